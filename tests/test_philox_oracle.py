"""The noise oracle of the perturbation sweep: Philox4x32-10 pinned on the published known-answer vectors of Random123
(kat_vectors: zero / all-ones / pi-digit counter and key), plus sanity of the Box-Muller normals."""

import numpy as np

from oracle.philox_oracle import normal_images, philox4x32_10

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = tuple(int(x) for x in philox4x32_10(ctr, key))
        assert got == want, (ctr, key, [hex(g) for g in got])


def test_normals_are_standard_and_keyed_by_image_index():
    z = normal_images(3, 3 * 64 * 64, seed=5, first_image=10)
    assert z.dtype == np.float32 and z.shape == (3, 3 * 64 * 64)
    assert abs(float(z.mean())) < 1e-2 and abs(float(z.std()) - 1.0) < 1e-2 and np.isfinite(z).all()
    assert np.array_equal(z[1], normal_images(1, 3 * 64 * 64, seed=5, first_image=11)[0])  # batching invariant
    assert not np.array_equal(z[0], normal_images(1, 3 * 64 * 64, seed=6, first_image=10)[0])
    assert abs(float(np.corrcoef(z[0], z[1])[0, 1])) < 2e-2
