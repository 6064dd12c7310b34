"""CPU oracle of the device-side noise generator of the perturbation sweep — TEST INFRASTRUCTURE ONLY.

The reference draws its perturbation directions on the host with ``torch.rand`` / ``torch.randn``
(apps/plots/loss_landscape.py:172-191 style sweeps); there is no reference stream to reproduce bit for bit, so this
path defines its own counter-based generator: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11; Random123) keyed by the seed, with the IMAGE INDEX in the counter, so the noise of image i does not
depend on how images are batched or sharded over ranks. The integer stream is pinned on Random123's published
known-answer vectors (tests/test_philox_oracle.py); the normals are Box-Muller in float64 here and fp32 on the device.

    counter = (j, 0, image_lo, image_hi), key = (seed_lo, seed_hi)  ->  r0..r3  ->  elements 4j .. 4j+3 of the image
    u(r) = ((r >> 9) + 0.5) * 2^-23;  z0 = sqrt(-2 ln u(r0)) cos(2 pi u(r1)),  z1 = ... sin(...);  same for (r2, r3)
"""

from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 uint32 scalars -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) for x in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return [np.broadcast_to(x, np.broadcast(*c).shape).astype(np.uint32) for x in c]


def _unit(r):
    return ((r >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0**-23


def normal_images(n_images: int, elems_per_image: int, seed: int, first_image: int = 0) -> np.ndarray:
    """float32 [n_images, elems_per_image] standard normals; row i is image (first_image + i)."""
    assert elems_per_image % 4 == 0
    j = np.arange(elems_per_image // 4, dtype=np.uint64)
    out = np.empty((n_images, elems_per_image), np.float32)
    for i in range(n_images):
        img = first_image + i
        r = philox4x32_10((j, np.uint64(0), np.uint64(img & 0xFFFFFFFF), np.uint64(img >> 32)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        z = np.empty((elems_per_image // 4, 4), np.float64)
        for p in range(2):
            rad = np.sqrt(-2.0 * np.log(_unit(r[2 * p])))
            ang = 2.0 * np.pi * _unit(r[2 * p + 1])
            z[:, 2 * p], z[:, 2 * p + 1] = rad * np.cos(ang), rad * np.sin(ang)
        out[i] = z.reshape(-1).astype(np.float32)
    return out
