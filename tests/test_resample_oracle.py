"""CPU suite for the input pipeline: the numpy oracle against Pillow / torchvision's own outputs (committed fixture, and
live where PIL is importable), and the product's host-side logic (tap tables, Normalize table, crop sampling) against the
oracle / torchvision."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import resample_oracle as R

GOLDEN = np.load(Path(__file__).parent / "golden" / "resample.npz")


def test_oracle_matches_pillow_fixture():
    imgs, resized, boxes, crops = GOLDEN["imgs"], GOLDEN["resized"], GOLDEN["boxes"], GOLDEN["crops"]
    for i, im in enumerate(imgs):
        assert np.array_equal(R.resample_u8(im, 224, 224), resized[i])
        t, l, h, w, f = (int(v) for v in boxes[i])
        assert np.array_equal(R.resized_crop_u8(im, t, l, h, w, 224, bool(f)), crops[i])
    assert np.array_equal(R.eval_transform(imgs[2]), GOLDEN["test_out"])  # torchvision's whole "test" transform, fp32 bits
    assert np.array_equal(R.resample_u8(GOLDEN["big"], 30, 40), GOLDEN["big_small"])  # down-scaling: wider filter


def test_oracle_matches_pillow_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for h, w, oh, ow in [(32, 32, 224, 224), (17, 29, 224, 224), (1, 1, 8, 8), (64, 48, 20, 31), (3, 200, 224, 224)]:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(im).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(R.resample_u8(im, oh, ow), ref), (h, w, oh, ow)


def test_product_tap_tables_equal_oracle():
    from vit_plasticity_b200.preprocess import bilinear_taps

    for in_size in list(range(1, 40)) + [100, 224, 256, 375, 500]:
        for out_size in (224, 32, 7):
            ks, bounds, kk = R.precompute_coeffs(in_size, out_size)
            ks2, b2, k2 = bilinear_taps(in_size, out_size)
            assert ks == ks2
            assert np.array_equal(bounds, np.asarray(b2, dtype=np.int32))
            assert np.array_equal(kk, np.asarray(k2, dtype=np.int32))


def test_product_normalize_table_equals_oracle_and_torch():
    from vit_plasticity_b200.preprocess import normalize_table

    lut = normalize_table("cpu").numpy()
    assert np.array_equal(lut, R.normalize_lut())
    # and both equal what ToTensor + Normalize do to every byte value
    v = torch.arange(256, dtype=torch.uint8).view(1, 256, 1).expand(3, 256, 1).contiguous()
    t = v.to(torch.float32).div(255)
    mean = torch.tensor(R.IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(R.IMAGENET_STD).view(3, 1, 1)
    assert np.array_equal(lut, t.sub(mean).div(std).view(3, 256).numpy())


def test_product_crop_sampling_follows_torchvision():
    T = pytest.importorskip("torchvision.transforms")
    from vit_plasticity_b200.preprocess import random_resized_crop_params

    img = torch.zeros(3, 32, 32)
    for seed in range(5):
        torch.manual_seed(seed)
        ref = [T.RandomResizedCrop.get_params(img, scale=(0.08, 1.0), ratio=(3 / 4, 4 / 3)) + (float(torch.rand(1)),) for _ in range(20)]
        torch.manual_seed(seed)
        got = [random_resized_crop_params(32, 32) + (float(torch.rand(1)),) for _ in range(20)]
        assert ref == got
