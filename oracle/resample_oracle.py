"""CPU oracle for the reference's image preprocessing — TEST INFRASTRUCTURE ONLY.

The reference's eval / analysis / linear-probing loaders apply ``Resize(size) -> CenterCrop(size) -> ToTensor ->
Normalize(ImageNet mean/std)`` and its train loader ``RandomResizedCrop(size) -> RandomHorizontalFlip -> ToTensor ->
Normalize`` (src/vitef/data/images/utils.py:337-366) to PIL images; CIFAR-10 samples are 32x32x3 uint8 arrays turned
into PIL images (src/vitef/data/images/cifar10.py:92-102). The arithmetic therefore lives in two third-party
dependencies that are NOT in /root/reference: torchvision (un-pinned in pyproject.toml; image has 0.26) and Pillow
(12.2.0 in this image), whose ``Image.resize(BILINEAR)`` is ``ImagingResample`` (Pillow src/libImaging/Resample.c):
separable, horizontal pass first, 8-bit intermediate, coefficients in 22-bit fixed point. This module restates that
published algorithm in numpy:

* ``precompute_coeffs``   = Resample.c ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (bilinear filter, support 1)
* ``resample_u8``         = ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``
* ``resized_crop_u8``     = torchvision ``F.resized_crop`` (crop, then resize the crop) + ``F.hflip``
* ``to_normalized_f32``   = ``ToTensor`` (uint8 -> float32 / 255, HWC -> CHW) + ``Normalize`` ((x - mean) / std in fp32)

Parity pinning: checked bit for bit against Pillow / torchvision themselves (``tests/test_resample_oracle.py`` does it at
test time wherever PIL is importable, and ``oracle/make_resample_golden.py`` stored PIL's own outputs in
``tests/golden/resample.npz``). Only tests, ``__graft_entry__.smoke()`` and bench baselines may import this module.
"""

from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: coefficients of 8-bit images are scaled by 2^22
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # src/vitef/data/images/utils.py:337
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bilinear(x: float) -> float:
    """Resample.c ``bilinear_filter`` (support 1.0)."""
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int, box0: float = 0.0, box1: float | None = None):
    """Resample.c ``precompute_coeffs`` for the bilinear filter followed by ``normalize_coeffs_8bpc``.

    Returns (ksize, bounds[out_size, 2] = (first source index, tap count), kk[out_size, ksize] int32 fixed-point taps).
    All floating-point work is in C ``double`` = Python float, in the same order as the C code.
    """
    if box1 is None:
        box1 = float(in_size)
    scale = filterscale = (box1 - box0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale  # bilinear support
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = box0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)  # C (int) cast truncates toward zero; the operand is >= -0.5 + 0.5
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            w = _bilinear((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx] = (xmin, xmax)
        for x in range(ksize):  # normalize_coeffs_8bpc: round half away from zero
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if k[x] < 0 else int(v + 0.5)
    return ksize, bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    """Resample.c ``clip8``: (in >> PRECISION_BITS) saturated to [0, 255]."""
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``Image.resize((out_w, out_h), BILINEAR)`` of an (H, W, C) uint8 image: horizontal pass, then vertical pass on the
    8-bit intermediate (ImagingResample; a pass is skipped when that dimension does not change)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    cur = img
    if out_w != w:
        _, bx, kx = precompute_coeffs(w, out_w)
        acc = np.full((h, out_w, img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for xx in range(out_w):
            x0, n = bx[xx]
            for x in range(n):
                acc[:, xx, :] += cur[:, x0 + x, :].astype(np.int64) * int(kx[xx, x])
        cur = _clip8(acc)
    if out_h != h:
        _, by, ky = precompute_coeffs(h, out_h)
        acc = np.full((out_h, cur.shape[1], img.shape[2]), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for yy in range(out_h):
            y0, n = by[yy]
            for y in range(n):
                acc[yy, :, :] += cur[y0 + y, :, :].astype(np.int64) * int(ky[yy, y])
        cur = _clip8(acc)
    return cur


def resized_crop_u8(img: np.ndarray, top: int, left: int, height: int, width: int, out: int, flip: bool = False) -> np.ndarray:
    """torchvision ``F.resized_crop`` on a PIL image = ``img.crop(box)`` then ``resize`` of the crop, then ``F.hflip``
    (RandomResizedCrop / RandomHorizontalFlip of the train transform, utils.py:341-346)."""
    res = resample_u8(np.ascontiguousarray(img[top : top + height, left : left + width]), out, out)
    return res[:, ::-1] if flip else res


def normalize_lut() -> np.ndarray:
    """lut[c, v] = fp32((fp32(v) / 255 - mean[c]) / std[c]): ToTensor's ``.div(255)`` and Normalize's ``sub_().div_()``,
    each rounded to fp32 as torch does. 3 x 256 entries cover every possible pixel."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)[:, None]
    std = np.asarray(IMAGENET_STD, dtype=np.float32)[:, None]
    return ((v[None, :] - mean) / std).astype(np.float32)


def to_normalized_f32(img_u8: np.ndarray) -> np.ndarray:
    """(H, W, 3) uint8 -> (3, H, W) float32, ``Normalize(ToTensor(img))``."""
    lut = normalize_lut()
    return np.stack([lut[c][img_u8[:, :, c]] for c in range(3)], axis=0)


def eval_transform(img_u8: np.ndarray, size: int = 224) -> np.ndarray:
    """The "val"/"test" transform for a square source (Resize(size) -> CenterCrop(size) is then a plain resize)."""
    assert img_u8.shape[0] == img_u8.shape[1], "non-square sources need the shorter-side resize + centre crop"
    return to_normalized_f32(resample_u8(img_u8, size, size))
