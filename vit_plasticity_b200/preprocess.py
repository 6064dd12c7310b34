"""Device-side input pipeline: the reference's image transforms applied on the GPU to uint8 batches.

Reference: ``build_transform`` (src/vitef/data/images/utils.py:313-369) composes torchvision transforms over PIL
images, per sample, in the DataLoader process (single process for the train loader, utils.py:197):

* ``"val"`` / ``"test"`` (eval.py, analysis.py, linear_probing.py): ``Resize(size) -> CenterCrop(size) -> ToTensor ->
  Normalize(ImageNet mean / std)``;
* ``"train"`` (train.py): ``RandomResizedCrop(size) -> RandomHorizontalFlip -> ToTensor -> Normalize``.

Here the uint8 batch (CIFAR-10: 32 x 32 x 3, 3 KB per image) is copied to the GPU as it is and ONE kernel
(``vb_preprocess_u8``) produces either the fp32 NCHW tensor the reference model is fed or directly the bf16 patch rows of
the patch-embedding GEMM. The result is bit-identical to PIL + torchvision: the resize is Pillow's two-pass 8-bit
bilinear ``ImagingResample`` driven by the fixed-point tap tables computed below (host side, float64, once per source
size), and ``Normalize(ToTensor(v))`` is a 3 x 256 table evaluated with torch's own fp32 ops. The random crop boxes and
flips are drawn on the host with torch's global generator in the order torchvision draws them, so a seeded run sees the
same augmentations as the reference loader.
"""

from __future__ import annotations

import math

import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # data/images/utils.py:337
IMAGENET_STD = (0.229, 0.224, 0.225)
_PRECISION_BITS = 22  # Pillow: 32 - 8 - 2 for 8-bit channels


def bilinear_taps(in_size: int, out_size: int) -> tuple[int, list[tuple[int, int]], list[list[int]]]:
    """Tap table of Pillow's bilinear resample of ``in_size`` -> ``out_size`` samples: (ksize, [(first, count)], [[taps]])
    with taps scaled by 2^22 and rounded half away from zero. Up- and down-scaling (the filter widens by the scale)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = fscale  # bilinear: support 1, stretched when shrinking
    ksize = 2 * math.ceil(support) + 1
    bounds, taps = [], []
    for i in range(out_size):
        center = (i + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        inv = 1.0 / fscale  # Pillow multiplies by the reciprocal
        w = [max(0.0, 1.0 - abs((j - center + 0.5) * inv)) for j in range(lo, hi)]
        total = 0.0
        for v in w:
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        fixed = [int(v * (1 << _PRECISION_BITS) + 0.5) for v in w]
        bounds.append((lo, hi - lo))
        taps.append(fixed + [0] * (ksize - len(fixed)))
    return ksize, bounds, taps


def normalize_table(device) -> torch.Tensor:
    """lut[c, v] = Normalize(ToTensor(v))[c], built with the same fp32 torch ops the reference's transforms run."""
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)  # ToTensor
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1)
    return v.view(1, 256).repeat(3, 1).sub_(mean).div_(std).contiguous().to(device)  # Normalize


def random_resized_crop_params(height: int, width: int, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)) -> tuple[int, int, int, int]:
    """torchvision ``RandomResizedCrop.get_params``: same draws from torch's global generator, in the same order."""
    area = height * width
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect = math.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1]).item())
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)  # fallback: central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


class DevicePreprocessor:
    """``build_transform(size, mode)`` for batches of equally sized uint8 HWC images, on the GPU.

    ``mode``: "train" | "val" | "test" (data/images/utils.py:338-367). ``__call__(batch_u8)`` takes a uint8 tensor
    [n, H, W, 3] (host or device; host tensors are copied as uint8) and returns the fp32 NCHW batch; ``patches(batch_u8)``
    returns the bf16 patch rows for ``Embedding.forward_patches`` instead (no fp32 image is materialised).
    """

    def __init__(self, size: int = 224, mode: str = "test", device="cuda", patch: int = 16):
        mode = mode.lower()
        if mode not in ("train", "val", "test"):
            raise ValueError(f"Mode {mode} not found. Options are 'train', 'val' and 'test'.")
        self.size, self.mode, self.device, self.patch = size, mode, torch.device(device), patch
        self.lut = normalize_table(self.device)
        self._tables: dict[tuple[int, int], tuple] = {}

    # one table per source extent 1..max_side (train crops can have any height / width up to the image's)
    def _get_tables(self, max_side: int, strip: int):
        key = (max_side, strip)
        if key not in self._tables:
            sizes = range(1, max_side + 1) if self.mode == "train" else [max_side]
            all_b, all_k, ksize, max_rows = [], [], 0, 1
            tabs = [bilinear_taps(s, self.size) for s in sizes]
            ksize = max(t[0] for t in tabs)
            for _, b, k in tabs:
                all_b.append(b)
                all_k.append([row + [0] * (ksize - len(row)) for row in k])
                for o0 in range(0, self.size, strip):
                    o1 = min(o0 + strip, self.size) - 1
                    max_rows = max(max_rows, b[o1][0] + b[o1][1] - b[o0][0])
            bounds = torch.tensor(all_b, dtype=torch.int32, device=self.device).contiguous()
            coef = torch.tensor(all_k, dtype=torch.int32, device=self.device).contiguous()
            self._tables[key] = (bounds, coef, max_rows, {s: i for i, s in enumerate(sizes)})
        return self._tables[key]

    def _params(self, n: int, h: int, w: int, index: dict[int, int]) -> torch.Tensor | None:
        if self.mode != "train":
            return None
        rows = []
        for _ in range(n):
            top, left, ch, cw = random_resized_crop_params(h, w)  # RandomResizedCrop(size) defaults
            flip = int(torch.rand(1).item() < 0.5)  # RandomHorizontalFlip(p=0.5)
            rows.append([top, left, ch, cw, flip, index[ch], index[cw], 0])
        return torch.tensor(rows, dtype=torch.int32).to(self.device)

    def _run(self, batch_u8: torch.Tensor, want_f32: bool, patch: int, params: torch.Tensor | None = None):
        if batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4 or batch_u8.shape[3] != 3:
            raise ValueError("expected a uint8 batch [n, H, W, 3]")
        n, h, w, _ = batch_u8.shape
        if self.mode != "train" and h != w:
            raise NotImplementedError("non-square sources need the shorter-side Resize + CenterCrop window")
        strip = patch if patch else 16
        bounds, coef, max_rows, index = self._get_tables(max(h, w), strip)
        if params is None:
            params = self._params(n, h, w, index)
        src = batch_u8.to(self.device, non_blocking=True).contiguous()
        return L.preprocess_u8(src, params, bounds, coef, max_rows, self.lut, self.size, want_f32=want_f32, patch=patch)

    def __call__(self, batch_u8: torch.Tensor, params: torch.Tensor | None = None) -> torch.Tensor:
        return self._run(batch_u8, True, 0, params)[0]

    def patches(self, batch_u8: torch.Tensor, params: torch.Tensor | None = None) -> torch.Tensor:
        return self._run(batch_u8, False, self.patch, params)[1]

    def make_params(self, n: int, h: int, w: int) -> torch.Tensor | None:
        """The per-sample crop / flip table of the next batch (train mode), drawn from torch's global generator."""
        _, _, _, index = self._get_tables(max(h, w), 16)
        return self._params(n, h, w, index)
