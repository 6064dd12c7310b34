"""Golden vectors for the input pipeline, produced by Pillow / torchvision THEMSELVES (build container; both are
dependencies of the reference's loaders, src/vitef/data/images/utils.py:313-369 and data/images/cifar10.py:92-102):

    python oracle/make_resample_golden.py    ->  tests/golden/resample.npz

Stores, for seeded random 32x32 uint8 images: PIL's 224x224 bilinear resize, torchvision's full "test" transform output,
torchvision's RandomResizedCrop boxes (seeded) with the resized crops + flips, and one down-scaling case; and asserts
the numpy oracle reproduces each bit for bit before writing.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torchvision.transforms as T
import torchvision.transforms.functional as TF
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import resample_oracle as R  # noqa: E402

rng = np.random.default_rng(1234)
imgs = rng.integers(0, 256, (4, 32, 32, 3), dtype=np.uint8)
imgs[0] = 0
imgs[1] = 255
normalize = T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
test_tf = T.Compose([T.Resize(224), T.CenterCrop(224), T.ToTensor(), normalize])

resized = np.stack([np.asarray(Image.fromarray(im).resize((224, 224), Image.BILINEAR)) for im in imgs])
test_out = test_tf(Image.fromarray(imgs[2])).numpy()  # torchvision's full "test" transform of one image (fp32, 600 KB)
torch.manual_seed(7)
boxes, crops = [], []
for im in imgs:
    pil = Image.fromarray(im)
    top, left, h, w = T.RandomResizedCrop.get_params(pil, scale=(0.08, 1.0), ratio=(3 / 4, 4 / 3))
    flip = int(torch.rand(1).item() < 0.5)
    out = TF.resized_crop(pil, top, left, h, w, [224, 224])
    if flip:
        out = TF.hflip(out)
    boxes.append([top, left, h, w, flip])
    crops.append(np.asarray(out))
boxes, crops = np.asarray(boxes, dtype=np.int32), np.stack(crops)
big = rng.integers(0, 256, (75, 100, 3), dtype=np.uint8)
big_small = np.asarray(Image.fromarray(big).resize((40, 30), Image.BILINEAR))

for i, im in enumerate(imgs):
    assert np.array_equal(R.resample_u8(im, 224, 224), resized[i])
    t, l, h, w, f = boxes[i]
    assert np.array_equal(R.resized_crop_u8(im, t, l, h, w, 224, bool(f)), crops[i])
assert np.array_equal(R.eval_transform(imgs[2]), test_out)
assert np.array_equal(R.resample_u8(big, 30, 40), big_small)
np.savez_compressed(ROOT / "tests" / "golden" / "resample.npz", imgs=imgs, resized=resized, test_out=test_out, boxes=boxes, crops=crops, big=big, big_small=big_small)
print("oracle == Pillow / torchvision on every case; wrote tests/golden/resample.npz")
