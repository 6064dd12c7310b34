"""GPU parity of the device-side input pipeline (vb_preprocess_u8) against the numpy oracle of Pillow's resample +
torchvision's ToTensor / Normalize: BIT-EXACT (integer resample, table-driven normalisation), plus the committed Pillow /
torchvision fixture. bf16 patch rows must equal im2col of the fp32 output rounded to bf16."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import resample_oracle as R

pytestmark = pytest.mark.gpu
GOLDEN = np.load(Path(__file__).parent / "golden" / "resample.npz")


def _pre(mode):
    from vit_plasticity_b200.preprocess import DevicePreprocessor

    return DevicePreprocessor(224, mode, "cuda")


def test_eval_transform_bit_exact_vs_oracle_and_fixture():
    rng = np.random.default_rng(0)
    imgs = np.concatenate([GOLDEN["imgs"], rng.integers(0, 256, (29, 32, 32, 3), dtype=np.uint8)])
    out = _pre("test")(torch.from_numpy(imgs)).cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(out[i], R.eval_transform(im)), i
    assert np.array_equal(out[2], GOLDEN["test_out"])  # torchvision's own output for fixture image 2


def test_train_transform_bit_exact_with_given_boxes():
    pre = _pre("train")
    imgs, boxes = GOLDEN["imgs"], GOLDEN["boxes"]
    _, _, _, index = pre._get_tables(32, 16)
    params = torch.tensor([[t, l, h, w, f, index[int(h)], index[int(w)], 0] for t, l, h, w, f in boxes.tolist()], dtype=torch.int32, device="cuda")
    out = pre(torch.from_numpy(imgs), params=params).cpu().numpy()
    for i in range(len(imgs)):
        assert np.array_equal(out[i], R.to_normalized_f32(GOLDEN["crops"][i])), i  # Pillow / torchvision's crops


def test_train_transform_random_boxes_and_edge_crops():
    pre = _pre("train")
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (64, 32, 32, 3), dtype=np.uint8)
    torch.manual_seed(11)
    params = pre.make_params(64, 32, 32)
    # force the degenerate boxes too: a single pixel, a single row, the whole image flipped
    params[0] = torch.tensor([31, 31, 1, 1, 0, 0, 0, 0], dtype=torch.int32)
    params[1] = torch.tensor([5, 0, 1, 32, 1, 0, 31, 0], dtype=torch.int32)
    params[2] = torch.tensor([0, 0, 32, 32, 1, 31, 31, 0], dtype=torch.int32)
    out = pre(torch.from_numpy(imgs), params=params).cpu().numpy()
    for i, (t, l, h, w, f, _, _, _) in enumerate(params.cpu().tolist()):
        ref = R.to_normalized_f32(R.resized_crop_u8(imgs[i], t, l, h, w, 224, bool(f)))
        assert np.array_equal(out[i], ref), (i, t, l, h, w, f)


def test_patch_rows_equal_im2col_of_the_image():
    from vit_plasticity_b200 import _lib as L

    pre = _pre("test")
    imgs = torch.from_numpy(np.random.default_rng(4).integers(0, 256, (9, 32, 32, 3), dtype=np.uint8))
    img = pre(imgs)
    patches = pre.patches(imgs)
    assert torch.equal(patches, L.im2col_patches(img, 16))


def test_model_accepts_patch_rows():
    from vit_plasticity_b200.models import build_model

    torch.manual_seed(0)
    model = build_model({"implementation": "vit", "model_name": "base", "pretrained": False, "finetuning": True, "n_classes": 10}, device="cuda")
    pre = _pre("test")
    imgs = torch.from_numpy(np.random.default_rng(5).integers(0, 256, (4, 32, 32, 3), dtype=np.uint8))
    with torch.no_grad():
        a = model(pre(imgs))
        b = model(pre.patches(imgs))
    assert torch.equal(a, b)
