"""Pin the oracle against the REAL reference and write the golden fixtures under tests/golden/.

Run in the build container only (it imports ``vitef`` and ``apps.vit`` from /root/reference, which does not exist on
the GPU box):

    python oracle/make_golden.py

For each configuration it (1) draws deterministic weights with ``vit_oracle.init_state_dict`` and loads them into the
reference ``nn.Module`` (``build_model`` from src/vitef/models/utils.py), (2) runs the reference's own forward,
``F.cross_entropy`` + ``backward`` + ``clip_grad_norm_`` + ``torch.optim.SGD`` step exactly as apps/vit/train.py
does, ``get_decomposition`` + ``distance`` exactly as apps/vit/analysis.py does, ``get_probes`` as
apps/vit/linear_probing.py does, (3) asserts the oracle reproduces every one of those to fp32 round-off, and
(4) stores the REFERENCE's outputs (never the oracle's) as the fixture. ``fire`` / ``omegaconf`` are absent from the
image and only used inside the apps' ``main()``: empty stub modules are injected before import.
"""

from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REF))
for missing in ("fire", "omegaconf"):
    if missing not in sys.modules:
        try:
            __import__(missing)
        except ImportError:
            stub = types.ModuleType(missing)
            stub.OmegaConf = object  # apps import `from omegaconf import OmegaConf`
            stub.Fire = lambda *a, **k: None
            sys.modules[missing] = stub

from apps.vit.analysis import distance as ref_distance  # noqa: E402
from apps.vit.utils import freeze_model as ref_freeze_model  # noqa: E402
from vitef.models import build_model as ref_build_model  # noqa: E402

from oracle import vit_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

CONFIGS = {
    # name: (arch, batch for train step, n pairs for plasticity, store full tensors?)
    "tiny": (O.Arch(emb_dim=128, n_heads=2, n_layers=2, ffn_dim=512, image_dim=(3, 32, 32)), 8, 6, True),
    "small": (O.Arch(emb_dim=256, n_heads=4, n_layers=3, ffn_dim=1024, image_dim=(3, 64, 64)), 6, 4, False),
    "vit_base": (O.vit_arch("base", n_classes=10), 4, 4, False),
    # BASELINE.json configs[3] / [4]: E = 1024, 16 heads, 24 layers (src/vitef/models/vit.py:130-134)
    "vit_large": (O.vit_arch("large", n_classes=10), 2, 2, False),
}
EPS_GRID = (10.0, 1.0, 1e-1, 1e-2, 1e-3)  # log grid of the perturbation sweep (modelled on apps/plots/loss_landscape.py:180-191)
SLAB = 4096  # leading contiguous elements of every large gradient stored verbatim (full-precision check of whole rows)
FREEZE_SETS = {  # BASELINE.json configs[2]; apps/vit/scripts/finetuning.sh:14
    "full": [],
    "attention_only": ["emb", "attn_norm", "ffn_norm", "ffn_fc1", "ffn_fc2"],
    "mlp_only": ["emb", "attn_norm", "mha", "ffn_norm"],
}
LR, MOMENTUM, CLIP = 1e-2, 0.9, 1.0  # apps/vit/configs/cifar10.yaml


def build_reference(name: str, arch: O.Arch):
    """Reference model + the prefix its inner Transformer's keys carry in state_dict()."""
    if name in ("vit_base", "vit_large"):
        cfg = dict(implementation="vit", model_name=name.split("_")[1], pretrained=False, in21k=True, patch_size=arch.patch_size,
                   image_dim=arch.image_dim, finetuning=True, n_classes=arch.n_classes)
        return ref_build_model(cfg, device="cpu"), "model."
    cfg = dict(implementation="transformer", image_dim=arch.image_dim, patch_type="computer_vision", image_patch="hybrid",
               patch_size=arch.patch_size, emb_type="linear", emb_dim=arch.emb_dim, pos_emb=True, n_heads=arch.n_heads,
               attn_bias=True, flash=False, causal=False, activation="gelu", ffn_dim=arch.ffn_dim, ffn_bias=True,
               norm="layer", norm_bias=True, norm_eps=arch.norm_eps, pre_norm=True, n_layers=arch.n_layers, dropout=0.0,
               cls_token=True, output_type="classification", weight_tying=False, n_classes=arch.n_classes)
    return ref_build_model(cfg, device="cpu"), ""


def checksum(t: torch.Tensor) -> list[float]:
    t = t.double().flatten()
    return [float(t.sum()), float(t.abs().sum()), float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64) % 7).sum())]


def sd_checksum(sd) -> dict[str, list[float]]:
    keys = sorted(sd)
    pick = keys[:: max(1, len(keys) // 12)]
    return {k: checksum(sd[k]) for k in pick}


def summarise(t: torch.Tensor, full: bool, slab: bool = False):
    t = t.detach().float()
    if full:
        return {"full": t.clone()}
    flat = t.flatten()
    if flat.numel() <= SLAB:
        return {"full": t.clone()}
    stride = max(1, flat.numel() // 256)
    out = {"norm": float(flat.double().norm()), "sample": flat[::stride][:256].clone(), "stride": stride}
    if slab:
        out["slab"] = flat[:SLAB].clone()
    return out


def assert_close(a, b, what, rtol=2e-4, atol=2e-5):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    if err > atol + rtol * scale:
        raise AssertionError(f"oracle != reference for {what}: max abs err {err:.3e} (scale {scale:.3e})")


def run_config(name: str):
    arch, batch, n_pairs, full = CONFIGS[name]
    torch.manual_seed(0)
    sd = O.init_state_dict(arch, seed=42)
    model, prefix = build_reference(name, arch)
    missing = model.load_state_dict({prefix + k: v for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    inner = model.model if prefix else model
    gold = {"name": name, "arch": arch.__dict__.copy(), "weights_seed": 42, "weights_checksum": sd_checksum(sd),
            "n_params": sum(v.numel() for v in sd.values()), "state_dict_keys": sorted(prefix + k for k in sd),
            "state_dict_shapes": {prefix + k: tuple(v.shape) for k, v in sd.items()}}

    # ---------------- forward + one finetuning step per freeze set (apps/vit/train.py:263-283) ----------------
    x = O.synthetic_images(batch, arch, seed=1)
    y = O.synthetic_labels(batch, arch, seed=2)
    gold["x_seed"], gold["y_seed"], gold["batch"] = 1, 2, batch
    gold["x_checksum"] = checksum(x)
    gold["labels"] = y.clone()
    model.eval()
    with torch.no_grad():
        logits_ref = model(x)
    assert_close(O.forward(sd, x, arch), logits_ref, f"{name}: logits")
    gold["logits"] = logits_ref.clone()
    gold["train"] = {}
    for fs_name, comps in FREEZE_SETS.items():
        model.load_state_dict({prefix + k: v for k, v in sd.items()})
        for p in model.parameters():
            p.requires_grad_(True)
            p.grad = None
        model.train()
        if prefix:
            ref_freeze_model(model, comps)  # apps/vit/utils.py:54 (expects the ViT wrapper)
        else:
            ref_freeze_model(model, [c for c in comps if c != "emb"])
            if "emb" in comps:  # freeze_model only reaches `.model.embedding` (apps/vit/utils.py:83)
                for p in model.embedding.parameters():
                    p.requires_grad = False
        opt = torch.optim.SGD(model.parameters(), lr=LR, weight_decay=0.0, momentum=MOMENTUM)  # optim.py:84-89
        preds = model(x)
        loss = F.cross_entropy(preds, y)
        loss.backward()
        grads_ref = {k[len(prefix):]: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), CLIP)
        opt.step()
        new_ref = {k[len(prefix):]: v.clone() for k, v in model.state_dict().items()}

        frozen = O.frozen_keys(sd, comps)
        o_loss, _, o_grads = O.loss_and_grads(sd, x, y, arch, frozen)
        assert set(o_grads) == set(grads_ref), f"{name}/{fs_name}: trainable sets differ: {set(o_grads) ^ set(grads_ref)}"
        assert_close(o_loss, loss, f"{name}/{fs_name}: loss")
        for k in grads_ref:
            assert_close(o_grads[k], grads_ref[k], f"{name}/{fs_name}: grad {k}", rtol=5e-4, atol=1e-6)
        sd2, bufs = dict(sd), {}
        o_norm = O.sgd_step(sd2, bufs, o_grads, LR, MOMENTUM, CLIP)
        assert_close(o_norm, gnorm, f"{name}/{fs_name}: grad_norm", rtol=5e-4)  # 24-layer fp32 round-off: 1.3e-4 on ViT-L
        for k in sd:
            assert_close(sd2[k], new_ref[k], f"{name}/{fs_name}: param after step {k}", rtol=1e-5, atol=1e-7)
        gold["train"][fs_name] = {
            "components": comps, "loss": float(loss), "grad_norm": float(gnorm), "trainable": sorted(grads_ref),
            "n_trainable": sum(v.numel() for v in grads_ref.values()),
            "grads": {k: summarise(v, full and fs_name == "full", slab=fs_name == "full") for k, v in grads_ref.items()},
            "param_delta_norm": {k: float((new_ref[k] - sd[k]).double().norm()) for k in grads_ref},
        }
        print(f"  [{name}/{fs_name}] loss {float(loss):.6f} grad_norm {float(gnorm):.6f} "
              f"trainable {gold['train'][fs_name]['n_trainable']}")

    # ---------------- plasticity estimator (apps/vit/analysis.py:216-233, apps/plots/analysis.py:97) ----------------
    model.load_state_dict({prefix + k: v for k, v in sd.items()})
    model.eval()
    x1 = O.synthetic_images(n_pairs, arch, seed=10)
    x2 = O.synthetic_images(n_pairs, arch, seed=11)
    out1, out2 = model.get_decomposition(x1), model.get_decomposition(x2)
    dist_ref = {k: ref_distance(out1[k], out2[k], reduction="none").numpy() for k in out1}
    dist_or = O.pair_distances(sd, x1, x2, arch)
    assert list(dist_ref) == list(dist_or), "decomposition key order differs"
    for k in dist_ref:
        assert_close(dist_or[k], dist_ref[k], f"{name}: distance {k}", rtol=2e-4)
    ratio_ref = {k: v / dist_ref["embedding"] for k, v in dist_ref.items() if k != "embedding"}  # plots/analysis.py:97
    plast = O.plasticity(dist_or)
    for comp, per_layer in plast.items():
        for i, r in enumerate(per_layer):
            assert_close(r, ratio_ref[f"block{i}_{comp}"], f"{name}: plasticity {comp}[{i}]", rtol=3e-4)
    gold["plasticity"] = {"x1_seed": 10, "x2_seed": 11, "n_pairs": n_pairs, "keys": list(dist_ref),
                          "distances": {k: torch.from_numpy(np.asarray(v)).clone() for k, v in dist_ref.items()},
                          "ratios": {k: torch.from_numpy(np.asarray(v)).clone() for k, v in ratio_ref.items()}}
    # small-perturbation pair (x, x + eps * noise): the regime of BASELINE.json configs[4]
    gold["plasticity_eps"] = {}
    noise = O.synthetic_images(n_pairs, arch, seed=12)
    # the same reference modules in float64: tells the reference's own fp32 round-off (which grows as 1 / eps) apart from
    # the error of the implementation under test; informational, the fp32 outputs above stay the fixture
    import copy

    model64 = copy.deepcopy(model).double()
    out1_64 = model64.get_decomposition(x1.double())
    gold["plasticity_eps_f64"] = {}
    for eps in EPS_GRID:
        xe = x1 + eps * noise
        oe = model.get_decomposition(xe)
        d = {k: ref_distance(out1[k], oe[k], reduction="none").numpy() for k in out1}
        gold["plasticity_eps"][eps] = {k: torch.from_numpy(np.asarray(v / d["embedding"])).clone() for k, v in d.items() if k != "embedding"}
        oe64 = model64.get_decomposition(x1.double() + eps * noise.double())
        d64 = {k: ref_distance(out1_64[k], oe64[k], reduction="none").numpy() for k in out1_64}
        gold["plasticity_eps_f64"][eps] = {k: torch.from_numpy(np.asarray(v / d64["embedding"])).clone() for k, v in d64.items() if k != "embedding"}
        worst = max(float(np.max(np.abs(gold["plasticity_eps"][eps][k].numpy() / gold["plasticity_eps_f64"][eps][k].numpy() - 1))) for k in gold["plasticity_eps"][eps])
        print(f"  [{name}] eps {eps:g}: reference fp32 vs fp64 ratios differ by at most {worst:.2e}")
    del model64, out1_64
    gold["plasticity_eps_noise_seed"] = 12

    # ---------------- probes (apps/vit/linear_probing.py:92-103: CLS row / token mean) ----------------
    if name not in ("vit_base", "vit_large"):
        pr = model.get_probes(x)
        pr_or = O.probes(sd, x, arch)
        assert list(pr) == list(pr_or)
        for k in pr:
            assert_close(pr_or[k], pr[k], f"{name}: probe {k}", rtol=5e-4, atol=1e-5)
        gold["probes"] = {k: {"cls": v[:, 0, :].clone(), "mean": v.mean(1).clone()} for k, v in pr.items()}

    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / f"{name}.pt"
    torch.save(gold, path)
    print(f"  wrote {path} ({path.stat().st_size / 1e6:.2f} MB)")


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "8")
    torch.set_num_threads(8)
    for cfg in sys.argv[1:] or list(CONFIGS):
        print(f"== {cfg}")
        run_config(cfg)
