"""Model-level parity on the GPU: the CUDA path (through the nn.Module drop-in surface and the C-ABI) against
(a) the committed fixtures = outputs of the REAL reference in fp32, and (b) the CPU oracle on the same seeded inputs.

Stated tolerances (bf16 storage / fp32 accumulate vs the reference's fp32):
  logits        relative L2 error <= 2e-2   (observed bf16-vs-fp32 max abs 0.02 on ~1-magnitude logits)
  loss          abs <= 2e-2
  gradients     per-tensor relative L2 error <= 4e-2 over the FULL tensor (tiny / small fixtures hold full tensors; for
                ViT-B/16 the full tensors are compared with the CPU oracle in test_against_cpu_oracle_same_inputs);
                total grad-norm within 2e-2 relative. The ViT-B/16 fixture holds 256 strided samples + the norm per
                tensor: that sampled estimate of the same relative error has ~±15 % estimator noise on top of the
                2-3.5 % bf16 error of the deepest fc1 / qkv weights, so it is held to 6e-2.
                Where the fixture also holds a "slab" (the first 4096 contiguous elements of a large gradient, stored
                verbatim), the relative L2 error over the slab is held to 5e-2 (ViT-B/16) / 1e-1 (ViT-L/16: 24 layers).
  plasticity    ratios to the embedding distance within 5e-3 relative (attention), 1e-3 (LayerNorm / fc1 / fc2), for
                independent pairs AND for perturbation pairs (x, x + eps n) at every eps of the fixture grid
                {10, 1, 1e-1, 1e-2, 1e-3} — the estimator carries (x, d), so its accuracy does not depend on eps. The
                fixtures are the reference's fp32 outputs, whose own round-off is <= 2e-4 at eps = 1e-3 (the fixture
                also holds the same reference modules run in float64 to tell the two apart; reported, not asserted).
"""

import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
DEV = "cuda"
REPORT = {}
VIT_NAMES = ("vit_base", "vit_large")  # fixtures built through the ViT wrapper (state_dict keys carry "model.")


def _dump_report():
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / "parity_report.json").write_text(json.dumps(REPORT, indent=1, sort_keys=True))
    except OSError:
        pass


def load(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


def arch_of(gold):
    a = dict(gold["arch"])
    a["image_dim"] = tuple(a["image_dim"])
    return O.Arch(**a)


def rel_l2(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def build(name, gold, arch, sd):
    from vit_plasticity_b200 import build_model

    if name in VIT_NAMES:
        cfg = dict(implementation="vit", model_name=name.split("_")[1], pretrained=False, in21k=True, finetuning=True, n_classes=arch.n_classes)
        model = build_model(cfg, device=DEV)
        model.load_state_dict({"model." + k: v for k, v in sd.items()})
    else:
        cfg = dict(implementation="transformer", image_dim=arch.image_dim, patch_type="computer_vision", image_patch="hybrid",
                   patch_size=arch.patch_size, emb_type="linear", emb_dim=arch.emb_dim, pos_emb=True, n_heads=arch.n_heads,
                   attn_bias=True, activation="gelu", ffn_dim=arch.ffn_dim, ffn_bias=True, norm="layer", norm_bias=True,
                   norm_eps=arch.norm_eps, pre_norm=True, n_layers=arch.n_layers, cls_token=True, output_type="classification",
                   weight_tying=False, n_classes=arch.n_classes)
        model = build_model(cfg, device=DEV)
        model.load_state_dict(sd)
    assert sorted(model.state_dict()) == gold["state_dict_keys"]
    return model


def check_summary(got, summ, tol, what):
    got = got.detach().float().cpu()
    if "full" in summ:
        err = rel_l2(got, summ["full"])
    else:
        sample = got.flatten()[:: summ["stride"]][:256]
        # sampled relative error, normalised by the full tensor's RMS so that near-zero samples do not dominate
        rms = summ["norm"] / max(1.0, got.numel() ** 0.5)
        err = float((sample.double() - summ["sample"].double()).norm() / (256**0.5 * rms + 1e-30))
        nerr = abs(float(got.double().norm()) - summ["norm"]) / (summ["norm"] + 1e-30)
        err = max(err, nerr)
        if "slab" in summ:  # whole leading rows stored verbatim: a real relative L2 error, not a sampled estimate
            slab = summ["slab"]
            serr = rel_l2(got.flatten()[: slab.numel()], slab)
            REPORT.setdefault("grad_slab_rel_l2", {})[what] = serr
            # 4 leading rows of a matrix are a local sample: their relative error sits above the whole tensor's (measured
            # worst 3.7e-2 on ViT-B/16; 5.5e-2 on ViT-L/16, block 0 query rows, 24 bf16 layers of back-propagation behind
            # them, whose full tensor is within 4e-2 of the oracle in test_against_cpu_oracle_same_inputs)
            stol = 1e-1 if what.startswith("vit_large") else 5e-2
            assert serr <= stol, f"{what}: slab relative L2 error {serr:.3e} > {stol}"
    REPORT.setdefault("grad_rel_l2", {})[what] = err
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol}"
    return err


@pytest.mark.parametrize("name", ["tiny", "small", "vit_base", "vit_large"])
def test_logits_loss_grads_match_reference(name):
    from vit_plasticity_b200.finetune import freeze_model

    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build(name, gold, arch, sd)
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"]).to(DEV)
    y = O.synthetic_labels(gold["batch"], arch, gold["y_seed"]).to(DEV)
    model.train()
    prefix = "model." if name in VIT_NAMES else ""
    for fs, ref in gold["train"].items():
        for p in model.parameters():
            p.requires_grad_(True)
            p.grad = None
        freeze_model(model, ref["components"]) if prefix else _freeze_inner(model, ref["components"])
        logits = model(x)
        assert logits.dtype == torch.float32 and logits.shape == gold["logits"].shape
        e_log = rel_l2(logits, gold["logits"])
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        grads = {k[len(prefix):]: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert sorted(grads) == ref["trainable"], f"{name}/{fs}: trainable set differs"
        assert sum(g.numel() for g in grads.values()) == ref["n_trainable"]
        gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
        REPORT[f"{name}/{fs}"] = {"logits_rel_l2": e_log, "loss_abs": abs(float(loss) - ref["loss"]), "grad_norm_rel": abs(gnorm - ref["grad_norm"]) / ref["grad_norm"]}
        assert e_log <= 2e-2, f"{name}/{fs}: logits rel L2 {e_log:.3e}"
        assert abs(float(loss) - ref["loss"]) <= 2e-2
        assert abs(gnorm - ref["grad_norm"]) <= 2e-2 * ref["grad_norm"], f"grad norm {gnorm} vs {ref['grad_norm']}"
        worst = max(check_summary(g, ref["grads"][k], 4e-2 if "full" in ref["grads"][k] else 6e-2, f"{name}/{fs}/{k}") for k, g in grads.items())
        REPORT[f"{name}/{fs}"]["worst_grad_rel_l2"] = worst
    _dump_report()


def _freeze_inner(model, comps):
    from vit_plasticity_b200.finetune import freeze_model

    freeze_model(model, comps)


@pytest.mark.parametrize("name", ["tiny", "small", "vit_base", "vit_large"])
def test_against_cpu_oracle_same_inputs(name):
    """Same check against the oracle run here on the host (different seeds than the fixture); full tensors, ViT-B/16
    (4 images) and ViT-L/16 (2 images) included: a few seconds of fp32 CPU work each."""
    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=7)
    model = build(name, gold, arch, sd)
    nb = {"vit_base": 4, "vit_large": 2}.get(name, 5)
    x = O.synthetic_images(nb, arch, 21)
    y = O.synthetic_labels(nb, arch, 22)
    o_loss, o_logits, o_grads = O.loss_and_grads(sd, x, y, arch)
    model.train()
    logits = model(x.to(DEV))
    loss = torch.nn.functional.cross_entropy(logits, y.to(DEV))
    loss.backward()
    assert rel_l2(logits, o_logits) <= 2e-2
    assert abs(float(loss) - float(o_loss)) <= 2e-2
    prefix = "model." if name in VIT_NAMES else ""
    errs = {k: rel_l2(p.grad, o_grads[k[len(prefix):]]) for k, p in model.named_parameters()}
    REPORT[f"{name}/oracle_full_tensor"] = {"worst_grad_rel_l2": max(errs.values()), "worst_tensor": max(errs, key=errs.get),
                                            "logits_rel_l2": rel_l2(logits, o_logits)}
    _dump_report()
    for k, e in errs.items():
        assert e <= 4e-2, f"{k}: {e:.3e}"


@pytest.mark.parametrize("name", ["tiny", "small", "vit_base"])
def test_train_step_matches_reference(name):
    from vit_plasticity_b200.finetune import build_optimizer, train_step

    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build(name, gold, arch, sd)
    prefix = "model." if name in VIT_NAMES else ""
    ref = gold["train"]["full"]
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"]).to(DEV)
    y = O.synthetic_labels(gold["batch"], arch, gold["y_seed"]).to(DEV)
    model.train()
    opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9)
    loss, gnorm = train_step(model, opt, [(x, y)], grad_clip=1.0)
    assert abs(float(loss) - ref["loss"]) <= 2e-2
    assert abs(float(gnorm) - ref["grad_norm"]) <= 2e-2 * ref["grad_norm"]
    new = {k[len(prefix):]: v for k, v in model.state_dict().items()}
    worst = 0.0
    for k, dn in ref["param_delta_norm"].items():
        got = float((new[k].cpu().double() - sd[k].double()).norm())
        worst = max(worst, abs(got - dn) / (dn + 1e-30))
    REPORT[f"{name}/train_step_worst_delta_norm_rel"] = worst
    assert worst <= 5e-2
    assert all(p.grad is None for p in model.parameters())  # optimizer.zero_grad() ran
    # two micro-batches with grad accumulation == one batch of the concatenation (mean-of-means, equal sizes)
    model.load_state_dict({prefix + k: v for k, v in sd.items()})
    opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9)
    h = gold["batch"] // 2
    loss2, gnorm2 = train_step(model, opt, [(x[:h], y[:h]), (x[h : 2 * h], y[h : 2 * h])], grad_clip=1.0)
    if 2 * h == gold["batch"]:
        assert abs(float(gnorm2) - ref["grad_norm"]) <= 3e-2 * ref["grad_norm"]
    _dump_report()


@pytest.mark.parametrize("components", [[], ["emb", "attn_norm", "ffn_norm", "ffn_fc1", "ffn_fc2"]])
def test_fused_sgd_matches_torch_sgd(components):
    """FusedSGD (flat arena, in-place wgrad accumulation, one clip+momentum+update pass) against the reference sequence
    clip_grad_norm_ + torch.optim.SGD on the same model, three steps incl. one with two accumulated micro-batches."""
    from vit_plasticity_b200.finetune import FusedSGD, build_optimizer, freeze_model, train_step

    gold = load("small")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    models = [build("small", gold, arch, sd), build("small", gold, arch, sd)]
    for m in models:
        m.train()
        freeze_model(m, components)
    opts = [build_optimizer(models[0], "sgd", lr=1e-2, momentum=0.9), build_optimizer(models[1], "sgd", lr=1e-2, momentum=0.9, fused=True)]
    assert isinstance(opts[1], FusedSGD)
    xs = [O.synthetic_images(4, arch, 40 + i).to(DEV) for i in range(4)]
    ys = [O.synthetic_labels(4, arch, 50 + i).to(DEV) for i in range(4)]
    steps = [[(xs[0], ys[0])], [(xs[1], ys[1]), (xs[2], ys[2])], [(xs[3], ys[3])]]
    for batches in steps:
        out = [train_step(m, o, batches, grad_clip=1.0) for m, o in zip(models, opts)]
        assert abs(float(out[0][0]) - float(out[1][0])) <= 1e-4 * max(1.0, abs(float(out[0][0])))
        assert abs(float(out[0][1]) - float(out[1][1])) <= 1e-4 * float(out[0][1]), (float(out[0][1]), float(out[1][1]))
    for (k, a), (_, b) in zip(models[0].state_dict().items(), models[1].state_dict().items()):
        assert rel_l2(b, a) <= 2e-5, (k, rel_l2(b, a))
        if k.split(".")[0] == "embedding" and "emb" in components:
            assert torch.equal(a.cpu(), sd[k]), k  # frozen parameters untouched
    # gradients stay views of the arena (zeroed, not dropped)
    arena = opts[1].arena
    assert float(arena.abs().max()) == 0.0
    assert all(p.grad is not None and p.grad.data_ptr() >= arena.data_ptr() for p in opts[1].trainable)


def test_fused_sgd_train_step_matches_reference_fixture():
    from vit_plasticity_b200.finetune import build_optimizer, train_step

    gold = load("vit_base")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build("vit_base", gold, arch, sd)
    ref = gold["train"]["full"]
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"]).to(DEV)
    y = O.synthetic_labels(gold["batch"], arch, gold["y_seed"]).to(DEV)
    model.train()
    opt = build_optimizer(model, "sgd", lr=1e-2, momentum=0.9, fused=True)
    loss, gnorm = train_step(model, opt, [(x, y)], grad_clip=1.0)
    assert abs(float(loss) - ref["loss"]) <= 2e-2
    assert abs(float(gnorm) - ref["grad_norm"]) <= 2e-2 * ref["grad_norm"]
    new = {k[len("model."):]: v for k, v in model.state_dict().items()}
    worst = max(abs(float((new[k].cpu().double() - sd[k].double()).norm()) - dn) / (dn + 1e-30) for k, dn in ref["param_delta_norm"].items())
    REPORT["vit_base/fused_sgd_worst_delta_norm_rel"] = worst
    assert worst <= 5e-2
    _dump_report()


@pytest.mark.parametrize("name", ["tiny", "small", "vit_base", "vit_large"])
def test_plasticity_matches_reference(name):
    from vit_plasticity_b200.plasticity import PlasticityEstimator, get_plasticity

    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build(name, gold, arch, sd).eval()
    p = gold["plasticity"]
    x1 = O.synthetic_images(p["n_pairs"], arch, p["x1_seed"]).to(DEV)
    x2 = O.synthetic_images(p["n_pairs"], arch, p["x2_seed"]).to(DEV)
    est = PlasticityEstimator(model)
    dist = est.pair_distances(x1, x2)
    assert list(dist) == p["keys"]
    worst = {}
    for k, ref in p["distances"].items():
        comp = k.split("_", 1)[1] if "_" in k else k
        err = float(np.max(np.abs(dist[k] - ref.numpy()) / ref.numpy()))
        worst[comp] = max(worst.get(comp, 0.0), err)
    ratios = get_plasticity(dist)
    for comp, per_layer in ratios.items():
        tol = 5e-3 if comp == "attn" else 1e-3
        for i, r in enumerate(per_layer):
            ref = p["ratios"][f"block{i}_{comp}"].numpy()
            err = float(np.max(np.abs(r - ref) / ref))
            worst["ratio_" + comp] = max(worst.get("ratio_" + comp, 0.0), err)
            assert err <= tol, f"{name}: plasticity {comp}[{i}] rel err {err:.3e} > {tol}"
    REPORT[f"{name}/plasticity_worst_rel"] = worst
    # size-independent properties: identical inputs -> all distances exactly 0; symmetry in (x1, x2)
    same = est.pair_distances(x1, x1)
    nonzero = {k: float(np.abs(v).max()) for k, v in same.items() if float(np.abs(v).max()) != 0.0}
    assert not nonzero, f"identical inputs must give exactly zero distances, got {nonzero}"
    swapped = est.pair_distances(x2, x1)  # the pair is carried as (base, difference): swapping changes the base point
    for k in dist:
        assert np.allclose(swapped[k], dist[k], rtol=5e-3 if k.endswith("_attn") else 1e-3), k
    # perturbation pairs x' = x + eps * noise over the whole grid (fixture: the fp32 reference's ratios)
    noise = O.synthetic_images(p["n_pairs"], arch, gold["plasticity_eps_noise_seed"]).to(DEV)
    for eps, ref_r in gold["plasticity_eps"].items():
        d = est.pair_distances(x1, x1 + eps * noise)
        errs, errs64 = {}, {}
        for k, r in ref_r.items():
            comp = k.split("_", 1)[1]
            got = d[k] / d["embedding"]
            errs[comp] = max(errs.get(comp, 0.0), float(np.max(np.abs(got - r.numpy()) / r.numpy())))
            r64 = gold["plasticity_eps_f64"][eps][k].numpy()
            errs64[comp] = max(errs64.get(comp, 0.0), float(np.max(np.abs(got - r64) / r64)))
        REPORT[f"{name}/plasticity_eps_{eps}"] = errs
        REPORT[f"{name}/plasticity_eps_{eps}_vs_f64_reference"] = errs64
        for comp, err in errs.items():
            tol = 5e-3 if comp == "attn" else 1e-3  # measured worst: 3.6e-3 / 6.8e-4 (tiny); ViT-B/L: 3.9e-4 / 3.6e-5
            assert err <= tol, f"{name}: eps={eps} {comp} rel err {err:.3e} > {tol}"
    _dump_report()


def test_perturbation_sweep_matches_oracle_and_is_shard_invariant():
    """configs[4] driver on one GPU: ratios against the CPU oracle on the same (x, x + eps n) pairs over a log grid down to
    1e-3 (n from the device generator, reproduced on the host by oracle/philox_oracle.py); running the two halves as the
    two ranks of a world-size-2 job gives the same table (noise is keyed by image index, shards are contiguous)."""
    from oracle.philox_oracle import normal_images

    from vit_plasticity_b200.plasticity import PlasticityEstimator, perturbation_sweep

    gold = load("small")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build("small", gold, arch, sd).eval()
    n = 6
    x = O.synthetic_images(n, arch, 31)
    eps_list = [10.0, 1.0, 0.1, 1e-2, 1e-3]
    full = perturbation_sweep(model, x, eps_list, noise_seed=5, pairs_per_call=4)
    noise = torch.from_numpy(normal_images(n, x[0].numel(), 5, 0)).reshape(x.shape)
    sd64 = {k: v.double() for k, v in sd.items()}
    for eps in eps_list:
        # float64 oracle: at eps = 1e-3 the fp32 oracle's own round-off (~2e-4) would eat into the tolerance
        ref = O.pair_distances(sd64, x.double(), x.double() + eps * noise.double(), arch)
        worst = {}
        for k, v in ref.items():
            if k == "embedding":
                continue
            r_ref = v / ref["embedding"]
            r = full[eps][k] / full[eps]["embedding"]
            comp = k.split("_", 1)[1]
            worst[comp] = max(worst.get(comp, 0.0), float(np.max(np.abs(r - r_ref) / r_ref)))
        REPORT[f"small/sweep_eps_{eps}"] = worst
        for comp, err in worst.items():
            tol = 5e-3 if comp == "attn" else 1e-3
            assert err <= tol, (eps, comp, err)
    est = PlasticityEstimator(model)
    halves = [perturbation_sweep(model, x, eps_list, noise_seed=5, pairs_per_call=4, rank=r, world=2, estimator=est) for r in range(2)]
    for eps in eps_list:
        for k in full[eps]:
            both = np.concatenate([halves[0][eps][k], halves[1][eps][k]])
            assert np.allclose(both, full[eps][k], rtol=1e-5, atol=0), (eps, k)
    _dump_report()


def test_analysis_loop_writes_reference_distances_pkl(tmp_path):
    """analysis(): the reference's loop and on-disk format (apps/vit/analysis.py:203-248) on the fused estimator."""
    import pickle

    from vit_plasticity_b200.plasticity import analysis, get_plasticity

    gold = load("tiny")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build("tiny", gold, arch, sd).eval()
    mk = lambda seed: [(O.synthetic_images(3, arch, seed + i), torch.zeros(3, dtype=torch.long)) for i in range(2)]
    l1, l2 = mk(60), mk(70)
    dist = analysis(model, l1, l2, n_steps=3, save_dir=tmp_path)  # 3 steps over 2-batch loaders: the loaders wrap around
    saved = pickle.load(open(tmp_path / "distances.pkl", "rb"))
    assert list(saved) == list(dist) and len(saved) == 1 + 5 * arch.n_layers
    for k, v in saved.items():
        assert isinstance(v, np.ndarray) and v.dtype == np.float32 and v.shape == (9,)
    ref = O.pair_distances(sd, l1[1][0], l2[1][0], arch)  # second batch pair sits at positions 3..5
    for k in ref:
        assert np.allclose(saved[k][3:6] / saved["embedding"][3:6], ref[k] / ref["embedding"], rtol=2e-2), k
    assert np.allclose(saved["embedding"][6:9], saved["embedding"][0:3])  # wrap-around: third step = first batch pair again
    plast = get_plasticity(saved)
    assert set(plast) == {"attn_norm", "attn", "ffn_norm", "ffn_fc1", "ffn_fc2"} and all(len(v) == arch.n_layers for v in plast.values())


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_decomposition_and_probes_api(name):
    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    model = build(name, gold, arch, sd).eval()
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"])
    probes = model.get_probes(x.to(DEV))
    assert list(probes) == list(gold["probes"])
    for k, v in probes.items():
        assert v.device.type == "cpu" and v.dtype == torch.float32
        assert rel_l2(v[:, 0, :], gold["probes"][k]["cls"]) <= 3e-2, k
        assert rel_l2(v.mean(1), gold["probes"][k]["mean"]) <= 3e-2, k
    # on-device pooling for linear probing == pooling + row normalisation of the reference's taps
    for cls_pooling, field in ((True, "cls"), (False, "mean")):
        pooled = model.get_pooled_probes(x.to(DEV), cls_pooling=cls_pooling)
        assert list(pooled) == list(gold["probes"])
        for k, v in pooled.items():
            ref_rows = gold["probes"][k][field].float()
            ref_rows = ref_rows / ref_rows.norm(dim=-1, keepdim=True)
            assert v.device.type == "cpu" and v.shape == ref_rows.shape
            assert rel_l2(v, ref_rows) <= 3e-2, (k, field)
            assert torch.allclose(v.norm(dim=-1), torch.ones(v.shape[0]), atol=1e-4)
    # the linear-probing driver (apps/vit/linear_probing.py:58-116 get_embeddings): two batches, loader order kept
    from vit_plasticity_b200.probing import get_embeddings

    h = gold["batch"] // 2
    labels = torch.arange(gold["batch"])
    emb, lab = get_embeddings(model, [(x[:h], labels[:h]), (x[h:], labels[h:])], cls_pooling=True)
    assert np.array_equal(lab, labels.numpy()) and list(emb) == list(gold["probes"])
    for k, v in emb.items():
        ref_rows = gold["probes"][k]["cls"].float()
        ref_rows = (ref_rows / ref_rows.norm(dim=-1, keepdim=True)).numpy()
        assert v.dtype == np.float32 and v.shape == ref_rows.shape
        assert rel_l2(v, ref_rows) <= 3e-2, k
    dec = model.get_decomposition(x.to(DEV))
    ref = O.decomposition(sd, x, arch)
    assert list(dec) == list(ref) and len(dec) == 1 + 5 * arch.n_layers
    for k, v in dec.items():
        assert v.device.type == "cpu" and v.shape == ref[k].shape
        assert rel_l2(v, ref[k].detach()) <= 2e-2, k
    # verbose=True returns the attention maps (n_layers, N, h, L, L), rows summing to one
    logits, att = model(x.to(DEV), verbose=True)
    assert att.shape == (arch.n_layers, gold["batch"], arch.n_heads, arch.seq_len, arch.seq_len)
    assert torch.allclose(att.sum(-1), torch.ones_like(att.sum(-1)), atol=1e-3)
    assert rel_l2(logits, gold["logits"]) <= 2e-2


def test_block_forward_backward_vs_oracle():
    from vit_plasticity_b200.models import TransformerConfig
    from vit_plasticity_b200.models.layers import TransformerBlock

    arch = O.Arch(emb_dim=256, n_heads=4, n_layers=1, ffn_dim=1024, image_dim=(3, 64, 64))
    sd = O.init_state_dict(arch, seed=3)
    cfg = TransformerConfig(emb_dim=256, n_heads=4, ffn_dim=1024, attn_bias=True, ffn_bias=True, norm="layer", norm_bias=True, norm_eps=1e-12)
    blk = TransformerBlock(cfg).to(DEV)
    blk.load_state_dict({k[len("blocks.0."):]: v for k, v in sd.items() if k.startswith("blocks.0.")})
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 17, 256, generator=g)
    dy = torch.randn(3, 17, 256, generator=g)
    xr = x.clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.block(params, 0, xr, arch)
    ref.backward(dy)
    xg = x.to(DEV).requires_grad_(True)
    out = blk(xg)
    assert out.dtype == torch.float32  # fp32 in -> fp32 out at the module boundary
    out.backward(dy.to(DEV))
    assert rel_l2(out, ref.detach()) <= 1e-2
    assert rel_l2(xg.grad, xr.grad) <= 3e-2
    for k, p in blk.named_parameters():
        assert rel_l2(p.grad, params["blocks.0." + k].grad) <= 3e-2, k


def test_no_cpu_fallback():
    from vit_plasticity_b200 import build_model

    model = build_model({"implementation": "transformer", "image_dim": (3, 32, 32), "patch_type": "computer_vision", "emb_type": "linear",
                         "emb_dim": 128, "n_heads": 2, "n_layers": 1, "attn_bias": True, "ffn_bias": True, "norm_bias": True,
                         "cls_token": True, "output_type": "classification", "n_classes": 3}, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 32, 32))


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_checkpoint_round_trip_like_reference_checkpointer(opt_name, tmp_path):
    """src/vitef/monitor/checkpoint.py:208-237: ``get_state_dict(model, optimizers)`` -> ``dcp.save`` -> ``dcp.load`` ->
    ``set_state_dict`` on the drop-in modules with the fused arena optimizers: parameters, moment buffers and step count
    come back, and the next step of the restored pair equals the next step of the original."""
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint.state_dict import get_state_dict, set_state_dict

    from vit_plasticity_b200.finetune import build_optimizer, train_step

    gold = load("tiny")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    xs = [O.synthetic_images(4, arch, 70 + i).to(DEV) for i in range(3)]
    ys = [O.synthetic_labels(4, arch, 80 + i).to(DEV) for i in range(3)]

    def make():
        m = build("tiny", gold, arch, sd)
        m.train()
        kw = dict(lr=1e-2, momentum=0.9) if opt_name == "sgd" else dict(lr=1e-3, weight_decay=1e-2)
        return m, build_optimizer(m, opt_name, fused=True, **kw)

    model, opt = make()
    for i in range(2):
        train_step(model, opt, [(xs[i], ys[i])], grad_clip=1.0)
    model_sd, optim_sd = get_state_dict(model=model, optimizers=opt)
    dcp.save({"model": model_sd, "optim": optim_sd}, checkpoint_id=str(tmp_path / "ckpt"))

    model2, opt2 = make()
    m2, o2 = get_state_dict(model=model2, optimizers=opt2)
    state = {"model": m2, "optim": o2}
    dcp.load(state, checkpoint_id=str(tmp_path / "ckpt"))
    set_state_dict(model=model2, optimizers=opt2, model_state_dict=state["model"], optim_state_dict=state["optim"])
    assert opt2._steps == opt._steps == 2
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    for name in ("momentum_arena", "exp_avg", "exp_avg_sq"):
        if getattr(opt, name, None) is not None:
            assert torch.equal(getattr(opt, name), getattr(opt2, name)), name
    out = [train_step(m, o, [(xs[2], ys[2])], grad_clip=1.0) for m, o in ((model, opt), (model2, opt2))]
    assert abs(float(out[0][0]) - float(out[1][0])) <= 1e-5
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert rel_l2(b, a) <= 1e-5, (k, rel_l2(b, a))


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_graphed_train_step_matches_eager(opt_name):
    """GraphedTrainStep (whole step captured in a CUDA graph, hyper-parameters read from device memory) against the eager
    train_step on the same model / batches / LR schedule: same kernels in the same order, so after 6 steps (1 eager +
    1 capture + 4 replays, cosine schedule with warm-up, Adam's per-step bias corrections) the parameters agree to fp32
    summation-order noise (the qkv bias gradient is reduced with atomics: not bit-reproducible run to run): rel L2 <= 1e-5."""
    from vit_plasticity_b200.finetune import GraphedTrainStep, build_optimizer, build_scheduler, train_step

    gold = load("small")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    xs = [O.synthetic_images(4, arch, 90 + i).to(DEV) for i in range(6)]
    ys = [O.synthetic_labels(4, arch, 95 + i).to(DEV) for i in range(6)]
    kw = dict(lr=1e-2, momentum=0.9) if opt_name == "sgd" else dict(lr=1e-3, weight_decay=1e-2)

    def make():
        m = build("small", gold, arch, sd)
        m.train()
        o = build_optimizer(m, opt_name, fused=True, **kw)
        return m, o, build_scheduler(o, "cosine", n_steps=6, warmup=2)

    m1, o1, s1 = make()
    m2, o2, s2 = make()
    graphed = GraphedTrainStep(m2, o2, grad_clip=1.0, scheduler=s2)
    for i in range(6):
        l1, g1 = train_step(m1, o1, [(xs[i], ys[i])], grad_clip=1.0, scheduler=s1)
        l2, g2 = graphed([(xs[i], ys[i])])
        assert abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l1)) and abs(float(g1) - float(g2)) <= 1e-5 * float(g1), (i, float(l1), float(l2), float(g1), float(g2))
    assert graphed.graph is not None and graphed.launches_per_step > 20
    assert o1._steps == o2._steps == 6
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert rel_l2(b, a) <= 1e-5, (k, rel_l2(b, a))
    # the bf16 shadows written by the optimizer kernel are what a fresh cast of the parameters gives
    from vit_plasticity_b200 import ops

    for p in o2._shadowed:
        assert torch.equal(ops.shadow_bf16(p), p.detach().reshape(p.shape[0], -1).to(torch.bfloat16))


@pytest.mark.parametrize("name,graphed", [("small", False), ("small", True), ("vit_large", False)])
def test_side_stream_and_fused_bias_gradients_match_plain_path(name, graphed, monkeypatch):
    """The step with weight gradients on the side stream, proj / fc2 bias gradients summed inside the LayerNorm backward and
    the qkv bias gradient split (query third in the attention backward, value third from the proj-dgrad epilogue, key third
    zero) against the same step with all three switched off (one stream, stand-alone column sums, all thirds in the attention
    backward): gradients of one step per tensor to 1e-4 relative L2 (same kernels; the proj / fc2 bias gradients are sums of the same
    bf16 values in another order), the qkv bias gradient to 1e-2 (its value third is a different, mathematically equal,
    expression; its key third is exactly zero instead of round-off), and parameters after 3 steps to 1e-4, eager and replayed from a graph, gradient accumulation over two micro-batches included."""
    from vit_plasticity_b200 import ops
    from vit_plasticity_b200.finetune import GraphedTrainStep, build_optimizer, train_step

    gold = load(name)
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    n = 2 if name == "vit_large" else 4
    xs = [O.synthetic_images(n, arch, 70 + i).to(DEV) for i in range(6)]
    ys = [O.synthetic_labels(n, arch, 80 + i).to(DEV) for i in range(6)]

    def run(new_path: bool):
        monkeypatch.setenv("VB_WGRAD_STREAM", "1" if new_path else "0")
        monkeypatch.setattr(ops, "_FUSED_BIAS", new_path)
        monkeypatch.setattr(ops, "_QBIAS", new_path)
        m = build(name, gold, arch, sd)
        m.train()
        o = build_optimizer(m, "sgd", lr=1e-2, momentum=0.9, fused=True)
        # one backward through the same code path, gradients kept (train_step would consume them)
        with ops.wgrad_overlap(True):
            torch.nn.functional.cross_entropy(m(xs[0]), ys[0]).backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        o.zero_grad()
        step = GraphedTrainStep(m, o, 1.0, grad_acc_steps=2) if graphed else None
        for i in range(3):
            batches = [(xs[2 * i], ys[2 * i]), (xs[2 * i + 1], ys[2 * i + 1])]
            if step is not None:
                step(batches)
            else:
                train_step(m, o, batches, grad_clip=1.0)
        torch.cuda.synchronize()
        return grads, {k: v.detach().clone() for k, v in m.state_dict().items()}

    g_new, p_new = run(True)
    g_old, p_old = run(False)
    # every tensor but the qkv bias comes from the same kernels / the same bf16 values summed in another order
    worst = max((rel_l2(g_new[k], g_old[k]), k) for k in g_old if not k.endswith("qkv_mat.bias"))
    worst_qb = max((rel_l2(g_new[k], g_old[k]), k) for k in g_old if k.endswith("qkv_mat.bias"))
    tag = "/graphed" if graphed else ""
    REPORT[f"{name}/side_stream_fused_bias_worst_grad_rel_l2{tag}"] = worst[0]
    REPORT[f"{name}/qkv_bias_split_worst_grad_rel_l2{tag}"] = worst_qb[0]
    assert worst[0] <= 1e-4, worst
    assert worst_qb[0] <= 1e-2, worst_qb
    e = arch.emb_dim
    k0 = next(k for k in g_new if k.endswith("blocks.0.attn.qkv_mat.bias"))
    assert float(g_new[k0][e:2 * e].abs().max()) == 0.0  # the key third is not computed at all on the new path ...
    assert float(g_old[k0][e:2 * e].abs().max()) <= 1e-2 * float(g_old[k0].abs().max())  # ... and is round-off on the old one
    worst_p = max((rel_l2(p_new[k], p_old[k]), k) for k in p_old)
    REPORT[f"{name}/side_stream_fused_bias_worst_param_rel_l2_after_3_steps{tag}"] = worst_p[0]
    _dump_report()
    assert worst_p[0] <= 1e-4, worst_p  # (measured: <= 1e-5 on the 4-layer model, 3.7e-5 on the 24-layer one)


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_fused_and_torch_optimizer_state_dicts_interchange(opt_name):
    """A checkpoint written with torch.optim.SGD / AdamW resumes under FusedSGD / FusedAdamW and vice versa (same state keys:
    momentum_buffer / exp_avg, exp_avg_sq, step): the step after the hand-over equals the step of the uninterrupted run."""
    from vit_plasticity_b200.finetune import build_optimizer, train_step

    gold = load("tiny")
    arch = arch_of(gold)
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    xs = [O.synthetic_images(4, arch, 70 + i).to(DEV) for i in range(3)]
    ys = [O.synthetic_labels(4, arch, 80 + i).to(DEV) for i in range(3)]
    kw = dict(lr=1e-2, momentum=0.9) if opt_name == "sgd" else dict(lr=1e-3, weight_decay=1e-2)

    def make(fused):
        m = build("tiny", gold, arch, sd)
        m.train()
        return m, build_optimizer(m, opt_name, fused=fused, **kw)

    for first_fused in (True, False):
        ma, oa = make(first_fused)  # runs steps 0, 1 then hands over
        mref, oref = make(first_fused)  # uninterrupted: steps 0, 1, 2
        for i in range(2):
            train_step(ma, oa, [(xs[i], ys[i])], grad_clip=1.0)
            train_step(mref, oref, [(xs[i], ys[i])], grad_clip=1.0)
        mb, ob = make(not first_fused)
        mb.load_state_dict(ma.state_dict())
        ob.load_state_dict(oa.state_dict())
        train_step(mb, ob, [(xs[2], ys[2])], grad_clip=1.0)
        train_step(mref, oref, [(xs[2], ys[2])], grad_clip=1.0)
        for (k, a), (_, b) in zip(mref.state_dict().items(), mb.state_dict().items()):
            assert rel_l2(b, a) <= 2e-5, (first_fused, k, rel_l2(b, a))
