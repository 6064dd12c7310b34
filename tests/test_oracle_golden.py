"""CPU suite: the oracle (oracle/vit_oracle.py) against the committed fixtures, which hold outputs of the REAL
reference (`vitef` imported from /root/reference by oracle/make_golden.py in the build container)."""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O

GOLDEN = Path(__file__).parent / "golden"


def load(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


def checksum(t):
    t = t.double().flatten()
    return [float(t.sum()), float(t.abs().sum()), float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64) % 7).sum())]


def arch_of(gold):
    a = dict(gold["arch"])
    a["image_dim"] = tuple(a["image_dim"])
    return O.Arch(**a)


def weights_of(gold, arch):
    sd = O.init_state_dict(arch, seed=gold["weights_seed"])
    for k, ref in gold["weights_checksum"].items():
        got = checksum(sd[k])
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-6), f"RNG drift: regenerated weight {k} differs from the fixture's"
    return sd


def close(a, b, rtol, atol=1e-6):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max()) <= atol + rtol * float(b.abs().max())


def check_summary(got, summ, rtol, what):
    got = got.detach().float()
    if "full" in summ:
        assert close(got, summ["full"], rtol), what
    else:
        assert abs(float(got.double().norm()) - summ["norm"]) <= rtol * summ["norm"] + 1e-7, what + " (norm)"
        sample = got.flatten()[:: summ["stride"]][:256]
        assert close(sample, summ["sample"], rtol, atol=1e-6 + rtol * summ["norm"] / max(1.0, got.numel() ** 0.5)), what + " (sample)"


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_schema_matches_reference(name):
    gold = load(name)
    arch = arch_of(gold)
    sd = weights_of(gold, arch)
    assert sorted(sd) == gold["state_dict_keys"]
    assert {k: tuple(v.shape) for k, v in sd.items()} == gold["state_dict_shapes"]
    assert sum(v.numel() for v in sd.values()) == gold["n_params"]


def test_vit_base_schema_and_param_counts():
    gold = load("vit_base")
    arch = arch_of(gold)
    assert gold["n_params"] == 85_806_346
    assert len(gold["state_dict_keys"]) == 152 and all(k.startswith("model.") for k in gold["state_dict_keys"])
    assert gold["train"]["attention_only"]["n_trainable"] == 28_357_642
    assert gold["train"]["mlp_only"]["n_trainable"] == 56_678_410
    shapes = O.init_state_dict(O.Arch(**{**gold["arch"], "n_layers": 1}), seed=0)
    for k, v in shapes.items():
        assert tuple(v.shape) == gold["state_dict_shapes"]["model." + k]
    assert arch.seq_len == 197


def test_vit_large_schema_and_param_counts():
    """BASELINE.json configs[3] / [4]: the ViT-L/16 fixture was produced by the reference's own ViT wrapper."""
    gold = load("vit_large")
    arch = arch_of(gold)
    assert (arch.emb_dim, arch.n_heads, arch.n_layers, arch.ffn_dim) == (1024, 16, 24, 4096)  # src/vitef/models/vit.py:130-134
    assert gold["n_params"] == 303_311_882
    assert len(gold["state_dict_keys"]) == 296 and all(k.startswith("model.") for k in gold["state_dict_keys"])
    assert gold["train"]["attention_only"]["n_trainable"] == 100_773_898
    assert gold["train"]["mlp_only"]["n_trainable"] == 201_461_770
    assert len(gold["plasticity"]["keys"]) == 1 + 5 * 24
    assert sorted(gold["plasticity_eps"]) == [1e-3, 1e-2, 1e-1, 1.0, 10.0]
    # the fixture holds whole leading rows ("slab") of every large gradient of the full-finetuning step
    assert "slab" in gold["train"]["full"]["grads"]["blocks.23.ffn.fc1.weight"]


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_forward_and_train_step(name):
    gold = load(name)
    arch = arch_of(gold)
    sd = weights_of(gold, arch)
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"])
    assert np.allclose(checksum(x), gold["x_checksum"], rtol=1e-9)
    y = O.synthetic_labels(gold["batch"], arch, gold["y_seed"])
    assert torch.equal(y, gold["labels"])
    with torch.no_grad():
        assert close(O.forward(sd, x, arch), gold["logits"], 2e-4)
    for fs, ref in gold["train"].items():
        frozen = O.frozen_keys(sd, ref["components"])
        loss, _, grads = O.loss_and_grads(sd, x, y, arch, frozen)
        assert sorted(grads) == ref["trainable"], fs
        assert abs(float(loss) - ref["loss"]) < 2e-5
        for k, g in grads.items():
            check_summary(g, ref["grads"][k], 1e-3, f"{name}/{fs}/{k}")
        sd2, bufs = dict(sd), {}
        gnorm = O.sgd_step(sd2, bufs, grads, 1e-2, 0.9, 1.0)
        assert abs(float(gnorm) - ref["grad_norm"]) < 1e-4 * ref["grad_norm"]
        for k, dn in ref["param_delta_norm"].items():
            assert abs(float((sd2[k] - sd[k]).double().norm()) - dn) <= 1e-3 * dn + 1e-9, k
        for k in frozen:
            assert torch.equal(sd2[k], sd[k])


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_plasticity_and_probes(name):
    gold = load(name)
    arch = arch_of(gold)
    sd = weights_of(gold, arch)
    p = gold["plasticity"]
    x1 = O.synthetic_images(p["n_pairs"], arch, p["x1_seed"])
    x2 = O.synthetic_images(p["n_pairs"], arch, p["x2_seed"])
    dist = O.pair_distances(sd, x1, x2, arch)
    assert list(dist) == p["keys"]
    assert len(dist) == 1 + 5 * arch.n_layers
    for k in dist:
        assert close(dist[k], p["distances"][k], 3e-4), k
    plast = O.plasticity(dist)
    assert set(plast) == {"attn_norm", "attn", "ffn_norm", "ffn_fc1", "ffn_fc2"}
    for comp, per_layer in plast.items():
        for i, r in enumerate(per_layer):
            assert close(r, p["ratios"][f"block{i}_{comp}"], 5e-4), (comp, i)
    noise = O.synthetic_images(p["n_pairs"], arch, gold["plasticity_eps_noise_seed"])
    for eps, ratios in gold["plasticity_eps"].items():
        d = O.pair_distances(sd, x1, x1 + eps * noise, arch)
        for k, r in ratios.items():
            # fp32 cancellation grows as eps shrinks: tolerance scales with 1/eps
            assert close(d[k] / d["embedding"], r, 5e-4 / min(1.0, eps * 10)), (eps, k)
    x = O.synthetic_images(gold["batch"], arch, gold["x_seed"])
    with torch.no_grad():
        pr = O.probes(sd, x, arch)
    assert list(pr) == list(gold["probes"]) and len(pr) == 8 * arch.n_layers
    for k, v in pr.items():
        assert close(v[:, 0, :], gold["probes"][k]["cls"], 5e-4, atol=1e-5)
        assert close(v.mean(1), gold["probes"][k]["mean"], 5e-4, atol=1e-5)


def test_edge_cases_distance_and_freeze():
    a = torch.zeros(3, 5, 4)
    assert torch.equal(O.distance(a, a), torch.zeros(3))
    b = a.clone()
    b[1, 2, 3] = 2.0
    assert torch.allclose(O.distance(a, b), torch.tensor([0.0, 2.0, 0.0]))
    arch = O.Arch(emb_dim=128, n_heads=2, n_layers=2, ffn_dim=512, image_dim=(3, 32, 32))
    sd = O.init_state_dict(arch, 0)
    assert O.frozen_keys(sd, []) == set()
    fr = O.frozen_keys(sd, ["emb", "attn_norm", "mha", "ffn_norm", "ffn_fc1", "ffn_fc2"])
    assert set(sd) - fr == {k for k in sd if k.startswith("output.")}  # final norm + head are never frozen
    with pytest.raises(KeyError):
        O.frozen_keys(sd, ["ffn_activation"])  # listed in the reference docstring but absent from its map
