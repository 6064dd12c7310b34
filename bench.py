"""Headline benchmark: ViT-B/16 bf16 finetuning images/s on B200 (BASELINE.json configs[1]); the plasticity estimator's
pairs/s (configs[0] shapes) and a bounded ViT-L/16 perturbation sweep (configs[4] shapes) ride along as secondary blocks
of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model base|large] [--components ...] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --workload sweep [--images 65536] [--eps 1e-3,1e-2,1e-1,1,10]   # configs[4] in full (ViT-L/16)
    python bench.py --impl reference     # the UNMODIFIED reference (baseline/_ref) on the host cores

Finetuning: a step = forward + cross-entropy + backward + grad-norm clip + SGD-momentum step on one synthetic batch
(apps/vit/train.py:263-283 semantics). N = 1 runs batch 512 (configs[1]). N > 1 runs the split BASELINE.json configs[2]/[3]
and SURVEY.md 8(e) state: GLOBAL batch 512, 512 / N images per rank ("scaling": "strong"); `--scaling weak` keeps
`--batch` images per rank instead. ``value`` times K steps with the batch already resident in HBM; ``e2e`` times the
same K steps through the public API with the batch in pinned HOST memory (H2D copy of the images and labels and a D2H
read of the loss inside the timed region, every step). Timing: CUDA events around the K steps, barrier + synchronize on
both sides, max over ranks. L2 (126 MB) cannot carry anything between steps: one step streams > 30 GB of activations at
batch 512, > 3.8 GB at batch 64 (config.l2 = "inputs>L2").
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FWD_GFLOP_PER_IMG = {"base": 35.126, "large": 123.107}  # SURVEY.md Appendix D (2 FLOPs per MAC)
VIT = {"base": dict(e=768, h=12, nl=12, f=3072), "large": dict(e=1024, h=16, nl=24, f=4096)}
SEQ, NP, KP = 197, 196, 3 * 16 * 16
EPS_GRID = "1e-3,1e-2,1e-1,1,10"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="finetune", choices=["finetune", "sweep"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="N > 1: strong = global batch split over the ranks (default), weak = --batch per rank")
    ap.add_argument("--global-batch", type=int, default=512, help="global batch of the strong-scaling split (BASELINE.json configs[2]/[3])")
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch at N = 1 and under --scaling weak")
    ap.add_argument("--model", default=None, help="base | large (default: base for finetune, large for sweep)")
    ap.add_argument("--components", default="", help="comma-separated components to FREEZE (apps/vit/utils.py:67-74)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="plasticity pairs per call (0 disables the secondary block)")
    ap.add_argument("--images", type=int, default=65536, help="--workload sweep: images in total (sharded over the ranks)")
    ap.add_argument("--eps", default=EPS_GRID, help="perturbation magnitudes of the sweep")
    ap.add_argument("--sweep-images", type=int, default=256, help="images per rank of the secondary ViT-L sweep block (0 disables it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the informational block: reference modules eager on the same GPU")
    ap.add_argument("--torch-sgd", action="store_true", help="clip_grad_norm_ + torch.optim.SGD instead of the fused arena step")
    ap.add_argument("--skip-eager-roofline", action="store_true", help="do not re-run the steps eagerly with per-launch events (profiling runs under ncu)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel of the step from Python instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.model is None:
        args.model = "large" if args.workload == "sweep" else "base"
    if args.scaling is None:
        args.scaling = "strong"
    return args


def per_rank_batch(args, world: int) -> int:
    if world == 1 or args.scaling == "weak":
        return args.batch
    if args.global_batch % world:
        raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
    return args.global_batch // world


def finetune_config(model_name: str, batch: int, world: int, comps, n_trainable, scaling: str):
    """The `config` object of the finetuning line, identical for both arms (the reference arm times a bounded sample of it)."""
    return {"workload": f"ViT-{model_name}/16 finetuning (fwd + CE + bwd + clip 1.0 + SGD 1e-2 m0.9), batch {batch}/GPU, 10-class synthetic CIFAR-10-shaped 224x224, random init",
            "global_batch": batch * world, "parallelism": f"dp{world}", "freeze": comps, "trainable_params": n_trainable,
            "scaling": scaling if world > 1 else "single GPU", "l2": "inputs>L2 (one step streams GBs of activations)"}


def sweep_config(model_name: str, images: int, eps, world: int, pairs_per_call: int):
    return {"workload": f"ViT-{model_name}/16 plasticity sweep: {images} images x {len(eps)} perturbation magnitudes {eps} = {images * len(eps)} pairs (x, x + eps n), "
                        f"n ~ N(0,1) drawn on the device, 224x224 synthetic CIFAR-10-shaped inputs, random init",
            "images": images, "eps": eps, "parallelism": f"pairs sharded contiguously over {world} rank(s), no data-path collective, one gather at the end",
            "pairs_per_call": pairs_per_call, "l2": "inputs>L2 (one call streams > 4 GB of projections)"}


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, averaged over the launches of the committed
    `ncu --set full` captures of this command (the latest round's profiles/*ncu_full_gemm*.txt: forward launches and the
    backward launches of one block); None if no capture is committed."""
    files = sorted((ROOT / "profiles").glob("*ncu_full_gemm*.txt"))
    if not files:
        return None, None
    prefix = files[-1].name.split("_ncu_full_gemm")[0]  # e.g. "r02_z"
    files = [f for f in files if f.name.startswith(prefix + "_ncu_full_gemm")]
    tot, n = 0.0, 0
    for f in files:
        for line in f.read_text().splitlines():
            parts = line.split()
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val, unit = float(parts[1]), parts[2]
                tot += val * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
                n += 1
    return (tot / (n / 2) if n else None), " + ".join(f.name for f in files)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            p.update(json.loads(f.read_text()))
            p["source"] = "measured"
        except Exception:
            pass
    return p


def sweep_gflop(model_name: str, n_eps: int) -> dict:
    """FLOPs the fused estimator EXECUTES per image of a sweep over n_eps magnitudes (2 per MAC), and what the reference's
    procedure (two full get_decomposition passes per pair, fc2 on the zero-padded 4E input) would spend on the same pairs."""
    v = VIT[model_name]
    e, h, nl, f = v["e"], v["h"], v["nl"], v["f"]
    patch = 2.0 * NP * e * KP
    qkv = 2.0 * SEQ * e * 3 * e * nl
    fc1 = 2.0 * SEQ * e * f * nl
    fc2 = 2.0 * SEQ * e * e * nl
    proj = 2.0 * SEQ * e * e * nl
    core = 2.0 * SEQ * SEQ * 64 * h * nl  # one score-shaped or P V-shaped contraction over all heads and layers
    shared = 2 * patch + 2 * qkv + fc1 + fc2          # base + direction embeddings and projections, fc1 / fc2 on the direction
    per_eps = 8 * core + proj                          # S, 3 x dS, 4 x P V; output projection of the difference
    faithful_pair = 2 * (patch + qkv + 2 * core + proj + fc1 + 2.0 * SEQ * f * e * nl)  # reference-faithful count per pair
    return {"executed_per_image": (shared + n_eps * per_eps) / 1e9, "executed_per_pair": (shared / n_eps + per_eps) / 1e9,
            "reference_faithful_per_pair": faithful_pair / 1e9}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """The timed region starts now (nvidia-smi itself is started earlier: it can take over a second to come up)."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        timed_rows = [r for t, r in self.rows if t >= self.t_mark]
        window = "timed region"
        if not timed_rows:  # sampler slower than the timed region: fall back to the warm-up steps (same load)
            timed_rows, window = [r for _, r in self.rows], "warm-up + timed region"
        self.rows = timed_rows
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        # the upper half of the samples is "under load" (idle gaps between host phases pull the plain median down)
        load = sm[len(sm) // 2 :] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "window": window}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (baseline/_ref) on the host cores
# --------------------------------------------------------------------------------------------------
CPU_SAMPLE_BATCH = 8  # images per step of the CPU arm: a bounded sample of the batch-512 workload (CPU minutes otherwise)


def cpu_port_baseline(model_name: str, comps, steps: int, warmup: int, why: str):
    """Fallback CPU arm when the reference install is absent: the oracle's restatement of the same step (oracle/vit_oracle.py,
    torch fp32, all host threads) — the one other place bench.py may execute oracle/."""
    import statistics

    import torch

    from oracle import vit_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = O.vit_arch(model_name, n_classes=10)
    sd = O.init_state_dict(arch, seed=42)
    frozen = O.frozen_keys(sd, comps)
    x, y = O.synthetic_images(CPU_SAMPLE_BATCH, arch, 1), O.synthetic_labels(CPU_SAMPLE_BATCH, arch, 2)
    bufs, times = {}, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = O.loss_and_grads(sd, x, y, arch, frozen)
        O.sgd_step(sd, bufs, grads, 1e-2, 0.9, 1.0)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    n_trainable = sum(v.numel() for k, v in sd.items() if k not in frozen)
    return {"value": round(CPU_SAMPLE_BATCH / dt, 3), "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"oracle port (torch fp32, {cores} threads) of the same step at batch {CPU_SAMPLE_BATCH} (of 512), median of {steps} steps after {warmup} warm-up, {dt:.2f} s/step; "
                      f"the unmodified reference was not available here ({why})"}, n_trainable


def cpu_reference_baseline(model_name: str, comps, steps: int, warmup: int, with_omp1: bool, plasticity_pairs: int):
    """cpu_baseline object: the reference's own step on all host cores (`value`), the reference-faithful OMP_NUM_THREADS=1
    setting the apps force at import (train.py:16, analysis.py:14) beside it, and configs[0] in full for the estimator."""
    import torch

    from baseline import reference_arm as R

    why = R.available()
    if why is not None:  # baseline/_ref did not travel: time the oracle port instead, and say so (kind "port")
        return cpu_port_baseline(model_name, comps, steps, warmup, why)
    cores = os.cpu_count() or 1
    v, dt, n_trainable = R.finetune(model_name, CPU_SAMPLE_BATCH, steps, warmup, comps, cores)
    out = {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "reference",
           "sample": f"unmodified reference (baseline/_ref: vitef.models.build_model ViT-{model_name}/16 -> F.cross_entropy -> backward -> clip_grad_norm_ -> "
                     f"vitef.optim SGD) fp32 on {cores} host threads, batch {CPU_SAMPLE_BATCH} per step (of the batch-512 workload), median of {steps} steps after {warmup} warm-up, {dt:.2f} s/step"}
    if with_omp1:
        v1, dt1, _ = R.finetune(model_name, CPU_SAMPLE_BATCH, max(2, min(steps, 3)), 1, comps, 1)
        out["omp_num_threads_1"] = {"value": round(v1, 3), "unit": "img/s", "cores": 1, "s_per_step": round(dt1, 2),
                                    "note": "reference-faithful threading: the apps set OMP_NUM_THREADS=1 at import (apps/vit/train.py:16)"}
    if plasticity_pairs > 0:
        pv, pdt = R.plasticity("base", plasticity_pairs, 16, cores)
        out["plasticity"] = {"value": round(pv, 3), "unit": "pairs/s", "cores": cores, "kind": "reference", "pairs": plasticity_pairs, "seconds": round(pdt, 2),
                             "sample": f"BASELINE.json configs[0] in full: ViT-B/16, {plasticity_pairs} synthetic pairs, get_decomposition x 2 + apps.vit.analysis.distance per key, fp32 CPU, {cores} threads"}
    torch.set_num_threads(cores)
    return out, n_trainable


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    comps = [c for c in args.components.split(",") if c]
    world = max(1, args.gpus)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if args.workload == "sweep":
        eps = [float(e) for e in args.eps.split(",")]
        return run_reference_sweep(args, eps, world)
    cpu, n_trainable = cpu_reference_baseline(args.model, comps, steps, warmup, with_omp1=False, plasticity_pairs=0)
    v = cpu["value"]
    print(json.dumps({
        "impl": "reference", "metric": f"ViT-{args.model[0].upper()}/16 finetune img/s", "value": v, "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(CPU_SAMPLE_BATCH / v * 1e3, 2), "higher_is_better": True,
        "scaling": "strong" if (world > 1 and args.scaling == "strong") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": finetune_config(args.model, per_rank_batch(args, world), world, comps, n_trainable, args.scaling),
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference_sweep(args, eps, world):
    """Reference procedure for the sweep on the host cores: per image and magnitude, get_decomposition(x) and
    get_decomposition(x + eps n) + distance per key (f(x) recomputed per pair, exactly as apps/vit/analysis.py would)."""
    import torch

    from baseline import reference_arm as R

    why = R.available()
    if why is not None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    cores = os.cpu_count() or 1
    pairs = max(2, min(8, args.steps))  # bounded sample: a handful of ViT-L pairs (~4 s each on 16 threads)
    v, dt = R.plasticity(args.model, pairs, 2, cores)
    print(json.dumps({
        "impl": "reference", "metric": f"ViT-{args.model[0].upper()}/16 plasticity sweep pairs/s", "value": round(v, 4), "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / v, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": sweep_config(args.model, args.images, eps, world, 64),
        "cpu_baseline": {"value": round(v, 4), "unit": "pairs/s", "cores": cores, "kind": "reference",
                         "sample": f"unmodified reference estimator (ViT.get_decomposition x 2 + apps.vit.analysis.distance) on {pairs} ViT-{args.model}/16 pairs, fp32, {cores} threads, {dt:.1f} s"},
        "e2e": {"value": round(v, 4), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    torch.set_num_threads(cores)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the B200 arm: device, ranks, timing helper."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs CUDA devices: the product path has no CPU fallback (use --impl reference for the CPU arm)")
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        if self.world > 1:
            dist.init_process_group(backend="nccl", device_id=self.dev)
        self.pk = peaks()

    def timed(self, fn_step, steps):
        """CUDA events around `steps` calls, barrier + synchronize on both sides, MAX over ranks (ms)."""
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn_step(i)
        e.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_vit(ctx, model_name: str):
    from vit_plasticity_b200 import build_model

    ctx.torch.manual_seed(42)
    return build_model({"implementation": "vit", "model_name": model_name, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device=ctx.dev)


def dp_parity_check(ctx, model, dp, opt, global_batch: int):
    """One optimisation step on IDENTICAL data two ways, from identical weights: (a) through DataParallel, every rank on its
    contiguous shard of the batch, gradients averaged by the bucketed all-reduce; (b) every rank alone on the whole batch
    (no collective). Compares the pre-clip gradient norm and a checksum of the updated parameters between (a) and (b) and
    across ranks. Weights and optimizer state are restored afterwards."""
    torch, dist = ctx.torch, ctx.dist
    from vit_plasticity_b200.finetune import train_step

    n = min(global_batch, 512)
    n -= n % ctx.world
    g = torch.Generator().manual_seed(99)
    x, y = torch.randn(n, 3, 224, 224, generator=g).to(ctx.dev), torch.randint(0, 10, (n,), generator=g).to(ctx.dev)
    params = [p for p in model.parameters()]
    saved = [p.detach().clone() for p in params]

    def restore():
        with torch.no_grad():
            for p, s in zip(params, saved):
                p.copy_(s)
        if getattr(opt, "momentum_arena", None) is not None:
            opt.momentum_arena.zero_()
        opt._steps = 0
        opt.zero_grad()

    def checksum():
        return torch.stack([p.detach().double().sum() for p in params if p.requires_grad]).sum(), torch.stack([p.detach().double().abs().sum() for p in params if p.requires_grad]).sum()

    per = n // ctx.world
    lo = ctx.rank * per
    _, gn_dp = train_step(dp, opt, [(x[lo : lo + per], y[lo : lo + per])], grad_clip=1.0, after_backward=dp.finish_grad_sync)
    cs_dp = checksum()
    gn_dp = gn_dp.double().clone()
    delta_dp = torch.sqrt(sum(((p.detach() - s).double() ** 2).sum() for p, s in zip(params, saved) if p.requires_grad))
    restore()
    dp.require_grad_sync = False  # (b): local gradients only, no bucket is launched
    _, gn_one = train_step(model, opt, [(x, y)], grad_clip=1.0)
    dp.require_grad_sync = True
    cs_one = checksum()
    upd_diff = torch.sqrt(sum(((p.detach() - s).double() ** 2).sum() for p, s in zip(params, saved) if p.requires_grad))
    gn_one = gn_one.double().clone()
    restore()
    mine = torch.stack([gn_dp, cs_dp[0], cs_dp[1]]).to(ctx.dev)
    allr = [torch.empty_like(mine) for _ in range(ctx.world)]
    dist.all_gather(allr, mine)
    allr = torch.stack(allr)
    replicas_identical = bool((allr == allr[0]).all())
    rel = lambda a, b: abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)
    out = {"global_batch": n, "ranks": ctx.world, "grad_norm_dp": float(gn_dp), "grad_norm_single_process": float(gn_one),
           "grad_norm_rel_diff": rel(gn_dp, gn_one), "param_abs_checksum_rel_diff": rel(cs_dp[1], cs_one[1]),
           "update_norm_rel_diff": rel(delta_dp, upd_diff), "replicas_bit_identical": replicas_identical}
    out["pass"] = bool(replicas_identical and out["grad_norm_rel_diff"] <= 1e-3 and out["update_norm_rel_diff"] <= 1e-3)
    return out


def bench_finetune(ctx, args):
    torch, dist = ctx.torch, ctx.dist
    from vit_plasticity_b200 import _lib
    from vit_plasticity_b200.distributed import DataParallel
    from vit_plasticity_b200.finetune import GraphedTrainStep, build_optimizer, freeze_model, train_step
    from vit_plasticity_b200.plasticity import PlasticityEstimator

    rank, world, dev, pk = ctx.rank, ctx.world, ctx.dev, ctx.pk
    comps = [c for c in args.components.split(",") if c]
    model = build_vit(ctx, args.model)
    model.train()
    freeze_model(model, comps)
    n_trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    dp = DataParallel(model) if world > 1 else None
    # apps/vit/configs/cifar10.yaml: SGD lr 1e-2 momentum 0.9, grad_clip 1 — as the fused clip+SGD step over the flat
    # gradient arena (same arithmetic as clip_grad_norm_ + torch.optim.SGD; tests/test_model_gpu.py::test_fused_sgd_*)
    opt = build_optimizer(dp or model, "sgd", lr=1e-2, momentum=0.9, fused=not args.torch_sgd)
    B = per_rank_batch(args, world)
    parity = dp_parity_check(ctx, model, dp, opt, B * world) if (dp is not None and not args.torch_sgd) else None
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2  # distinct pinned host batches, alternated
    host = [(torch.randn(B, 3, 224, 224, generator=g).pin_memory(), torch.randint(0, 10, (B,), generator=g).pin_memory()) for _ in range(n_host)]
    devb = [(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)) for x, y in host]
    after = dp.finish_grad_sync if dp is not None else None
    use_graph = not (args.no_graph or args.torch_sgd)
    # the public API for fixed-shape training: the whole step (forward, loss, backward, bucket all-reduces, clip + SGD,
    # arena reset) captured once and replayed — one graph launch per step instead of ~300 launches from Python
    graphed = GraphedTrainStep(dp or model, opt, grad_clip=1.0, after_backward=after) if use_graph else None

    def step_eager(i):
        return train_step(dp or model, opt, [devb[i % n_host]], grad_clip=1.0, after_backward=after)

    def step_resident(i):
        if graphed is not None:
            return graphed([devb[i % n_host]])
        return step_eager(i)

    copy_stream = torch.cuda.Stream()
    staged = {}

    def prefetch(i):
        # next batch's H2D copy on a side stream, overlapped with the current step's compute
        with torch.cuda.stream(copy_stream):
            x, y = host[i % n_host]
            staged[i] = (x.to(dev, non_blocking=True), y.to(dev, non_blocking=True), torch.cuda.Event())
            staged[i][2].record(copy_stream)

    def step_e2e(i, last):
        if i not in staged:
            prefetch(i)
        x, y, ev = staged.pop(i)
        torch.cuda.current_stream().wait_event(ev)
        x.record_stream(torch.cuda.current_stream())
        y.record_stream(torch.cuda.current_stream())
        if not last:
            prefetch(i + 1)
        if graphed is not None:
            loss, _ = graphed([(x, y)])
        else:
            loss, _ = train_step(dp or model, opt, [(x, y)], grad_clip=1.0, after_backward=after)
        return float(loss)  # D2H read of the step's result

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    # ---- warm-up (untimed): the first call is eager, the second captures the graph ----
    warm = max(3, args.warmup)
    for i in range(warm):
        step_resident(i)
    torch.cuda.synchronize()

    # ---- timed: inputs resident in HBM ----
    if rank == 0:
        sampler.mark()
    _lib.reset_launch_count()
    ms_best = ctx.timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = graphed.launches_per_step * args.steps if graphed is not None else _lib.launch_count()
    graph_info = {"used": graphed is not None, "launches_per_step": graphed.launches_per_step if graphed is not None else None}
    ms_step = ms_best / args.steps
    value = world * B * args.steps / (ms_best / 1e3)
    # ---- the same steps issued kernel by kernel, every GEMM launch bracketed by CUDA events on its stream (roofline) ----
    _lib.GEMM_EVENTS = []
    ms_total = ms_best
    if not args.skip_eager_roofline:
        for i in range(2):
            step_eager(i)
        _lib.GEMM_EVENTS = []
        ms_total = ctx.timed(step_eager, args.steps)
    gemm_events, _lib.GEMM_EVENTS = _lib.GEMM_EVENTS, None
    gemm_ms = sum(s.elapsed_time(e) for s, e, _ in gemm_events)
    gemm_flops = sum(f for _, _, f in gemm_events)

    # ---- timed: end to end from pinned host memory ----
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            step_e2e(i, last=(i == 1))
        staged.clear()
        ms_e2e = ctx.timed(lambda i: step_e2e(i, last=(i == args.steps - 1)), args.steps)
        h2d = B * 3 * 224 * 224 * 4 + B * 8
        e2e = {"value": round(world * B * args.steps / (ms_e2e / 1e3), 2), "unit": "img/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
               "ms_per_step": round(ms_e2e / args.steps, 3)}

    # ---- secondary: the same step fed by the device-side input pipeline: the batch crosses PCIe as
    # uint8 32x32x3 CIFAR-shaped images; RandomResizedCrop / flip / ToTensor / Normalize (the reference's "train"
    # transform, bit-exact against PIL / torchvision) run on the GPU and emit the patch-embed GEMM's bf16 operand ----
    pipe = None
    if not args.no_e2e:
        from vit_plasticity_b200.preprocess import DevicePreprocessor

        pre = DevicePreprocessor(224, "train", dev)
        hu8 = [torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_host)]
        ustaged = {}
        graphed_u8 = GraphedTrainStep(dp or model, opt, grad_clip=1.0, after_backward=after) if use_graph else None

        def u_prep(i):
            # host side of the next batch (crop boxes / flips in torchvision's draw order, 1.5 MB uint8 copy) while the
            # GPU is busy with the current step
            ustaged[i] = (hu8[i % n_host].to(dev, non_blocking=True), pre.make_params(B, 32, 32))

        def step_u8(i, last=False):
            if i not in ustaged:
                u_prep(i)
            xu8, params = ustaged.pop(i)
            x = pre.patches(xu8, params)
            if graphed_u8 is not None:
                loss, _ = graphed_u8([(x, devb[i % n_host][1])])
            else:
                loss, _ = train_step(dp or model, opt, [(x, devb[i % n_host][1])], grad_clip=1.0, after_backward=after)
            if not last:
                u_prep(i + 1)
            return float(loss)

        for i in range(3):
            step_u8(i, last=(i == 2))
        ustaged.clear()
        ms_u8 = ctx.timed(lambda i: step_u8(i, last=(i == args.steps - 1)), args.steps)
        pipe = {"value": round(world * B * args.steps / (ms_u8 / 1e3), 2), "unit": "img/s", "h2d_bytes_per_step": (B * 32 * 32 * 3 + B * 32) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": round(ms_u8 / args.steps, 3),
                "source": "uint8 32x32x3 batch in pinned host memory; crop boxes / flips drawn on the host in torchvision's RNG order"}

    # ---- secondary: plasticity estimator pairs/s, configs[0] shapes (each rank takes its own pairs; no collective) ----
    plast = None
    if args.pairs > 0:
        model.eval()
        est = PlasticityEstimator(model)
        P = args.pairs
        hx1, hx2 = torch.randn(P, 3, 224, 224, generator=g).pin_memory(), torch.randn(P, 3, 224, 224, generator=g).pin_memory()
        dx1, dx2 = hx1.to(dev), hx2.to(dev)
        for _ in range(2):
            est.squared_distances(dx1, dx2)
        reps = 5
        ms_p = ctx.timed(lambda i: est.squared_distances(dx1, dx2), reps)
        # end to end: the pair images start in pinned host memory; call i + 1's H2D copy runs on the side stream into the
        # other of two fixed staging buffers while call i computes; the distance table is read back to the host every call
        stage = [(torch.empty_like(dx1), torch.empty_like(dx2)) for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        pstaged = set()

        def p_prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])  # the call that last read this staging buffer has finished
                stage[s][0].copy_(hx1, non_blocking=True)
                stage[s][1].copy_(hx2, non_blocking=True)
                copied[s].record(copy_stream)
            pstaged.add(i)

        def p_step(i, last):
            if i not in pstaged:
                p_prefetch(i)
            pstaged.discard(i)
            s = i % 2
            torch.cuda.current_stream().wait_event(copied[s])
            if not last:
                p_prefetch(i + 1)
            out = est.pair_distances(stage[s][0], stage[s][1])  # numpy on the host: D2H of the (1 + 5 n_layers) x P table
            consumed[s].record()
            return out

        p_step(0, True)
        ms_pe = ctx.timed(lambda i: p_step(i, i == reps - 1), reps)
        pps = world * P * reps / (ms_p / 1e3)
        gf = sweep_gflop(args.model, 1)
        plast = {"metric": f"ViT-{args.model[0].upper()}/16 plasticity pairs/s", "value": round(pps, 1), "unit": "pairs/s", "pairs_per_call": P,
                 "e2e": {"value": round(world * P * reps / (ms_pe / 1e3), 1), "unit": "pairs/s", "h2d_bytes_per_step": 2 * P * 3 * 224 * 224 * 4 * world,
                         "d2h_bytes_per_step": (1 + 5 * len(model.model.blocks)) * P * 4 * world},
                 "roofline": {"bound": "tensor", "achieved": round(pps / world * gf["executed_per_pair"] / 1e3, 1), "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": round(pps / world * gf["executed_per_pair"] / 1e3 / pk["bf16_tflops_sustained"], 4),
                              "gflop_per_pair_executed": round(gf["executed_per_pair"], 3), "gflop_per_pair_reference_faithful": round(gf["reference_faithful_per_pair"], 3)}}
        del est, dx1, dx2, stage
        model.train()

    scale_parity = None
    if world == 1 and not comps and not args.no_gpu_eager and not args.torch_sgd:
        graphed = graphed_u8 = None  # release the graphs' private pools before the fp32 reference allocates its activations
        torch.cuda.empty_cache()
        scale_parity = scale_parity_block(ctx, args, model, opt, devb[0])
    # free the finetuning state before the ViT-L secondary block
    del devb, host, staged
    graphed = graphed_u8 = None
    opt = dp = None
    torch.cuda.empty_cache()
    sweep = None
    if args.sweep_images > 0:
        try:
            sweep = bench_sweep(ctx, args, model_name="large", images_total=args.sweep_images * world, secondary=True)
        except Exception as exc:  # a secondary block must not take the headline down
            sweep = {"error": repr(exc)[:300]}

    if rank != 0:
        return

    # ---- CPU baseline: the unmodified reference on the host cores (bounded sample), N = 1 only ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu, _ = cpu_reference_baseline(args.model, comps, steps=5, warmup=1, with_omp1=True, plasticity_pairs=64 if args.pairs > 0 else 0)
    eager = None
    if not args.no_gpu_eager and world == 1:
        eager = gpu_eager_block(ctx, args, comps)

    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_src = ncu_traffic_per_launch()
    step_flops = 3 * FWD_GFLOP_PER_IMG.get(args.model, 0) * 1e9 * B if not comps else None
    out = {
        "metric": f"ViT-{args.model[0].upper()}/16 finetune img/s", "value": round(value, 2), "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong" if (world > 1 and args.scaling == "strong") else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": finetune_config(args.model, B, world, comps, n_trainable, args.scaling),
        "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1) if achieved else None, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": round(achieved / pk["bf16_tflops_sustained"], 4) if achieved else None,
                     "traffic": round(traffic) if traffic else None, "traffic_unit": "bytes per launch (mean over the launches of " + str(traffic_src) + ")" if traffic else None,
                     "flops_per_launch": round(gemm_flops / max(1, len(gemm_events))), "us_per_launch": round(gemm_ms * 1e3 / max(1, len(gemm_events)), 2),
                     "kernel": "gemm_tcgen05_kernel (all fwd/dgrad/wgrad launches of the timed steps, CUDA events per launch)", "peak_source": pk["source"] + " bf16_tflops_sustained",
                     "gemm_share_of_step": round(gemm_ms / ms_total, 4), "gemm_launches": len(gemm_events),
                     "whole_step_frac": round(step_flops * args.steps / (ms_best / 1e3) / 1e12 / pk["bf16_tflops_sustained"], 4) if step_flops else None,
                     "note": "per-launch durations from the eager run of the same steps (events cannot bracket nodes of a replayed graph); gemm_share_of_step is relative to that run"},
        "cpu_baseline": cpu, "clocks": clocks, "dp_parity": parity, "scale_parity": scale_parity, "plasticity": plast, "sweep": sweep, "e2e_u8_input_pipeline": pipe, "gpu_eager": eager,
        "cuda_graph": graph_info, "ms_per_step_eager_with_per_launch_events": round(ms_total / args.steps, 3),
    }
    print(json.dumps(out))


def scale_parity_block(ctx, args, model, opt, batch):
    """Parity at BENCH scale (N = 1): loss and every parameter gradient of one forward + backward on the bench batch (512
    images: M = 100 864 token rows, weight gradients reduced over 100 864 tokens by split-K + TMA reduce-add) against the
    UNMODIFIED reference run in fp32 (TF32 off) on this same GPU with the same weights. Stated tolerances, as in
    tests/test_model_gpu.py: loss abs 2e-2, gradient norm 2e-2 relative, per-tensor relative L2 4e-2."""
    torch = ctx.torch
    import torch.nn.functional as F

    from baseline import reference_arm as R

    if R.available() is not None:
        return {"unavailable": R.available()}
    x, y = batch
    try:
        opt.zero_grad()
        model.train()
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        mine = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        opt.zero_grad()
        torch.cuda.empty_cache()
        ref_loss, ref = R.loss_and_grads_on(str(ctx.dev), args.model, {k: v.detach().clone() for k, v in model.state_dict().items()}, x, y)
        errs = {k: float((mine[k].double() - g.double()).norm() / g.double().norm().clamp_min(1e-30)) for k, g in ref.items() if k in mine}
        gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in mine.values())))
        gn_ref = float(torch.sqrt(sum((g.double() ** 2).sum() for g in ref.values())))
        worst = max(errs, key=errs.get)
        out = {"batch": int(x.shape[0]), "token_rows": int(x.shape[0]) * SEQ, "loss": float(loss), "loss_reference_fp32": ref_loss, "loss_abs_diff": abs(float(loss) - ref_loss),
               "grad_norm": gn, "grad_norm_reference_fp32": gn_ref, "grad_norm_rel_diff": abs(gn - gn_ref) / gn_ref, "tensors_compared": len(errs),
               "worst_grad_rel_l2": errs[worst], "worst_tensor": worst, "median_grad_rel_l2": sorted(errs.values())[len(errs) // 2],
               "tolerances": {"loss_abs": 2e-2, "grad_norm_rel": 2e-2, "grad_rel_l2": 4e-2}}
        out["pass"] = bool(out["loss_abs_diff"] <= 2e-2 and out["grad_norm_rel_diff"] <= 2e-2 and out["worst_grad_rel_l2"] <= 4e-2 and set(ref) == set(mine))
        del ref, mine
        torch.cuda.empty_cache()
        return out
    except Exception as exc:  # informational block: never take the headline down (e.g. out of memory for the fp32 reference)
        return {"error": repr(exc)[:300]}


def gpu_eager_block(ctx, args, comps):
    """Informational (SURVEY.md 8d 'GPU baseline'): the reference's own modules, eager, on this same GPU at the same batch —
    fp32 as shipped (TF32 off) and under torch.autocast(bfloat16). Not the reference arm (that one is the CPU path)."""
    torch = ctx.torch
    from baseline import reference_arm as R

    if R.available() is not None:
        return {"unavailable": R.available()}
    out = {}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name, autocast in (("fp32", False), ("autocast_bf16", True)):
        try:
            torch.cuda.empty_cache()
            v, dt, _ = R.finetune(args.model, args.batch, 3, 1, comps, os.cpu_count() or 1, device=str(ctx.dev), autocast=autocast)
            out[name] = {"value": round(v, 1), "unit": "img/s", "ms_per_step": round(dt * 1e3, 2), "batch": args.batch}
        except Exception as exc:  # e.g. out of memory: the reference materialises two fp32 score tensors per layer
            out[name] = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()
    try:  # the reference estimator as shipped, on this GPU: every component output crosses PCIe twice (analysis.py:216-226)
        R.plasticity(args.model, 16, 16, os.cpu_count() or 1, device=str(ctx.dev))
        v, dt = R.plasticity(args.model, 64, 16, os.cpu_count() or 1, device=str(ctx.dev))
        out["estimator_as_shipped"] = {"value": round(v, 1), "unit": "pairs/s", "pairs": 64, "seconds": round(dt, 2),
                                       "note": "get_decomposition x 2 (outputs to the host) + re-upload + distance per key, fp32, batches of 16"}
    except Exception as exc:
        out["estimator_as_shipped"] = {"error": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out


def bench_sweep(ctx, args, model_name: str, images_total: int, secondary: bool = False):
    """BASELINE.json configs[4]: ViT-L/16, `images_total` images x the eps grid, sharded over the ranks."""
    torch, dist = ctx.torch, ctx.dist
    from vit_plasticity_b200 import _lib
    from vit_plasticity_b200.distributed import shard_range
    from vit_plasticity_b200.plasticity import PlasticityEstimator, gather_tables, sweep_local
    from vit_plasticity_b200.preprocess import DevicePreprocessor

    rank, world, dev, pk = ctx.rank, ctx.world, ctx.dev, ctx.pk
    eps = [float(e) for e in args.eps.split(",")]
    ppc = 64
    model = build_vit(ctx, model_name).eval()
    est = PlasticityEstimator(model)
    lo, hi = shard_range(images_total, rank, world)
    n_local = hi - lo
    # resident fp32 images of this rank's shard (seeded per image block, so the data does not depend on the sharding)
    x_dev = torch.empty(n_local, 3, 224, 224, device=dev)
    gen = torch.Generator(device=dev)
    for s0 in range(0, n_local, 256):
        gen.manual_seed(7_000_000 + lo + s0)
        x_dev[s0 : s0 + 256].normal_(generator=gen)
    # warm-up: one chunk through every kernel of the path
    for _ in range(max(1, min(args.warmup, 3))):
        est.sweep_squared_distances(x_dev[: min(ppc, n_local)], x_dev[: min(ppc, n_local)].flip(0), eps)
    torch.cuda.synchronize()
    sampler = ClockSampler(ctx.local)
    if rank == 0 and not secondary:
        sampler.start()
        sampler.mark()
    _lib.reset_launch_count()
    result = {}

    def run_resident(_i):
        result["t"] = sweep_local(est, x_dev, eps, noise_seed=11, first_image=lo, pairs_per_call=ppc)  # [n_eps, rows, n_local] on the device

    ms = ctx.timed(run_resident, 1)
    launches = _lib.launch_count()
    clocks = sampler.stop() if (rank == 0 and not secondary) else None
    # ---- the only collective: ONE gather of the distance tables, outside the device-timed region, timed on its own ----
    gather_s = None
    if world > 1:
        gather_tables(result["t"][:, :1], images_total, (images_total + world - 1) // world)  # untimed: NCCL sets the gather's channels up on first use
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        table = gather_tables(result["t"], images_total, (images_total + world - 1) // world)
        torch.cuda.synchronize()
        gather_s = time.perf_counter() - t0
        if rank == 0:
            assert table.shape == (len(eps), len(est.keys()), images_total)
    n_pairs = images_total * len(eps)
    value = n_pairs / (ms / 1e3)
    # ---- end to end: the shard starts as uint8 CIFAR-shaped samples in pinned host memory, the reference's "test" transform
    # (Resize 224 + ToTensor + Normalize) runs on the device, the distance table is read back to the host ----
    e2e = None
    if not args.no_e2e:
        gcpu = torch.Generator().manual_seed(1000 + rank)
        hu8 = torch.randint(0, 256, (n_local, 32, 32, 3), dtype=torch.uint8, generator=gcpu).pin_memory()
        pre = DevicePreprocessor(224, "test", dev)
        sweep_local(est, hu8[: min(ppc, n_local)], eps, 11, lo, ppc, pre).cpu()
        ms_e = ctx.timed(lambda _i: sweep_local(est, hu8, eps, 11, lo, ppc, pre).cpu(), 1)  # .cpu(): D2H of this rank's table
        n_rows = 1 + 5 * len(model.model.blocks)
        e2e = {"value": round(n_pairs / (ms_e / 1e3), 1), "unit": "pairs/s", "h2d_bytes_per_step": images_total * 32 * 32 * 3,
               "d2h_bytes_per_step": n_rows * images_total * len(eps) * 4, "seconds": round(ms_e / 1e3, 3),
               "source": "uint8 32x32x3 samples in pinned host memory -> device-side Resize/ToTensor/Normalize (bit-exact vs torchvision) -> estimator -> distance table on the host"}
    gf = sweep_gflop(model_name, len(eps))
    achieved = value / world * gf["executed_per_pair"] / 1e3
    out = {"metric": f"ViT-{model_name[0].upper()}/16 plasticity sweep pairs/s", "value": round(value, 1), "unit": "pairs/s", "n_gpus": world,
           "images": images_total, "eps": eps, "pairs": n_pairs, "seconds": round(ms / 1e3, 3), "gather_seconds": round(gather_s, 4) if gather_s is not None else None,
           "e2e": e2e, "gpu_launches": launches,
           "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops_sustained"], 4),
                        "gflop_per_pair_executed": round(gf["executed_per_pair"], 3), "gflop_per_pair_reference_faithful": round(gf["reference_faithful_per_pair"], 3),
                        "reference_faithful_tflops_equivalent": round(value / world * gf["reference_faithful_per_pair"] / 1e3, 1),
                        "note": "achieved = executed FLOPs (f(x) and everything linear in the noise shared by the eps grid) / device time; peak = measured sustained bf16"}}
    if clocks is not None:
        out["clocks"] = clocks
    del x_dev, est, model
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    ctx = Ctx()
    try:
        if args.workload == "sweep":
            eps = [float(e) for e in args.eps.split(",")]
            out = bench_sweep(ctx, args, args.model, args.images)
            if ctx.rank == 0:
                out.update({"steps": args.steps, "warmup": args.warmup, "ms_per_step": round(out["seconds"] * 1e3, 1), "higher_is_better": True, "scaling": "strong",
                            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": sweep_config(args.model, args.images, eps, ctx.world, 64)})
                if not args.no_cpu_baseline and ctx.world == 1:
                    from baseline import reference_arm as R

                    if R.available() is None:
                        cores = os.cpu_count() or 1
                        v, dt = R.plasticity(args.model, 4, 2, cores)
                        out["cpu_baseline"] = {"value": round(v, 4), "unit": "pairs/s", "cores": cores, "kind": "reference",
                                               "sample": f"unmodified reference estimator on 4 ViT-{args.model}/16 pairs, fp32, {cores} threads, {dt:.1f} s"}
                print(json.dumps(out))
        else:
            bench_finetune(ctx, args)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
