"""Headline benchmark: ViT-B/16 bf16 finetuning images/s on B200 (BASELINE.json configs[1]), plus the plasticity
estimator's pairs/s (configs[0] shapes) as a secondary block of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--model base|large] [--components ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference     # the reference algorithm's CPU path (oracle port) on the host cores

A step = forward + cross-entropy + backward + grad-norm clip + SGD-momentum step on one synthetic batch
(apps/vit/train.py:263-283 semantics). ``value`` times K steps with the batch already resident in HBM; ``e2e`` times
the same K steps through the public API with the batch in pinned HOST memory (H2D copy of the images and labels and a
D2H read of the loss inside the timed region, every step). Timing: CUDA events around the K steps, barrier +
synchronize on both sides, max over ranks. L2 (126 MB) cannot carry anything between steps: one step streams
> 30 GB of activations (config.l2 = "inputs>L2").
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
PLAST_GFLOP_PER_PAIR = {"base": 39.577, "large": 137.146}  # executed variant: fc1/fc2/patch/proj on the difference


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (weak scaling)")
    ap.add_argument("--model", default="base")
    ap.add_argument("--components", default="", help="comma-separated components to FREEZE (apps/vit/utils.py:67-74)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="plasticity pairs per call (0 disables the secondary block)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--torch-sgd", action="store_true", help="clip_grad_norm_ + torch.optim.SGD instead of the fused arena step")
    return ap.parse_args()


def workload_config(model_name: str, batch: int, world: int, comps, n_trainable=None):
    """The `config` object shared by both arms (the reference arm runs a bounded sample of the same workload)."""
    cfg = {"workload": f"ViT-{model_name}/16 finetuning (fwd + CE + bwd + clip 1.0 + SGD 1e-2 m0.9), batch {batch}/GPU, 10-class synthetic CIFAR-10-shaped 224x224, random init",
           "global_batch": batch * world, "parallelism": f"dp{world}", "freeze": comps}
    if n_trainable is not None:
        cfg["trainable_params"] = n_trainable
    return cfg


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
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_finetune_img_s(model_name: str, batch: int, steps: int, warmup: int, components):
    import torch

    from oracle import vit_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = O.vit_arch(model_name, n_classes=10)
    sd = O.init_state_dict(arch, seed=42)
    frozen = O.frozen_keys(sd, components)
    x, y = O.synthetic_images(batch, arch, 1), O.synthetic_labels(batch, arch, 2)
    bufs = {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = O.loss_and_grads(sd, x, y, arch, frozen)
        O.sgd_step(sd, bufs, grads, 1e-2, 0.9, 1.0)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    comps = [c for c in args.components.split(",") if c]
    batch = 8  # bounded sample of the batch-512 workload: same step, 8 images per step (CPU minutes otherwise)
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    v, dt, cores = cpu_finetune_img_s(args.model, batch, steps, warmup, comps)
    sample = f"oracle port (torch fp32 CPU) of apps/vit/train.py step, ViT-{args.model}/16, batch {batch} (of the 512 workload), {steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": f"ViT-{args.model[0].upper()}/16 finetune img/s", "value": round(v, 3), "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.model, args.batch, max(1, args.gpus), comps), sample="CPU arm: bounded sample of the workload, 8 images per step"),
        "cpu_baseline": {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from vit_plasticity_b200 import _lib, build_model
    from vit_plasticity_b200.distributed import DataParallel
    from vit_plasticity_b200.finetune import build_optimizer, freeze_model, train_step
    from vit_plasticity_b200.plasticity import PlasticityEstimator

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs CUDA devices: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    comps = [c for c in args.components.split(",") if c]
    pk = peaks()

    torch.manual_seed(42)
    model = build_model({"implementation": "vit", "model_name": args.model, "pretrained": False, "in21k": True, "finetuning": True, "n_classes": 10}, device=dev)
    model.train()
    freeze_model(model, comps)
    n_trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    dp = DataParallel(model) if world > 1 else None
    # apps/vit/configs/cifar10.yaml: SGD lr 1e-2 momentum 0.9, grad_clip 1 — as the fused clip+SGD step over the flat
    # gradient arena (same arithmetic as clip_grad_norm_ + torch.optim.SGD; tests/test_model_gpu.py::test_fused_sgd_*)
    opt = build_optimizer(dp or model, "sgd", lr=1e-2, momentum=0.9, fused=not args.torch_sgd)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2  # distinct pinned host batches, alternated
    host = [(torch.randn(B, 3, 224, 224, generator=g).pin_memory(), torch.randint(0, 10, (B,), generator=g).pin_memory()) for _ in range(n_host)]
    devb = [(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)) for x, y in host]
    after = dp.finish_grad_sync if dp is not None else None

    def step_resident(i):
        return train_step(dp or model, opt, [devb[i % n_host]], grad_clip=1.0, after_backward=after)

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
        loss, _ = train_step(dp or model, opt, [(x, y)], grad_clip=1.0, after_backward=after)
        return float(loss)  # D2H read of the step's result

    def timed(fn_step, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn_step(i)
        e.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- warm-up (untimed) ----
    for i in range(max(3, args.warmup)):
        step_resident(i)
    torch.cuda.synchronize()

    # ---- timed: inputs resident in HBM; GEMM launches individually event-timed for the roofline ----
    if rank == 0:
        sampler.mark()
    _lib.reset_launch_count()
    _lib.GEMM_EVENTS = []
    ms_total = timed(step_resident, args.steps)
    launches = _lib.launch_count()
    gemm_events, _lib.GEMM_EVENTS = _lib.GEMM_EVENTS, None
    clocks = sampler.stop() if rank == 0 else None
    gemm_ms = sum(s.elapsed_time(e) for s, e, _ in gemm_events)
    gemm_flops = sum(f for _, _, f in gemm_events)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- timed: end to end from pinned host memory ----
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            step_e2e(i, last=(i == 1))
        staged.clear()
        ms_e2e = timed(lambda i: step_e2e(i, last=(i == args.steps - 1)), args.steps)
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

        def u_prep(i):
            # host side of the next batch (crop boxes / flips in torchvision's draw order, 1.5 MB uint8 copy) while the
            # GPU is busy with the current step
            ustaged[i] = (hu8[i % n_host].to(dev, non_blocking=True), pre.make_params(B, 32, 32))

        def step_u8(i, last=False):
            if i not in ustaged:
                u_prep(i)
            xu8, params = ustaged.pop(i)
            x = pre.patches(xu8, params)
            loss, _ = train_step(dp or model, opt, [(x, devb[i % n_host][1])], grad_clip=1.0, after_backward=after)
            if not last:
                u_prep(i + 1)
            return float(loss)

        for i in range(2):
            step_u8(i, last=(i == 1))
        ustaged.clear()
        ms_u8 = timed(lambda i: step_u8(i, last=(i == args.steps - 1)), args.steps)
        pipe = {"value": round(world * B * args.steps / (ms_u8 / 1e3), 2), "unit": "img/s", "h2d_bytes_per_step": (B * 32 * 32 * 3 + B * 32) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": round(ms_u8 / args.steps, 3),
                "source": "uint8 32x32x3 batch in pinned host memory; crop boxes / flips drawn on the host in torchvision's RNG order"}

    # ---- secondary: plasticity estimator pairs/s (each rank takes its own shard of pairs; no collective) ----
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
        ms_p = timed(lambda i: est.squared_distances(dx1, dx2), reps)
        # end to end: the pair images start in pinned host memory; the next call's H2D copy runs on the side stream
        # while the current call computes; the distance table is read back to the host every call
        pstaged = {}

        def p_prefetch(i):
            with torch.cuda.stream(copy_stream):
                pstaged[i] = (hx1.to(dev, non_blocking=True), hx2.to(dev, non_blocking=True), torch.cuda.Event())
                pstaged[i][2].record(copy_stream)

        def p_step(i, last):
            if i not in pstaged:
                p_prefetch(i)
            a, b, ev = pstaged.pop(i)
            torch.cuda.current_stream().wait_event(ev)
            a.record_stream(torch.cuda.current_stream())
            b.record_stream(torch.cuda.current_stream())
            if not last:
                p_prefetch(i + 1)
            return est.pair_distances(a, b)  # numpy on the host: D2H of the (1 + 5 n_layers) x P table

        p_step(0, True)
        ms_pe = timed(lambda i: p_step(i, i == reps - 1), reps)
        pps = world * P * reps / (ms_p / 1e3)
        plast = {"metric": f"ViT-{args.model[0].upper()}/16 plasticity pairs/s", "value": round(pps, 1), "unit": "pairs/s", "pairs_per_call": P,
                 "e2e": {"value": round(world * P * reps / (ms_pe / 1e3), 1), "unit": "pairs/s", "h2d_bytes_per_step": 2 * P * 3 * 224 * 224 * 4 * world,
                         "d2h_bytes_per_step": (1 + 5 * len(model.model.blocks)) * P * 4 * world},
                 "frac_of_tensor_peak": round(pps / world * PLAST_GFLOP_PER_PAIR.get(args.model, 0) / 1e3 / pk["bf16_tflops_sustained"], 4),
                 "gflop_per_pair_executed": PLAST_GFLOP_PER_PAIR.get(args.model)}
        model.train()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port on the host cores; bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, cores = cpu_finetune_img_s(args.model, 8, 2, 1, comps)
        cpu = {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "port",
               "sample": f"oracle port (torch fp32) of the same step at batch 8 (of 512), 2 timed steps after 1 warm-up, {dt:.2f} s/step"}

    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_src = ncu_traffic_per_launch()
    step_flops = 3 * FWD_GFLOP_PER_IMG.get(args.model, 0) * 1e9 * B if not comps else None
    out = {
        "metric": f"ViT-{args.model[0].upper()}/16 finetune img/s", "value": round(value, 2), "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": dict(workload_config(args.model, B, world, comps, n_trainable), l2="inputs>L2 (one step streams >30 GB)"),
        "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1) if achieved else None, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": round(achieved / pk["bf16_tflops_sustained"], 4) if achieved else None,
                     "traffic": round(traffic) if traffic else None, "traffic_unit": "bytes per launch (mean over the launches of " + str(traffic_src) + ")" if traffic else None,
                     "flops_per_launch": round(gemm_flops / max(1, len(gemm_events))), "us_per_launch": round(gemm_ms * 1e3 / max(1, len(gemm_events)), 2),
                     "kernel": "gemm_tcgen05_kernel (all fwd/dgrad/wgrad launches of the timed steps, CUDA events per launch)", "peak_source": pk["source"] + " bf16_tflops_sustained",
                     "gemm_share_of_step": round(gemm_ms / ms_total, 4), "gemm_launches": len(gemm_events),
                     "whole_step_frac": round(step_flops * args.steps / (ms_total / 1e3) / 1e12 / pk["bf16_tflops_sustained"], 4) if step_flops else None},
        "cpu_baseline": cpu, "clocks": clocks, "plasticity": plast, "e2e_u8_input_pipeline": pipe,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
