"""Per-kernel timings at the ViT-B/16 batch-512 shapes (CUDA events, L2 flushed between timed iterations)."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L  # noqa: E402

dev = "cuda"
peaks = {"hbm_gbs": 6539.2, "bf16_tflops": 1627.3}
try:
    peaks.update(json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json")))
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


MODES = [int(c) for c in os.environ.get("KB_GEMM_MODES", "")] if os.environ.get("KB_GEMM_MODES") else None
MODE_NAMES = {0: "single", 1: "pair", 2: "pair/6-stage", 3: "pair/5-stage", 4: "pair+steal"}


def timeit_gemm(name, fn, flops):
    """Times a GEMM call; with KB_GEMM_MODES (e.g. "1230") once per tile mapping, back to back on the same operands, so
    the variants are compared under the same clocks / temperature."""
    if not MODES:
        report(name, timeit(fn), flops=flops)
        return
    before = L.lib().vb_get_gemm_cta_pair()
    parts = []
    for mode in MODES:
        L.lib().vb_set_gemm_cta_pair(1 if mode == 4 else mode)
        L.lib().vb_set_gemm_scheduler(1 if mode == 4 else 0)
        ms = timeit(fn)
        parts.append(f"{MODE_NAMES[mode]} {ms*1e3:7.1f} us ({flops / ms / 1e9:6.0f} TF)")
    L.lib().vb_set_gemm_cta_pair(before)
    L.lib().vb_set_gemm_scheduler(0)
    print(f"{name:34s} " + " | ".join(parts), flush=True)


def rnd(*s, dt=torch.bfloat16, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(dt)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
GEMM_ONLY = os.environ.get("KB_ONLY") == "gemm"
print(f"GEMM tile mapping: {'CTA pairs (256x256, cta_group::2)' if L.lib().vb_get_gemm_cta_pair() else 'single CTAs (128x256)'}", flush=True)
M, E, F3, FF = B * 197, 768, 2304, 3072
res = []


def report(name, ms, flops=None, bytes_=None):
    line = f"{name:34s} {ms*1e3:9.1f} us"
    if flops:
        tf = flops / ms / 1e9
        line += f"  {tf:8.1f} TFLOP/s ({tf/peaks['bf16_tflops']*100:5.1f}% of measured burst)"
    if bytes_:
        gb = bytes_ / ms / 1e6
        line += f"  {gb:8.1f} GB/s ({gb/peaks['hbm_gbs']*100:5.1f}% of measured)"
    print(line, flush=True)
    res.append({"name": name, "ms": ms})


x = rnd(M, E)
for name, n, k, epi in [("fwd qkv  bias", F3, E, L.EPI_BF16), ("fwd proj bias+resid", E, E, L.EPI_BF16_RESID), ("fwd fc1  bias+gelu (+z)", FF, E, L.EPI_BF16_GELU), ("fwd fc1  bias+gelu+gelu' (train)", FF, E, L.EPI_BF16_GELU_GRAD), ("fwd fc1  bias only", FF, E, L.EPI_BF16), ("fwd fc2  bias+resid", E, FF, L.EPI_BF16_RESID)]:
    a = rnd(M, k)
    w = rnd(n, k, scale=0.02)
    bias = torch.randn(n, device=dev)
    out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, n, device=dev, dtype=torch.bfloat16) if epi in (L.EPI_BF16_GELU, L.EPI_BF16_GELU_GRAD) else None
    aux = rnd(M, n) if epi == L.EPI_BF16_RESID else None
    timeit_gemm(name, lambda: L.gemm(a, w, m=M, n=n, k=k, epilogue=epi, bias=bias, aux=aux, out=out, out2=out2), 2.0 * M * n * k)
    ref = timeit(lambda: torch.nn.functional.linear(a, w))
    report("   (torch F.linear bf16, no epi)", ref, flops=2.0 * M * n * k)
    del a, w, out, out2, aux

for name, n_out, k_in, epi in [("dgrad fc2 (+dgelu)", E, FF, L.EPI_BF16_DGELU), ("dgrad fc1", FF, E, L.EPI_BF16), ("dgrad qkv", F3, E, L.EPI_BF16), ("dgrad proj", E, E, L.EPI_BF16)]:
    dy = rnd(M, n_out)
    w = rnd(n_out, k_in, scale=0.02)
    z = rnd(M, k_in) if epi == L.EPI_BF16_DGELU else None
    out = torch.empty(M, k_in, device=dev, dtype=torch.bfloat16)
    timeit_gemm(name, lambda: L.gemm(dy, w, m=M, n=k_in, k=n_out, b_layout=1, epilogue=epi, aux=z, out=out), 2.0 * M * n_out * k_in)
    if epi == L.EPI_BF16_DGELU:  # the training path: saved derivative, one multiply; optionally + fused column sums (fc1 bias grad)
        timeit_gemm("dgrad fc2 (x saved gelu')", lambda: L.gemm(dy, w, m=M, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_MULAUX, aux=z, out=out), 2.0 * M * n_out * k_in)
        cs = torch.zeros(k_in, device=dev)
        timeit_gemm("dgrad fc2 (x gelu' + colsum)", lambda: L.gemm(dy, w, m=M, n=k_in, k=n_out, b_layout=1, epilogue=L.EPI_BF16_MULAUX, aux=z, out=out, out_colsum=cs), 2.0 * M * n_out * k_in)
    del dy, w, z, out

for name, n_out, k_in in [("wgrad fc1", FF, E), ("wgrad fc2", E, FF), ("wgrad qkv", F3, E), ("wgrad proj", E, E)]:
    dy = rnd(M, n_out)
    xx = rnd(M, k_in)
    dw = torch.zeros(n_out, k_in, device=dev)
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    for waves in ((0,) if MODES else (0, 1, 2, 4)):
        sk = max(1, (148 * waves) // tiles) if waves else 0  # 0: the library's own choice
        timeit_gemm(f"{name} split_k={sk}", lambda: L.gemm(dy, xx, m=n_out, n=k_in, k=M, a_layout=1, b_layout=1, epilogue=L.EPI_F32_ADD, out=dw, split_k=sk), 2.0 * M * n_out * k_in)
    ref = timeit(lambda: torch.matmul(dy.T, xx))
    report("   (torch matmul dy^T x)", ref, flops=2.0 * M * n_out * k_in)
    del dy, xx, dw

if GEMM_ONLY:
    sys.exit(0)
g, b = torch.ones(E, device=dev), torch.zeros(E, device=dev)
ms = timeit(lambda: L.layernorm_fwd(x, g, b, 1e-12))
report("layernorm fwd", ms, bytes_=2.0 * M * E * 2)
y, mean, rstd = L.layernorm_fwd(x, g, b, 1e-12)
dy = rnd(M, E)
dg, db = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
ms = timeit(lambda: L.layernorm_bwd(dy, x, g, mean, rstd, dres=dy, dgamma=dg, dbeta=db))
report("layernorm bwd (+dres)", ms, bytes_=4.0 * M * E * 2)
ms = timeit(lambda: torch.nn.functional.layer_norm(x, (E,), g.bfloat16(), b.bfloat16(), 1e-12))
report("   (torch layer_norm fwd)", ms, bytes_=2.0 * M * E * 2)

qkv = rnd(M, 3 * E)
ms = timeit(lambda: L.attention_fwd(qkv, B, 197, 12, 64))
report("attention fwd", ms, flops=4.0 * B * 12 * 197 * 197 * 64)
out, lse = L.attention_fwd(qkv, B, 197, 12, 64)
do = rnd(M, E)
ms = timeit(lambda: L.attention_bwd(qkv, out, do, lse, B, 197, 12, 64))
report("attention bwd", ms, flops=10.0 * B * 12 * 197 * 197 * 64)
q4 = qkv.view(B, 197, 3, 12, 64)
qq, kk, vv = (q4[:, :, i].transpose(1, 2) for i in range(3))
ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
report("   (torch sdpa fwd)", ms, flops=4.0 * B * 12 * 197 * 197 * 64)

ms = timeit(lambda: L.colsum_bf16(qkv, torch.zeros(3 * E, device=dev)))
report("colsum [M,2304]", ms, bytes_=M * 3 * E * 2.0)
img = torch.randn(B, 3, 224, 224, device=dev)
ms = timeit(lambda: L.im2col_patches(img, 16))
report("im2col", ms, bytes_=B * 3 * 224 * 224 * 6.0)
w32 = torch.randn(86_000_000, device=dev)
ms = timeit(lambda: L.cast_f32_to_bf16(w32))
report("cast 86M f32->bf16", ms, bytes_=86e6 * 6)
