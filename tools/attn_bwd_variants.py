"""Attention backward (key-domain tcgen05 kernel) with the three bias-gradient modes: none, all three thirds of the qkv bias
gradient reduced in the kernel, query third only. Isolated timing (CUDA events, L2 flushed between launches), ViT-B shapes.
With VITB200_DBG_TIMING=1 the library prints the stamped timeline of CTA 0 for every launch instead (use `once`).

Usage: python tools/attn_bwd_variants.py [batch] [once]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
once = len(sys.argv) > 2 and sys.argv[2] == "once"
H, S, E = 12, 197, 768
qkv = torch.randn(B * S, 3 * E, device="cuda").bfloat16()
out, lse = L.attention_fwd(qkv, B, S, H, 64)
do = torch.randn(B * S, E, device="cuda").bfloat16()
delta = (do.float() * out.float()).view(B, S, H, 64).sum(-1).permute(0, 2, 1).contiguous()
dbias = torch.zeros(3 * E, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)

variants = {
    "forward": lambda: L.attention_fwd(qkv, B, S, H, 64),
    "no bias gradient": lambda: L.attention_bwd(qkv, None, do, lse, B, S, H, 64, delta=delta),
    "all three thirds in the kernel": lambda: L.attention_bwd(qkv, None, do, lse, B, S, H, 64, dbias=dbias, delta=delta),
    "query third only": lambda: L.attention_bwd(qkv, None, do, lse, B, S, H, 64, dbias=dbias, delta=delta, q_bias_only=True),
}
for name, fn in variants.items():
    if once:
        print(f"== {name}", flush=True)
        fn()
        torch.cuda.synchronize()
        continue
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print(f"{name:34s} {ts[len(ts) // 2]:8.1f} us (min {ts[0]:.1f})", flush=True)
