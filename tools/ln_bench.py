"""LayerNorm forward / backward timings at the ViT-B/16 batch-512 shape (CUDA events, L2 flushed), for grid-size A/B:
VB_LN_FWD_BLOCKS_PER_SM=<n> python tools/ln_bench.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L  # noqa: E402

M, E = int(sys.argv[1]) if len(sys.argv) > 1 else 512 * 197, int(sys.argv[2]) if len(sys.argv) > 2 else 768
x = torch.randn(M, E, device="cuda").bfloat16()
dy = torch.randn(M, E, device="cuda").bfloat16()
g, b = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=15):
    for _ in range(3):
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


y, mean, rstd = L.layernorm_fwd(x, g, b, 1e-12)
ms = timeit(lambda: L.layernorm_fwd(x, g, b, 1e-12))
print(f"blocks/SM={os.environ.get('VB_LN_FWD_BLOCKS_PER_SM', 'default')}  layernorm fwd [{M}x{E}] {ms*1e3:7.1f} us  {2.0*M*E*2/ms/1e6:7.0f} GB/s")
dg, db = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
ms = timeit(lambda: L.layernorm_bwd(dy, x, g, mean, rstd, dres=dy, dgamma=dg, dbeta=db))
print(f"                   layernorm bwd (+dres) {ms*1e3:7.1f} us  {4.0*M*E*2/ms/1e6:7.0f} GB/s")
