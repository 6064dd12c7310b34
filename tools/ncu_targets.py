"""Launches each hot kernel a few times so that `ncu -k regex:... -s N -c M` can capture them."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = 512
M, E, FF = B * 197, 768, 3072
dev = "cuda"
rnd = lambda *s: torch.randn(*s, device=dev).bfloat16()
if which in ("all", "gemm"):
    x, w1, b1 = rnd(M, E), rnd(FF, E) * 0.02, torch.randn(FF, device=dev)
    a, z = torch.empty(M, FF, device=dev, dtype=torch.bfloat16), torch.empty(M, FF, device=dev, dtype=torch.bfloat16)
    dy, w2 = rnd(M, E), rnd(E, FF) * 0.02
    dz = torch.empty(M, FF, device=dev, dtype=torch.bfloat16)
    q = torch.empty(M, 3 * E, device=dev, dtype=torch.bfloat16)
    wq, bq = rnd(3 * E, E) * 0.02, torch.randn(3 * E, device=dev)
    for _ in range(3):
        L.gemm(x, wq, m=M, n=3 * E, k=E, epilogue=L.EPI_BF16, bias=bq, out=q)                      # plain
        L.gemm(x, w1, m=M, n=FF, k=E, epilogue=L.EPI_BF16_GELU, bias=b1, out=a, out2=z)             # gelu
        L.gemm(dy, w2, m=M, n=FF, k=E, b_layout=1, epilogue=L.EPI_BF16_DGELU, aux=z, out=dz)        # dgelu
if which in ("all", "attn"):
    qkv = rnd(M, 3 * E)
    for _ in range(3):
        out, lse = L.attention_fwd(qkv, B, 197, 12, 64)
        L.attention_bwd(qkv, out, rnd(M, E), lse, B, 197, 12, 64)
if which in ("all", "ln"):
    x = rnd(M, E)
    g, b = torch.ones(E, device=dev), torch.zeros(E, device=dev)
    for _ in range(3):
        y, mean, rstd = L.layernorm_fwd(x, g, b, 1e-12)
        L.layernorm_bwd(y, x, g, mean, rstd, dres=y, dgamma=torch.zeros(E, device=dev), dbeta=torch.zeros(E, device=dev))
torch.cuda.synchronize()
print("done")
