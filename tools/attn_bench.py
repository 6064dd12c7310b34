import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L
B = 512
qkv = torch.randn(B * 197, 2304, device="cuda").bfloat16()
out, lse = L.attention_fwd(qkv, B, 197, 12, 64)
do = torch.randn(B * 197, 768, device="cuda").bfloat16()
dbias = torch.zeros(2304, device="cuda")
for name, fn in [("attention_fwd", lambda: L.attention_fwd(qkv, B, 197, 12, 64)), ("attention_bwd", lambda: L.attention_bwd(qkv, out, do, lse, B, 197, 12, 64)), ("attention_bwd + qkv bias grad", lambda: L.attention_bwd(qkv, out, do, lse, B, 197, 12, 64, dbias=dbias))]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / 10 * 1e3:.1f} us")
