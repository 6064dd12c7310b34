"""Times vb_attention_fwd of several builds of libvitb200.so (build/variants/*.so, one per commit, plus the in-tree one)
next to torch's SDPA on the same operands: ViT-B/16 batch 512 (L = 197, 12 heads, d = 64), CUDA events, L2 flushed
between iterations, the variants interleaved round-robin so that clocks / temperature are shared."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
libs = sorted((ROOT / "build" / "variants").glob("libvitb200_*.so")) + [ROOT / "vit_plasticity_b200" / "libvitb200.so"]
B, L, H, D = 512, 197, 12, 64
dev = "cuda"
qkv = torch.randn(B * L, 3 * H * D, device=dev).bfloat16()
out = torch.empty(B * L, H * D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device=dev, dtype=torch.float32)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
q, k, v = (qkv.view(B, L, 3, H, D)[:, :, i].transpose(1, 2) for i in range(3))
handles = []
for p in libs:
    h = ctypes.CDLL(str(p))
    h.vb_attention_fwd.restype = ctypes.c_int32
    h.vb_attention_fwd.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int32] * 4 + [ctypes.c_void_p]
    handles.append((p.stem.replace("libvitb200_", "").replace("libvitb200", "HEAD"), h))


def call(h):
    rc = h.vb_attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, H, D, torch.cuda.current_stream().cuda_stream)
    assert rc == 0


fns = [(name, (lambda h=h: call(h))) for name, h in handles]
fns.append(("torch sdpa fwd", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)))
times = {name: [] for name, _ in fns}
for name, fn in fns:
    for _ in range(3):
        fn()
torch.cuda.synchronize()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 15):
    for name, fn in fns:
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times[name].append(s.elapsed_time(e) * 1e3)
for name, ts in times.items():
    ts.sort()
    print(f"{name:16s} median {ts[len(ts) // 2]:7.1f} us   min {ts[0]:7.1f} us")
