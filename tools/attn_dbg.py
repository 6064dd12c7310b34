import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_plasticity_b200 import _lib as L
B = 512
qkv = torch.randn(B * 197, 2304, device="cuda").bfloat16()
out, lse = L.attention_fwd(qkv, B, 197, 12, 64)
do = torch.randn(B * 197, 768, device="cuda").bfloat16()
L.attention_bwd(qkv, out, do, lse, B, 197, 12, 64)
torch.cuda.synchronize()
