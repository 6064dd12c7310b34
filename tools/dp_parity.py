"""2-rank NCCL data-parallel parity: DataParallel + FusedSGD (shared arena, in-place gradients, overlapped bucket
all-reduce) on two half batches == one process on the whole batch.   torchrun --nproc-per-node 2 tools/dp_parity.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import vit_oracle as O  # noqa: E402  (checker only: synthetic data + weights)
from vit_plasticity_b200 import build_model  # noqa: E402
from vit_plasticity_b200.distributed import DataParallel  # noqa: E402
from vit_plasticity_b200.finetune import build_optimizer, train_step  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
arch = O.Arch(emb_dim=256, n_heads=4, n_layers=3, ffn_dim=1024, image_dim=(3, 64, 64))
sd = O.init_state_dict(arch, seed=3)
cfg = dict(implementation="transformer", image_dim=arch.image_dim, patch_type="computer_vision", image_patch="hybrid", patch_size=arch.patch_size,
           emb_type="linear", emb_dim=arch.emb_dim, pos_emb=True, n_heads=arch.n_heads, attn_bias=True, activation="gelu", ffn_dim=arch.ffn_dim,
           ffn_bias=True, norm="layer", norm_bias=True, norm_eps=arch.norm_eps, pre_norm=True, n_layers=arch.n_layers, cls_token=True,
           output_type="classification", weight_tying=False, n_classes=arch.n_classes)
B = 8 * world
xs = [O.synthetic_images(B, arch, 100 + i) for i in range(3)]
ys = [O.synthetic_labels(B, arch, 200 + i) for i in range(3)]


def make():
    m = build_model(dict(cfg), device=dev)
    m.load_state_dict(sd)
    return m.train()


from vit_plasticity_b200 import ops  # noqa: E402

FUSED = os.environ.get("DP_FUSED", "1") == "1"
ops.INPLACE_GRADS = os.environ.get("DP_INPLACE", "1") == "1"
model = make()
dp = DataParallel(model, bucket_mb=1)  # several buckets, so the overlap path is exercised
opt = build_optimizer(dp, "sgd", lr=1e-2, momentum=0.9, fused=FUSED)
per = B // world
for x, y in zip(xs, ys):
    train_step(dp, opt, [(x[rank * per:(rank + 1) * per].to(dev), y[rank * per:(rank + 1) * per].to(dev))], grad_clip=1.0, after_backward=dp.finish_grad_sync)
torch.cuda.synchronize()
ok = True
if rank == 0:
    ref = make()
    ropt = build_optimizer(ref, "sgd", lr=1e-2, momentum=0.9, fused=True)
    for x, y in zip(xs, ys):
        train_step(ref, ropt, [(x.to(dev), y.to(dev))], grad_clip=1.0)
    worst = 0.0
    for (k, a), (_, b) in zip(ref.state_dict().items(), model.state_dict().items()):
        worst = max(worst, float((a.double() - b.double()).norm() / a.double().norm().clamp_min(1e-30)))
    print(f"dp_parity: {len(dp.buckets)} buckets, worst relative parameter difference after 3 steps = {worst:.3e}")
    ok = worst <= 1e-5  # per-sample arithmetic is identical; only the fp32 summation order of the batch reduction differs
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
# replicas must stay identical
same = True
for k, p in model.named_parameters():
    mx, mn = p.detach().clone(), p.detach().clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    d = float((mx - mn).abs().max())
    if d != 0.0:
        same = False
        if rank == 0 and d > 1e-6:
            print(f"dp_parity: replicas differ in {k}: max |diff| {d:.3e} (|p| max {float(p.abs().max()):.3e})")
dist.destroy_process_group()
if rank == 0:
    print("dp_parity:", "PASS" if (ok and same) else "FAIL", "(replicas identical)" if same else "(replicas DIVERGED)")
sys.exit(0 if (bool(flag.item()) and same) else 1)
