"""CPU suite: the N>1 path on world_size-2 `gloo` — bucketed gradient averaging == single-process gradient on the
concatenated batch; frozen parameters are excluded from the buckets; pair sharding covers every unit exactly once."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 32), nn.GELU(), nn.Linear(32, 32), nn.LayerNorm(32), nn.Linear(32, 4))


def _worker(rank, world, port, bucket_mb, freeze_first, outdir, overlap=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from vit_plasticity_b200.distributed import DataParallel
    from vit_plasticity_b200.finetune import train_step

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make_model()
        if rank == 1:  # replicas start different on purpose: the wrapper must broadcast rank 0's weights
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        if freeze_first:
            for p in model[0].parameters():
                p.requires_grad = False
        dp = DataParallel(model, bucket_mb=bucket_mb, overlap=overlap)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 16, generator=g), torch.randint(0, 4, (8,), generator=g)
        lo, hi = rank * 4, rank * 4 + 4
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        loss = torch.nn.functional.cross_entropy(dp(x[lo:hi]), y[lo:hi])
        loss.backward()
        dp.finish_grad_sync()
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        # a second step through the public train_step helper (hooks must re-arm after zero_grad)
        train_step(dp, opt, [(x[lo:hi], y[lo:hi])], grad_clip=1.0, after_backward=dp.finish_grad_sync)
        torch.save((rank, grads, float(gnorm), {k: v.clone() for k, v in model.state_dict().items()}, len(dp.buckets), dp.grad_bytes()), os.path.join(outdir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_mb,freeze_first,overlap", [(64, False, True), (0.002, False, True), (0.002, True, True), (0.002, True, False)])
def test_two_rank_gloo_matches_single_process(bucket_mb, freeze_first, overlap, tmp_path):
    """overlap=False: ONE all-reduce of the whole gradient arena after backward instead of one per bucket from the hooks."""
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bucket_mb, freeze_first, str(tmp_path), overlap)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    results = [torch.load(tmp_path / f"r{r}.pt", weights_only=False) for r in range(2)]
    # single-process reference on the concatenated batch
    model = _make_model()
    if freeze_first:
        for p in model[0].parameters():
            p.requires_grad = False
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 16, generator=g), torch.randint(0, 4, (8,), generator=g)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    for _ in range(2):
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        if _ == 0:
            ref_grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        ref_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        if _ == 0:
            first_norm = float(ref_norm)
        opt.step()
        opt.zero_grad()
    for rank, grads, gnorm, sd, n_buckets, nbytes in results:
        assert set(grads) == set(ref_grads)
        for k in ref_grads:
            assert torch.allclose(grads[k], ref_grads[k], atol=1e-6), (rank, k)
        assert abs(gnorm - first_norm) < 1e-5
        for k, v in model.state_dict().items():
            assert torch.allclose(sd[k], v, atol=1e-6), (rank, k)
        trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
        assert nbytes == 4 * trainable  # frozen parameters never enter a bucket
        assert n_buckets >= (2 if bucket_mb < 1 else 1)


def _acc_worker(rank, world, port, outdir, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from vit_plasticity_b200.distributed import DataParallel
    from vit_plasticity_b200.finetune import train_step

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make_model()
        dp = DataParallel(model, bucket_mb=0.002, overlap=overlap)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(16, 16, generator=g), torch.randint(0, 4, (16,), generator=g)
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        for step in range(2):
            lo = step * 8 + rank * 4  # this rank's 4 samples of the step's 8, as two micro-batches of 2
            batches = [(x[lo : lo + 2], y[lo : lo + 2]), (x[lo + 2 : lo + 4], y[lo + 2 : lo + 4])]
            train_step(dp, opt, batches, grad_clip=1.0, after_backward=dp.finish_grad_sync)
        torch.save({k: v.clone() for k, v in model.state_dict().items()}, os.path.join(outdir, f"acc{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_gradient_accumulation_matches_single_process(overlap, tmp_path):
    """grad_acc_steps = 2 under data parallelism (apps/vit/train.py:263-270 accumulates micro-batches before the step): only
    the last micro-batch's backward may launch the bucket all-reduces; replicas stay identical and equal the
    single-process step on the concatenated batch (equal micro-batch sizes: mean of means)."""
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_acc_worker, args=(r, 2, port, str(tmp_path), overlap)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    sds = [torch.load(tmp_path / f"acc{r}.pt", weights_only=False) for r in range(2)]
    model = _make_model()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(16, 16, generator=g), torch.randint(0, 4, (16,), generator=g)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    for step in range(2):
        torch.nn.functional.cross_entropy(model(x[step * 8 : step * 8 + 8]), y[step * 8 : step * 8 + 8]).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
    for k, v in model.state_dict().items():
        assert torch.equal(sds[0][k], sds[1][k]), f"replicas differ in {k}"
        assert torch.allclose(sds[0][k], v, atol=1e-6), k


def test_shard_range_partitions_units():
    from vit_plasticity_b200.distributed import shard_range

    for n, world in [(65536, 8), (64, 8), (10, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_env_contract_without_torchrun(monkeypatch):
    from vit_plasticity_b200 import distributed as D

    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    assert not D.is_distributed_job() and D.get_rank() == 0 and D.get_world_size() == 1 and D.is_master_process()
    mgr = D.build_manager({"device": "cpu", "bogus": 1})
    with mgr as m:
        model = m.build_model(nn.Linear(2, 2))
        assert isinstance(model, nn.Linear)  # single process: no wrapper, as in the reference
    with pytest.raises(NotImplementedError):
        D.build_manager({"device": "cpu", "tp": 2, "dp": 1})


def _gather_worker(rank, world, port, outdir, n_total):
    import numpy as np

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_plasticity_b200.distributed import shard_range
    from vit_plasticity_b200.plasticity import gather_tables

    n_eps, rows = 2, 4
    lo, hi = shard_range(n_total, rank, world)
    # column j of the full table holds j (+ 100 * eps index) in every row: the gathered table must be 0..n_total-1 in order
    local = torch.arange(lo, hi, dtype=torch.float32).expand(n_eps, rows, hi - lo) + 100.0 * torch.arange(n_eps).view(n_eps, 1, 1)
    table = gather_tables(local.contiguous(), n_total, (n_total + world - 1) // world)
    if rank == 0:
        np.save(os.path.join(outdir, "table.npy"), table.numpy())
    else:
        assert table is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 11), (3, 4), (2, 1)])
def test_sweep_gather_reassembles_shards_in_order(tmp_path, world, n_total):
    """The sweep's only collective: rank-ordered contiguous shards (ragged, possibly empty tail shards, every rank padded
    to the common shard size) come back as one [n_eps, rows, n_total] table on rank 0."""
    import numpy as np

    port = _free_port()
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path), n_total), nprocs=world, join=True)
    table = np.load(tmp_path / "table.npy")
    assert table.shape == (2, 4, n_total)
    for e in range(2):
        assert (table[e] == np.arange(n_total, dtype=np.float32) + 100.0 * e).all()


class _FakeProbeModel(nn.Module):
    """Stands in for the Transformer on the CPU: get_pooled_probes returns rows that identify their sample."""

    def __init__(self):
        super().__init__()
        self.p = nn.Parameter(torch.zeros(1))
        self.blocks = [0]

    def get_pooled_probes(self, x, cls_pooling=True, normalize=True):
        v = x.reshape(x.shape[0], -1)[:, :3].float()
        return {"block0_a": v.clone(), "block0_b": torch.cat([v, v], 1)}


def _probe_worker(rank, world, port, outdir):
    import numpy as np

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_plasticity_b200.probing import get_embeddings

    loader = [(torch.arange(i * 100, i * 100 + n * 4, dtype=torch.float32).reshape(n, 4), torch.arange(n) + 10 * i) for i, n in enumerate([5, 3, 1])]
    emb, lab = get_embeddings(_FakeProbeModel(), loader, True, device="cpu")
    if rank == 0:
        np.savez(os.path.join(outdir, "probe.npz"), a=emb["block0_a"], b=emb["block0_b"], lab=lab)
    else:
        assert emb is None and lab is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_probe_features_sharded_and_gathered_in_loader_order(tmp_path, world):
    """Linear-probing features: every rank pools its shard of every batch, rank 0 gets the rows of all
    taps and the labels back in the loader's sample order (ragged batches, a batch smaller than the world size)."""
    import numpy as np

    port = _free_port()
    mp.spawn(_probe_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "probe.npz")
    loader = [(torch.arange(i * 100, i * 100 + n * 4, dtype=torch.float32).reshape(n, 4), torch.arange(n) + 10 * i) for i, n in enumerate([5, 3, 1])]
    ref = torch.cat([x for x, _ in loader])[:, :3].numpy()
    assert np.array_equal(got["a"], ref)
    assert np.array_equal(got["b"], np.concatenate([ref, ref], 1))
    assert np.array_equal(got["lab"], torch.cat([y for _, y in loader]).numpy())
