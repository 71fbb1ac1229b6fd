"""CPU, world_size 2, gloo: the sharding / gather / all-reduce plumbing of the N>1 path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ionic_mpnn_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(n_total, rank, world)
        local = (np.arange(lo, hi, dtype=np.float32) * 0.5 + 1.0).reshape(-1, 1)  # stands for model.predict(shard)
        full = D.gather_predictions(local, n_total)
        g = torch.full((7,), float(rank + 1))
        D.allreduce_sum_(g)
        q.put((rank, full[:, 0].tolist(), g.tolist()))
    finally:
        dist.destroy_process_group()


def _bucket_worker(rank, world, port, n_total, q):
    """The training step's single collective with UNEVEN shards: every rank contributes sum-gradients of its own pairs plus
    the tail [sse, pair count, occurrence norms]; the mean over the all-reduced count must equal the global mean."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(n_total, rank, world)
        per_pair = torch.arange(n_total, dtype=torch.float64)[:, None] * torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
        n_params = 3
        bucket = torch.zeros(n_params + 4, dtype=torch.float64)
        bucket[:n_params] = per_pair[lo:hi].sum(0)               # sum-gradients of the local pairs
        bucket[n_params] = float(((per_pair[lo:hi, 0] - 1.0) ** 2).sum())  # local sse
        bucket[n_params + 1] = hi - lo                            # local pair count
        bucket[n_params + 2] = float((per_pair[lo:hi] ** 2).sum())
        D.allreduce_sum_(bucket)
        g, mse, occ = D.mean_from_summed_bucket(bucket, n_params)
        q.put((rank, g.tolist(), mse, occ[0]))
    finally:
        dist.destroy_process_group()


def test_two_rank_bucket_with_uneven_shards_gives_the_global_mean():
    world, n_total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per_pair = np.arange(n_total)[:, None] * np.array([1.0, -2.0, 0.5])
    want_g = per_pair.mean(0)
    want_mse = ((per_pair[:, 0] - 1.0) ** 2).mean()
    want_occ = (per_pair ** 2).sum() / n_total ** 2
    for rank, g, mse, occ in res:
        assert np.allclose(g, want_g) and abs(mse - want_mse) < 1e-12 and abs(occ - want_occ) < 1e-12


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 8, 1000, 16_777_216):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gather_and_allreduce():
    world, n_total = 2, 11  # odd: ranks own 6 and 5 pairs
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (np.arange(n_total) * 0.5 + 1.0).tolist()
    for rank, full, g in res:
        assert full == want
        assert g == [3.0] * 7
