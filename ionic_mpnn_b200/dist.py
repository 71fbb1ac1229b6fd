"""Data-parallel plumbing: one process per GPU, pairs sharded in contiguous ranges, no data-path
collective for inference; one flat all-reduce of the gradient bucket per training step.  torch.distributed is
plumbing only (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of ``n`` pairs owned by ``rank``: the first ``n % world`` ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_records(records, rank, world):
    lo, hi = shard_range(len(records), rank, world)
    return records[lo:hi]


def gather_predictions(local, n_total, group=None):
    """All ranks contribute their shard's predictions ``(n_local, 1)``; every rank gets the full ``(n_total, 1)``
    array in pair order.  Off the timed path (predictions are 4 bytes per pair)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local, np.float32).reshape(-1, 1)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(cap, dtype=torch.float32, device=dev)
    loc = torch.as_tensor(np.asarray(local, np.float32).reshape(-1), device=dev)
    assert loc.numel() == sizes[rank][1] - sizes[rank][0], "local shard size does not match shard_range"
    buf[: loc.numel()] = loc
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[: hi - lo].cpu().numpy() for o, (lo, hi) in zip(out, sizes)]).reshape(-1, 1)


def allreduce_sum_(flat, group=None):
    """In-place sum of the flat gradient bucket over ranks (the only collective of the training step)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def mean_from_summed_bucket(bucket, n_params):
    """What imp_clip_adam_sparse does with a bucket reduced in "sum" mode (train.TrainMixin.train_step): the bucket holds
    sum-gradients [0, n_params) and the tail [sse, pair count, occurrence norm^2 x 2]; after the all-reduce the mean
    gradient is bucket / count, the mean squared error sse / count, and the occurrence norms scale with 1 / count^2.
    Host-side restatement used by the CPU tests of the N > 1 path."""
    count = float(bucket[n_params + 1])
    return bucket[:n_params] / count, float(bucket[n_params]) / count, [float(bucket[n_params + 2]) / count ** 2,
                                                                         float(bucket[n_params + 3]) / count ** 2]
