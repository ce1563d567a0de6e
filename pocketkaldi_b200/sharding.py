"""Host-side sharding of an utterance corpus across ranks.

Every stage of the path is per-utterance, so ranks take disjoint sets of utterances and never
exchange data; the only cross-rank traffic is the reduction of timing scalars and a final host
gather of results. These helpers are backend-agnostic (`nccl` on the GPU box, `gloo` in the
CPU tests).
"""

import numpy as np


def shard_by_samples(num_samples, world_size):
    """Deals utterances to ranks so that total samples per rank are balanced (longest first,
    always to the least-loaded rank). Returns a list of index arrays, one per rank; indices
    within a rank are ascending so results concatenate in corpus order per rank."""
    num_samples = np.asarray(num_samples, dtype=np.int64)
    order = np.argsort(-num_samples, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    buckets = [[] for _ in range(world_size)]
    for u in order:
        r = int(np.argmin(load))
        buckets[r].append(int(u))
        load[r] += num_samples[u]
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def weak_scaling_ids(rank, utts_per_rank):
    """Weak-scaling benchmark shard: rank r owns utterance ids [r*n, (r+1)*n)."""
    return np.arange(rank * utts_per_rank, (rank + 1) * utts_per_rank, dtype=np.int64)


def reduce_timing(dist, elapsed_ms, frames, device=None):
    """(max over ranks of elapsed_ms, sum over ranks of frames). `dist` is torch.distributed or
    None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(elapsed_ms), float(frames)
    import torch
    kw = {"device": device} if device is not None else {}
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, **kw)
    f = torch.tensor([float(frames)], dtype=torch.float64, **kw)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(t.item()), float(f.item())


def gather_to_rank0(dist, array):
    """Final host gather: rank 0 receives the list of every rank's numpy array (others: None)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [array]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(array, out, dst=0)
    return out
