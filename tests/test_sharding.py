"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the shard
arithmetic, the max-over-ranks timing reduction and the final host gather."""

import os
import socket

import numpy as np
import pytest

from pocketkaldi_b200 import sharding
from pocketkaldi_b200.synth import synth_pcm


def test_shard_by_samples_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    lens = rng.integers(400, 480000, size=257)
    for world in (1, 2, 4, 8):
        shards = sharding.shard_by_samples(lens, world)
        allidx = np.concatenate(shards)
        assert sorted(allidx.tolist()) == list(range(len(lens)))
        loads = [int(lens[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= int(lens.max())
        for s in shards:
            assert np.all(np.diff(s) > 0)


def test_weak_scaling_ids_are_disjoint_and_regenerable():
    a, b = sharding.weak_scaling_ids(0, 16), sharding.weak_scaling_ids(1, 16)
    assert set(a).isdisjoint(set(b)) and a[0] == 0 and b[0] == 16
    # any rank can regenerate any shard bit-for-bit (counter-based PCM)
    assert np.array_equal(synth_pcm(1234, b, 800)[3], synth_pcm(1234, [19], 800)[0])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.weak_scaling_ids(rank, 3)
    pcm = synth_pcm(7, ids, 1000)
    ms, frames = sharding.reduce_timing(dist, 10.0 + 5.0 * rank, 100 * (rank + 1))
    gathered = sharding.gather_to_rank0(dist, pcm.sum(axis=1))
    q.put((rank, ms, frames, None if gathered is None else [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, frames, gathered in res:
        assert ms == 15.0          # max over ranks, seen by every rank
        assert frames == 300.0     # sum over ranks
    assert res[1][3] is None       # only rank 0 receives the gather
    want = [synth_pcm(7, sharding.weak_scaling_ids(r, 3), 1000).sum(axis=1).tolist() for r in range(2)]
    assert res[0][3] == want
