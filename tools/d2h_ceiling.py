"""Concurrent pinned device -> host copy bandwidth of this box: the ceiling of bench.py's e2e leg.
  python tools/d2h_ceiling.py                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
         --master-port 29511 tools/d2h_ceiling.py   # eight ranks at the same time
Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


class _Args:
    gpus = int(os.environ.get("WORLD_SIZE", "1"))


class _Env:
    def __init__(self):
        self.args = _Args()
        self.rank, self.world, self.local, self.dist = bench.dist_setup(self.args.gpus)


env = _Env()
if env.dist is None:
    import torch
    torch.cuda.set_device(env.local)
res = bench.d2h_ceiling(env, gib=2, reps=4)
if env.rank == 0:
    res["n_gpus"] = env.world
    print(json.dumps(res), flush=True)
if env.dist is not None:
    env.dist.destroy_process_group()
