#!/bin/bash
# 8-GPU box: scaling lines of bench.py (FP16R default) at 8 and 2 GPUs
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 5 --warmup 3 --no-modes > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?" >> gpurun_out/bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 --no-sub --no-modes > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
