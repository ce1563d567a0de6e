#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 900 python tools/decode_demo.py --utts 128 --precision fp16c8 > gpurun_out/decode_demo.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
PKB_STREAM_GRAPH=0 timeout 300 python bench.py --config 5 --steps 300 --warmup 20 > gpurun_out/bench_c5_eager.json 2>/dev/null
