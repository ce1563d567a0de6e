#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 900 python tools/decode_demo.py --utts 128 --ref --precision fp16c8 > gpurun_out/decode_demo.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active --clock-control none -k regex:fbank_kernel -c 1 --csv --log-file gpurun_out/ncu_fbank.csv python bench.py --utts 512 --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
