#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python bench.py --precision fp16c8 --no-sub --no-modes > gpurun_out/bench_c8.json 2> gpurun_out/bench_c8.err; echo "rc=$?" >> gpurun_out/bench_c8.err
timeout 300 python bench.py --no-sub > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?" >> gpurun_out/bench_c3.err
timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_c2.json 2>> gpurun_out/bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:cmvn_kernel -c 2 --csv --log-file gpurun_out/ncu_cmvn.csv python bench.py --no-cpu --no-sub --no-e2e --steps 1 --warmup 1 > /dev/null 2>&1
