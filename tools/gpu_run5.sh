#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python bench.py --precision fp16c8 --no-sub --no-cpu --no-e2e > gpurun_out/bench_c8.json 2> gpurun_out/bench_c8.err; echo "rc=$?" >> gpurun_out/bench_c8.err
timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_c2.json 2>> gpurun_out/bench_c8.err
# GEMM kernels of the FP16C8 mode at 512 utterances: full sections (tensor pipe, stalls, DRAM traffic)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -c 7 -o gpurun_out/r2_gemm_c8 python bench.py --precision fp16c8 --utts 512 --steps 1 --warmup 0 --no-cpu --no-e2e --no-sub > gpurun_out/ncu_gemm_c8.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_gemm_c8.log
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cmvn_kernel -c 1 --csv --log-file gpurun_out/ncu_cmvn.csv python bench.py --precision fp16c8 --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
