#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for g in 4 8; do
PKB_CMVN_GROUPS=$g timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:cmvn_kernel -c 1 --csv --log-file gpurun_out/ncu_cmvn_g$g.csv python bench.py --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
PKB_CMVN_GROUPS=$g timeout 300 python bench.py --no-cpu --no-sub --no-e2e > gpurun_out/bench_g$g.json 2>/dev/null
done
timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_c2.json 2>/dev/null
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cmvn_kernel -c 1 -o gpurun_out/r2_cmvn python bench.py --no-cpu --no-e2e --no-sub --steps 1 --warmup 0 > gpurun_out/ncu_cmvn_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fbank_kernel -c 1 --csv --log-file gpurun_out/ncu_fbank.csv python bench.py --utts 512 --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
