#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
for g in 4 8; do
PKB_CMVN_GROUPS=$g timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:cmvn_kernel -c 1 --csv --log-file gpurun_out/ncu_cmvn_g$g.csv python bench.py --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
PKB_CMVN_GROUPS=$g timeout 300 python bench.py --no-cpu --no-sub --no-e2e > gpurun_out/bench_g$g.json 2>/dev/null
done
PKB_CMVN_GROUPS=8 PKB_CMVN_BLOCKS_PER_SM=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cmvn_kernel -c 1 --csv --log-file gpurun_out/ncu_cmvn_g8b1.csv python bench.py --no-cpu --no-sub --no-e2e --steps 1 --warmup 0 > /dev/null 2>&1
timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_c2.json 2>/dev/null
PKB_CMVN_GROUPS=8 timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_c2_g8.json 2>/dev/null
