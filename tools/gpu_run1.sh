#!/bin/bash
# first GPU pass of round 2: tests, default bench, CMVN residency sweep, ncu over smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
for b in 1 2 3 4; do
  PKB_CMVN_BLOCKS_PER_SM=$b timeout 300 python bench.py --no-cpu --no-e2e --no-sub --steps 5 --warmup 3 > gpurun_out/bench_cmvn_b$b.json 2>> gpurun_out/bench_cmvn.err
done
timeout 300 python bench.py --config 2 --no-cpu --no-e2e > gpurun_out/bench_config2.json 2>> gpurun_out/bench_cmvn.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/smoke_ncu.log
