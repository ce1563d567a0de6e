#!/bin/bash
# round 2, pass 2: compact output tests, bench with compact e2e, ncu-clean smoke, CMVN ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python -c "
import os
print({k: v for k, v in os.environ.items() if 'INJECT' in k or 'NSIGHT' in k or 'PROFILER' in k or 'NV_' in k})
import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/smoke_ncu.log
# front-end kernels at 512 utterances, full sections, source view
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'cmvn_kernel|fbank_kernel' -c 2 -o gpurun_out/r2_front python bench.py --utts 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > gpurun_out/ncu_front.log 2>&1; echo "ncu front rc=$?" >> gpurun_out/ncu_front.log
