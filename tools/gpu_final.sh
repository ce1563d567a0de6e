#!/bin/bash
# end-of-round evidence: tests, default bench, smoke under ncu, launch list, ncu details
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/smoke_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -c 14 -o gpurun_out/r2_gemm_fp16r python bench.py --utts 512 --steps 1 --warmup 0 --no-cpu --no-e2e --no-sub --no-modes > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'cmvn_kernel|fbank_kernel' -c 2 -o gpurun_out/r2_front python bench.py --utts 512 --steps 1 --warmup 0 --no-cpu --no-e2e --no-sub > gpurun_out/ncu_front.log 2>&1
for c in 2 4 5; do timeout 300 python bench.py --config $c --no-cpu > gpurun_out/bench_config$c.json 2>/dev/null; done
timeout 300 python bench.py --precision fp16c8 --no-sub --no-modes > gpurun_out/bench_fp16c8.json 2>/dev/null
timeout 600 python tools/decode_demo.py --utts 128 --ref --precision fp16r > gpurun_out/decode_demo.txt 2>/dev/null
timeout 300 python tools/refine_margin_stats.py --net 3 --utts 24 > gpurun_out/margin_stats.txt 2>&1
timeout 300 python tools/refine_margin_stats.py --net 4 --utts 12 >> gpurun_out/margin_stats.txt 2>&1
for n in 128 1024; do timeout 120 python tools/viterbi_bench.py --utts $n >> gpurun_out/viterbi_bench.txt 2>&1; done
