#!/bin/bash
# Cycle-accurate A/B of kernel variants on the GPU box (bench.py's step time varies by +-5 % with the
# power state of the box; sm__cycles_elapsed.max of one launch under ncu repeats to 0.1 %).
#   usage: tools/ncu_ab.sh <launch-skip> <variant>[:ENV=1,...] ...     (variants: build_variants/<name>.so,
#          built with tools/build_variant.sh; launch-skip 6 = the FP16 output stage of a config-3 FP16R step)
skip=$1; shift
for spec in "$@"; do
  v=${spec%%:*}; envs="X=1"
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  cp build_variants/$v.so pocketkaldi_b200/libpkb200.so
  env $envs timeout 300 ncu --metrics sm__cycles_elapsed.max --clock-control none -k regex:gemm_kernel \
    --launch-skip $skip -c 1 --csv --log-file gpurun_out/ab_x.csv \
    python bench.py --utts 512 --steps 1 --warmup 0 --no-cpu --no-e2e --no-sub --no-modes > /dev/null 2>&1
  echo "$spec $(grep -E 'sm__cycles_elapsed' gpurun_out/ab_x.csv | awk -F'","' '{print $(NF)}' | tr -d '"') cycles"
done
