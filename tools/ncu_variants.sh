#!/bin/bash
# On the GPU box: cycles and tensor-pipe activity of the 7 GEMM launches of one step under ncu for ONE
# build_variants/<name>.so (one profiler run per gpurun call). Cycle counts are what to compare:
# durations depend on the clock the box happens to run at.
# usage: tools/ncu_variants.sh <name> [VAR=1 ...]
v=$1; shift
cp build_variants/$v.so pocketkaldi_b200/libpkb200.so
env "$@" timeout 600 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:gemm_kernel -s 7 -c 7 --csv --log-file gpurun_out/ncu_$v.csv \
  python bench.py --utts 512 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$v.csv')) if len(r)>10]
h=rows[0]; d={}
for r in rows[1:]:
    d.setdefault(r[h.index('ID')],{})[r[h.index('Metric Name')].split('.')[0]]=float(r[h.index('Metric Value')].replace(',',''))
print('$v $*', ' '.join('%dk/%.0f%%' % (v['sm__cycles_elapsed']/1e3, v['sm__pipe_tensor_cycles_active']) for k,v in sorted(d.items(), key=lambda kv:int(kv[0]))))
PY
