#!/bin/bash
# usage: tools/ncu_kernel.sh <variant .so name under build_variants> <kernel regex> [bench args]
# cycles / duration of the matching launches of one bench step under ncu (compare cycles).
v=$1; k=$2; shift 2
cp build_variants/$v.so pocketkaldi_b200/libpkb200.so
timeout 600 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:$k -c 2 --csv --log-file gpurun_out/ncuk_$v.csv \
  python bench.py ${@:---utts 512} --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncuk_$v.csv')) if len(r)>10]
h=rows[0]; d={}
for r in rows[1:]:
    d.setdefault(r[h.index('ID')],{})[r[h.index('Metric Name')].split('.')[0]]=float(r[h.index('Metric Value')].replace(',',''))
print('$v', ' | '.join('%dk cyc %.0f us occ %.0f%% smem wf %.1fM (conflicts %.1fM) inst %.1fM' % (v['sm__cycles_elapsed']/1e3, v['gpu__time_duration']/1e3, v['sm__warps_active'], v['l1tex__data_pipe_lsu_wavefronts_mem_shared']/1e6, v['l1tex__data_bank_conflicts_pipe_lsu_mem_shared']/1e6, v['smsp__inst_executed']/1e6) for k,v in sorted(d.items(), key=lambda kv:int(kv[0]))))
PY
