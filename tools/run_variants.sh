#!/bin/bash
# On the GPU box: times bench.py's kernel classes with each build_variants/<name>.so in turn.
# A variant name may carry environment settings after a colon: name:VAR=1,VAR2=x
# BENCH_ARGS overrides the bench arguments (default: config 3 at 1024 utterances).
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  cp build_variants/$v.so pocketkaldi_b200/libpkb200.so
  echo "== $spec"
  env $envs timeout 300 python bench.py ${BENCH_ARGS:---utts 1024} --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python tools/_kernel_ms.py
done
