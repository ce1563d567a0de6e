#!/bin/bash
# Builds an experimental variant of libpkb200.so with extra nvcc defines:
#   tools/build_variant.sh <name> [-DFOO=1 ...]   ->  build_variants/<name>.so
# (all sources are recompiled with the defines; the regular build/ objects are not touched)
set -e
cd "$(dirname "$0")/../pocketkaldi_b200/csrc"
name=$1; shift
out=../../build_variants; mkdir -p $out/obj_$name
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -I../../include -I. --expt-relaxed-constexpr"
objs=""
for f in *.cu; do
  b=${f%.cu}
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $f -o $out/obj_$name/$b.o &
  objs="$objs $out/obj_$name/$b.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $objs
echo built $out/$name.so
