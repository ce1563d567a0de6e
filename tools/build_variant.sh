#!/bin/bash
# Builds an experimental variant of libpkb200.so with extra nvcc defines:
#   tools/build_variant.sh <name> [-DFOO=1 ...]   ->  build_variants/<name>.so
set -e
cd "$(dirname "$0")/../pocketkaldi_b200/csrc"
name=$1; shift
out=../../build_variants; mkdir -p $out/obj_$name
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -I../../include -I. --expt-relaxed-constexpr"
for f in gemm_sm100 nnet; do
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $f.cu -o $out/obj_$name/$f.o &
done
wait
objs=""
for o in build/*.o; do case $o in *gemm_sm100.o|*nnet.o) ;; *) objs="$objs $o";; esac; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $objs $out/obj_$name/gemm_sm100.o $out/obj_$name/nnet.o
echo built $out/$name.so
