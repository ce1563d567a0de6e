#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED pocketkaldi reference from the
# sources where they lie (default /root/reference) into oracle/_ref/, which is
# git-ignored but travels to the GPU box with the gpurun snapshot.
#
#   oracle/_ref/libpkref.so      reference library + oracle/ref_capi.cc wrapper
#   oracle/_ref/pocketkaldi_ref  the reference CLI (src/main.cc)
#   oracle/_ref/obj/*.o          per-source objects (the shim build in
#                                pocketkaldi_b200/shim links the decoder-side ones)
#
# The reference's own autotools build is not used (autotools are absent); the
# source list is the one in Makefile.am:12-30. An empty cblas.h stub satisfies
# matrix.cc:9-11 / vector.cc:5, which include it but reference no cblas symbol.
# No reference source is copied into the repository.
set -euo pipefail
REF="${PK_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF/src not found; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/stub" "$OUT/obj"
: > "$OUT/stub/cblas.h"
CXX="${CXX:-g++}"
CXXFLAGS="-std=c++11 -O2 -g -fPIC -w -I$REF/src -I$OUT/stub"
SRCS="util fst matrix pcm_reader decoder srfft fbank strlcpy cmvn nnet am vector decodable symbol_table pocketkaldi hashtable configuration gemm gemm_haswell"
pids=()
for f in $SRCS; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/src/$f.cc" -nt "$OUT/obj/$f.o" ]; then
    $CXX $CXXFLAGS -c "$REF/src/$f.cc" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX $CXXFLAGS -c "$HERE/ref_capi.cc" -o "$OUT/obj/ref_capi.o"
OBJS=""
for f in $SRCS; do OBJS="$OBJS $OUT/obj/$f.o"; done
$CXX -shared -o "$OUT/libpkref.so" $OBJS "$OUT/obj/ref_capi.o" -lm -pthread
$CXX $CXXFLAGS "$REF/src/main.cc" $OBJS -lm -pthread -o "$OUT/pocketkaldi_ref"
echo "build_ref: ok -> $OUT"
