#!/usr/bin/env bash
# Builds the drop-in demonstration: the reference's own callers (src/main.cc,
# src/pocketkaldi.cc, src/decoder.cc -- compiled from where they lie, nothing copied)
# against the shim header, linked with the shim implementation, libpkb200.so and the
# reference's unchanged decoder-side objects from oracle/_ref/obj (built by
# oracle/build_ref.sh). Output: oracle/_ref/pocketkaldi_b200_cli (git-ignored, travels to
# the GPU box). The reference's fbank/cmvn/nnet/am/decodable/srfft objects are NOT linked.
set -euo pipefail
REF="${PK_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_shim: $REF/src not found; keeping any prebuilt $OUT/pocketkaldi_b200_cli" >&2
  exit 0
fi
[ -f "$OUT/obj/decoder.o" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$OUT/shim"
CXX="${CXX:-g++}"
FLAGS="-std=c++11 -O2 -g -fPIC -w -I$HERE -I$ROOT/include -I$REF/src -I$OUT/stub"
$CXX $FLAGS -c "$HERE/pkb_shim.cc" -o "$OUT/shim/pkb_shim.o"
for f in pocketkaldi decoder main; do
  $CXX $FLAGS -include "$HERE/pkb_shim.h" -c "$REF/src/$f.cc" -o "$OUT/shim/$f.o"
done
KEEP="util fst matrix pcm_reader strlcpy vector symbol_table hashtable configuration gemm gemm_haswell"
OBJS=""
for f in $KEEP; do OBJS="$OBJS $OUT/obj/$f.o"; done
$CXX -o "$OUT/pocketkaldi_b200_cli" "$OUT/shim/main.o" "$OUT/shim/pocketkaldi.o" \
  "$OUT/shim/decoder.o" "$OUT/shim/pkb_shim.o" $OBJS \
  -L"$ROOT/pocketkaldi_b200" -lpkb200 -Wl,-rpath,'$ORIGIN/../../pocketkaldi_b200' -lm -pthread
# batch driver: list ingestion + one GPU batch + the reference decoder on host threads
$CXX $FLAGS -include "$HERE/pkb_shim.h" -c "$HERE/pkb_batch_main.cc" -o "$OUT/shim/pkb_batch_main.o"
$CXX -o "$OUT/pocketkaldi_b200_batch" "$OUT/shim/pkb_batch_main.o" "$OUT/shim/pocketkaldi.o" \
  "$OUT/shim/decoder.o" "$OUT/shim/pkb_shim.o" $OBJS \
  -L"$ROOT/pocketkaldi_b200" -lpkb200 -Wl,-rpath,'$ORIGIN/../../pocketkaldi_b200' -lm -pthread
# the hot path must come from the shim, not from the reference objects
if nm "$OUT/pocketkaldi_b200_cli" | grep -q "pk_srfft_compute"; then
  echo "build_shim: reference FFT leaked into the shim binary" >&2
  exit 1
fi
echo "build_shim: ok -> $OUT/pocketkaldi_b200_cli, $OUT/pocketkaldi_b200_batch"
