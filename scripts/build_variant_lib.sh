#!/bin/bash
# Builds experiments/_build/<name>/libw2vseg.so: the product library with extra -D flags (A/B experiments).
# usage: scripts/build_variant_lib.sh <name> [-DFLAG=...]...   then   W2VSEG_LIB=$PWD/experiments/_build/<name>/libw2vseg.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OUT=$ROOT/experiments/_build/$NAME
mkdir -p "$OUT"
cd "$ROOT/wav2vecsegmenter_b200/csrc"
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       "$@" -c $f -o "$OUT/${f%.cu}.o" &
done
wait
nvcc -shared -o "$OUT/libw2vseg.so" "$OUT"/*.o -cudart static
echo "$OUT/libw2vseg.so"
