#!/bin/bash
# Builds experiments/_build/libw2vseg_epi.so: the same library with the clock64 trace hooks of the pair GEMM
# (-DW2VSEG_TRACE, scripts/prof_gemm_trace.py) and the extra epilogue activation variants
# (-DW2VSEG_EPI_EXPERIMENT, scripts/prof_gemm_act.py). Experiments only; the product library is built by
# `python -m wav2vecsegmenter_b200.build`. Use it with W2VSEG_LIB=$PWD/experiments/_build/libw2vseg_epi.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/experiments/_build
mkdir -p "$OUT"
cd "$ROOT/wav2vecsegmenter_b200/csrc"
for f in common gemm_tc gemm_tc2 posconv_tc conv0_tc kernels attention attention_tc attention_tc64 engine; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DW2VSEG_EPI_EXPERIMENT -DW2VSEG_TRACE "$@" -c $f.cu -o "$OUT/$f.o" &
done
wait
nvcc -shared -o "$OUT/libw2vseg_epi.so" "$OUT"/*.o -cudart static
echo "$OUT/libw2vseg_epi.so"
