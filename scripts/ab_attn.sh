#!/bin/bash
# same-box A/B of the attention kernels: baseline build (experiments/_build/libw2vseg_base.so) vs the in-tree build
for i in 1 2; do
  for lib in experiments/_build/libw2vseg_base.so wav2vecsegmenter_b200/csrc/libw2vseg.so; do
    echo "== $lib"; W2VSEG_LIB=$PWD/$lib python scripts/prof_attn.py 2>&1 | grep -v mma
  done
done
