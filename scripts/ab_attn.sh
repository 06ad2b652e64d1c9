#!/bin/bash
# same-box A/B of the attention kernels: baseline build (experiments/_build/libw2vseg_base.so, if present),
# the in-tree build with the round-1 layout (W2VSEG_ATT64_GROUPS=1) and the in-tree default
for i in 1 2; do
  if [ -f experiments/_build/libw2vseg_base.so ]; then
    echo "== base"; W2VSEG_LIB=$PWD/experiments/_build/libw2vseg_base.so python scripts/prof_attn.py 2>&1 | grep -v mma
  fi
  echo "== in-tree, 1 group per CTA (2 CTAs/SM)"; W2VSEG_ATT64_GROUPS=1 python scripts/prof_attn.py 2>&1 | grep -v mma
  echo "== in-tree default"; python scripts/prof_attn.py 2>&1 | grep -v mma
done
