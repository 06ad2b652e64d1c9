#!/bin/bash
# gpurun with retries on "transient" (pod busy): usage scripts/gp.sh <timeout_s> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
