#!/bin/bash
# GPU: correctness + timing of the attention kernel over its tuning knobs (BYA_FA_POLY16, BYA_FA_SKEW)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BYA_FA_POLY16=${FA_CHECK_POLY:-4} timeout 300 python tools/gpu_check_fa.py small 2>&1 | tail -8
for p in ${FA_POLYS:-0 4 8}; do
  for k in ${FA_SKEWS:-0 1000}; do
    echo "=== BYA_FA_POLY16=$p BYA_FA_SKEW=$k"
    BYA_FA_POLY16=$p BYA_FA_SKEW=$k timeout 300 python tools/gpu_check_fa.py big 2>&1 | grep -E "bench|FAIL|rror"
  done
done
