#!/bin/bash
# GPU: correctness + timing of the attention kernels over the polynomial-exp2 share
#   general kernel: BYA_FA_POLY16          bounded kernel (BOUNDED=1): BYA_FA_POLY16_BOUNDED
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in ${FA_POLYS:-0 4 8 12}; do
  echo "=== BOUNDED=${BOUNDED:-0} poly16=$p"
  BYA_FA_POLY16=$p BYA_FA_POLY16_BOUNDED=$p timeout 300 python tools/gpu_check_fa.py ${FA_WHICH:-all} 2>&1 | grep -E "fa batch|bench|FAIL|ALL|rror" | grep -v "sdpa 0.1"
done
