#!/bin/bash
# ncu --set full (with source-level stall sampling) of the three HBM-bound kernels furthest from their roofline, one
# report each, taken from a 2-layer step at the full grid.  Reports land in gpurun_out/ (read here with tools/ncu_hot.py).
set -u
mkdir -p gpurun_out
common="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -c 1"
ncu $common -k 'regex:ln_mod_kernel<\(int\)12' -s 3 -o gpurun_out/r1_ln12 -f python tools/profile_step.py c2 L=2 > gpurun_out/ncu_ln12.log 2>&1
ncu $common -k 'regex:xattn_kv32_kernel<\(int\)64' -o gpurun_out/r1_xattn64 -f python tools/profile_step.py c2 L=2 > gpurun_out/ncu_xattn64.log 2>&1
[ -n "${SKIP_SMALL:-}" ] || ncu $common -k 'regex:small_attention_kernel' -s 1 -o gpurun_out/r1_smallattn -f python tools/profile_step.py c2 L=2 > gpurun_out/ncu_smallattn.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu_ln12.log
