#!/bin/bash
# usage (under gpurun --gpus N): bash tools/run_multi_gpu.sh N "<tags>"   tags: check c2 c2nccl c3cfg c5
N=$1; TAGS=${2:-"check c2"}; STEPS=${3:-5}
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
mkdir -p gpurun_out
for t in $TAGS; do
  case $t in
    check)  run 29601 tools/gpu_check_sp.py > gpurun_out/sp${N}_r2.log 2>&1; grep -E "SP OK|SP FAILED|Error|error" gpurun_out/sp${N}_r2.log | tail -5; grep -c "bit_identical=True" gpurun_out/sp${N}_r2.log;;
    c2)     run 29602 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_r2_n${N}_c2.json 2> gpurun_out/bench_r2_n${N}_c2.err; grep '^{' gpurun_out/bench_r2_n${N}_c2.json | cut -c1-260; tail -2 gpurun_out/bench_r2_n${N}_c2.err;;
    c2nccl) BYA_SP_EXCHANGE=nccl run 29603 bench.py --gpus $N --steps $STEPS --warmup 3 --no-sp-check --no-loop > gpurun_out/bench_r2_n${N}_c2_nccl.json 2> gpurun_out/bench_r2_n${N}_c2_nccl.err; grep '^{' gpurun_out/bench_r2_n${N}_c2_nccl.json | cut -c1-260; tail -2 gpurun_out/bench_r2_n${N}_c2_nccl.err;;
    c3cfg)  run 29604 bench.py --gpus $N --steps $STEPS --warmup 3 --config c3 --cfg-parallel > gpurun_out/bench_r2_n${N}_c3cfg.json 2> gpurun_out/bench_r2_n${N}_c3cfg.err; grep '^{' gpurun_out/bench_r2_n${N}_c3cfg.json | cut -c1-260; tail -2 gpurun_out/bench_r2_n${N}_c3cfg.err;;
    c3)     run 29606 bench.py --gpus $N --steps $STEPS --warmup 3 --config c3 --no-sp-check > gpurun_out/bench_r2_n${N}_c3.json 2> gpurun_out/bench_r2_n${N}_c3.err; grep '^{' gpurun_out/bench_r2_n${N}_c3.json | cut -c1-260; tail -2 gpurun_out/bench_r2_n${N}_c3.err;;
    c5)     run 29605 bench.py --gpus $N --steps $STEPS --warmup 3 --config c5 --no-loop > gpurun_out/bench_r2_n${N}_c5.json 2> gpurun_out/bench_r2_n${N}_c5.err; grep '^{' gpurun_out/bench_r2_n${N}_c5.json | cut -c1-260; tail -2 gpurun_out/bench_r2_n${N}_c5.err;;
  esac
done
