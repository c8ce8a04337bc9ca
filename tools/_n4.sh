run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --steps 1 --warmup 3 --no-loop 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['sp_check'], d['ms_per_step'])"; }
echo "== fused=0"; BYA_ROUTER_FUSED=0 run 29621
echo "== rope packed=0"; BYA_ROPE_PACKED=0 run 29622
echo "== nccl"; BYA_SP_EXCHANGE=nccl run 29623
