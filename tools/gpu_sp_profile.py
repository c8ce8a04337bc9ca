"""torchrun --nproc-per-node P tools/gpu_sp_profile.py [c2] [L=<layers>]: per-kernel CUDA-event times of the sequence-parallel
exchange path (eager steps; tags: qkv_gemm, self_attention, out_gemm, peer_barrier, peer_pull), for BYA_SP_EXCHANGE=peer|nccl."""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bya_b200  # noqa: F401
from bench import build_model
from bya_b200 import ops, sp
from bya_b200.synth import CONFIGS, make_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = CONFIGS["c2"]
for a in sys.argv[1:]:
    if a.startswith("L="):
        cfg = dataclasses.replace(cfg, num_layers=int(a[2:]))
model = build_model(cfg, dev)
sp.enable(model)
inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    model(**inp)
torch.cuda.synchronize()
tags = ["qkv_gemm", "self_attention", "out_gemm", "peer_barrier", "peer_copy"]
ops.PROFILE = {t: [] for t in tags}
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
s.record()
model(**inp)
e.record()
torch.cuda.synchronize()
res = {t: [a.elapsed_time(b) for a, b in ops.PROFILE[t]] for t in tags}
ops.PROFILE = None
line = f"rank {rank} exchange={model.engine().sp_exchange} eager step {s.elapsed_time(e):.1f} ms | " + " | ".join(
    f"{t}: n={len(v)} mean={1e3 * sum(v) / max(len(v), 1):.1f}us max={1e3 * max(v or [0]):.0f}us sum={sum(v):.1f}ms" for t, v in res.items())
print(line, flush=True)
# the same step as a graph
model.use_cuda_graph = model.sp_cuda_graph = True
for _ in range(3):
    model(**inp)
torch.cuda.synchronize()
dist.barrier()
s.record()
for _ in range(3):
    model(**inp)
e.record()
torch.cuda.synchronize()
print(f"rank {rank} graph step {s.elapsed_time(e) / 3:.1f} ms", flush=True)
model.engine()._graphs.clear()
dist.barrier()
dist.destroy_process_group()
