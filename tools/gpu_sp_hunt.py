"""torchrun --nproc-per-node P tools/gpu_sp_hunt.py : where does the sequence-parallel step first leave the single-GPU step?
Every scratch / symmetric buffer can be poisoned (NaN or zero fill at allocation) to expose reads of never-written memory:
0 x finite garbage is exactly 0 (bit-identical result), 0 x NaN is NaN."""
import os, sys, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bya_b200
from bya_b200 import sp
from bya_b200.synth import CONFIGS, make_inputs
from bench import build_model

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dataclasses.replace(CONFIGS["c2"], num_layers=2, cross_attn_interval=2)
model = build_model(cfg, dev)
inp = make_inputs(cfg, 4321, device=dev, dtype=torch.bfloat16)
ref_taps = {}
ref = model(**inp, taps=lambda k, v: ref_taps.__setitem__(k, v.detach().clone()))[0].clone()
N, T = cfg.n_tokens, cfg.n_tokens - cfg.frames * cfg.grid_h * cfg.grid_w
R = N // world
n0 = rank * R
Tl = min(max(T - n0, 0), R)
v0 = max(n0 - T, 0)
Vl = R - Tl
modes = [("single-gpu nan", "1", True, False), ("nan", "1", True, True), ("none", "0", True, True), ("nan-unpacked", "1", False, True)]
for tag, val, packed, shard in modes:
    os.environ["BYA_POISON_SCRATCH"] = val
    model.packed_rope = packed
    model.invalidate()
    if shard:
        sp.enable(model, dist.group.WORLD)
    else:
        R, n0, Tl, v0, Vl = N, 0, T, 0, N - T
    first = []

    def tap(k, v):
        r = ref_taps.get(k)
        if r is None:
            return
        v = v.detach()
        if k.endswith(".video") or k == "embed_video":
            r = r[v0:v0 + Vl]
        elif k.endswith(".text"):
            return
        if r.shape != v.shape:
            return
        bad = ~((v.float() == r.float()) | (v.float().isnan() & r.float().isnan()))
        nanrows = v.float().isnan().reshape(v.shape[0], -1).any(1) if v.dim() >= 2 else v.float().isnan()
        if bool(bad.any()) or bool(nanrows.any()):
            rows = bad.reshape(bad.shape[0], -1).any(1).nonzero().flatten()
            nr = nanrows.nonzero().flatten()
            first.append(f"{k}: {int(rows.numel())} rows differ (first {rows[:4].tolist()} last {rows[-2:].tolist()}), "
                         f"{int(nr.numel())} rows with NaN (first {nr[:4].tolist()} last {nr[-2:].tolist()}) of {v.shape[0]}")

    if shard:
        R = N // world
        n0 = rank * R
        Tl = min(max(T - n0, 0), R)
        v0 = max(n0 - T, 0)
        Vl = R - Tl
    out = model(**inp, taps=tap)[0]
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    for r_ in range(world):
        dist.barrier()
        if r_ == rank:
            print(f"[{tag}] rank {rank}: out bit_identical={same} nan={bool(out.float().isnan().any())} exchange={getattr(model.engine(), 'sp_exchange', '-')}"
                  f" first bad taps: {first[:3]}", flush=True)
    model._sp_group = None
    model.invalidate()
    torch.cuda.empty_cache()
dist.destroy_process_group()
