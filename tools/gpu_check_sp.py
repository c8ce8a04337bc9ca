"""torchrun --nproc-per-node P tools/gpu_check_sp.py : sequence-parallel step == single-GPU step (same kernels)."""
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
ok = True
# (forced masks, frames, grid): the last case has an odd number of positions per frame, so the sharded router pads
for forced, frames, gh, gw in ((False, 13, 6, 9), (True, 13, 6, 9), (False, 4, 5, 9)):
    cfg = dataclasses.replace(CONFIGS["c1"], num_layers=2, frames=frames, grid_h=gh, grid_w=gw, cross_attn_interval=2)
    if (cfg.n_tokens % world) or (8 % world):
        continue
    model = build_model(cfg, dev)
    inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16, forced_masks=forced)
    ref = model(**inp)[0].float()
    for mode in (["nccl", "peer"] if os.environ.get("BYA_SP_EXCHANGE") is None else [os.environ["BYA_SP_EXCHANGE"]]):
        model.sp_exchange = mode            # "peer": NVLink peer-memory push / pull + device barrier; "nccl": all_to_all_single
        sp.enable(model)
        for graph in (False, True):
            model.use_cuda_graph = model.sp_cuda_graph = graph
            out = model(**inp)[0].float()
            out = model(**inp)[0].float()   # graph mode: first call captures, second replays
            a, b = out.flatten().double(), ref.flatten().double()
            cos = float((a @ b) / (a.norm() * b.norm()))
            err = float((out - ref).abs().max())
            print(f"rank {rank}/{world} {model.engine().sp_exchange:4s} graph={int(graph)} forced={int(forced)} frames={frames} grid={gh}x{gw}: "
                  f"cos={cos:.7f} max_abs={err:.4e} bit_identical={bool(torch.equal(out, ref))} ref_absmax={float(ref.abs().max()):.3f}",
                  flush=True)
            ok &= cos > 0.9999 and err < 0.05 * float(ref.abs().max()) and model.engine().sp_exchange == mode
        model.use_cuda_graph = model.sp_cuda_graph = False
        model.engine()._graphs.clear()
    del model
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SP OK" if t.item() == 1.0 else "SP FAILED", flush=True)
dist.destroy_process_group()
