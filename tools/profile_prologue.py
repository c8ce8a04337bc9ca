"""The per-generation prologue (engine.prologue, own kernels) between cudaProfilerStart/Stop for ncu, and timed as one
CUDA graph replay (no host gaps).  usage: [ncu --profile-from-start off ...] python tools/profile_prologue.py [c2|c3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bya_b200  # noqa: F401
from bench import build_model
from bya_b200 import ops
from bya_b200.synth import CONFIGS, make_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = CONFIGS[name]
import dataclasses

cfg = dataclasses.replace(cfg, num_layers=42)
dev = torch.device("cuda", 0)
model = build_model(cfg, dev)
inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16)
eng = model.engine()
args = (inp["id_cond"], inp["id_vit_hidden"], inp["audio_embeds"], cfg.frames, True)
for _ in range(2):
    eng.prologue(*args)
torch.cuda.synchronize()
l0 = ops.LAUNCHES
torch.cuda.profiler.start()
eng.prologue(*args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", ops.LAUNCHES - l0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.prologue(*args)
for _ in range(3):
    g.replay()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(5):
    flush.zero_()
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print("prologue as one graph replay (L2 flushed): ms", [round(t, 3) for t in ts])
