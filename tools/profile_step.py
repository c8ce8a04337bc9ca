"""One denoising step between cudaProfilerStart/Stop, for ncu (`--profile-from-start off`).
usage: ncu ... python tools/profile_step.py [config] [L=<layers>] [forced]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dataclasses
import torch
import bya_b200
from bya_b200.synth import CONFIGS, make_inputs
from bench import build_model

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith(("L=", "forced")) else "c2"
cfg = CONFIGS[name]
for a in sys.argv:
    if a.startswith("L="):
        cfg = dataclasses.replace(cfg, num_layers=int(a[2:]))
dev = torch.device("cuda", 0)
model = build_model(cfg, dev)
model.cache_prologue = False
inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16, forced_masks="forced" in sys.argv)
for _ in range(2):
    model(**inp)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(**inp)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
