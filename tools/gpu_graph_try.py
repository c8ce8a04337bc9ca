"""Experiment: how much of the step is launch gaps?  Times the eager step against a CUDA-graph replay of the same step."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bya_b200
from bya_b200.synth import CONFIGS, make_inputs
from bench import build_model

cfg = CONFIGS["c2"]
dev = torch.device("cuda", 0)
model = build_model(cfg, dev)
model.cache_prologue = os.environ.get("CACHE", "0") == "1"
inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16)

def step():
    return model(**inp)[0]

for _ in range(3):
    ref = step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s.record()
for _ in range(3):
    step()
e.record(); t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"eager: {s.elapsed_time(e)/3:.1f} ms/step (host issue time {t_issue/3*1e3:.1f} ms/step)", flush=True)

g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        out = step()
    torch.cuda.synchronize()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    s.record()
    for _ in range(3):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"graph: {s.elapsed_time(e)/3:.1f} ms/step; max|graph - eager| = {(out.float()-ref.float()).abs().max().item():.3e}", flush=True)
except Exception as ex:
    print("graph capture failed:", repr(ex)[:2000])
