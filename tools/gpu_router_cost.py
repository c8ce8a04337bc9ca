"""In-situ cost of the router chain: the graphed c2 step with the learned router (soft) against the same step with
forced routing masks (the router chain is then skipped, everything else is identical).  usage: python tools/gpu_router_cost.py [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bya_b200  # noqa: F401
from bya_b200.synth import CONFIGS, make_inputs
from bench import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = CONFIGS[name]
dev = torch.device("cuda", 0)
model = build_model(cfg, dev)
model.cache_prologue = False
model.use_cuda_graph = True
res = {}
for forced in (False, True):
    inp = make_inputs(cfg, 1234, device=dev, dtype=torch.bfloat16, forced_masks=forced)
    for _ in range(3):
        model(**inp)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        model(**inp)
    e.record()
    torch.cuda.synchronize()
    res["forced" if forced else "soft"] = s.elapsed_time(e) / 5
print(f"{name}: soft router {res['soft']:.2f} ms/step, forced masks {res['forced']:.2f} ms/step, "
      f"router chain in situ = {res['soft'] - res['forced']:.2f} ms ({100 * (res['soft'] - res['forced']) / res['soft']:.1f} % of the step)")
