"""Times the 3072-wide LayerNorm + adaLN-modulate kernel at the step's shape (17 776 rows) with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bya_b200  # noqa: E402,F401
from bya_b200 import ops  # noqa: E402

rows, dim, split = 17776, 3072, 226
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, dim, device="cuda", dtype=torch.bfloat16, generator=g)
gamma = torch.randn(dim, device="cuda", generator=g).bfloat16()
beta = torch.randn(dim, device="cuda", generator=g).bfloat16()
mods = [torch.randn(dim, device="cuda", generator=g) * 0.1 for _ in range(4)]
outs = {}
for _run in range(2):
    out = torch.empty_like(x)
    for _ in range(3):
        ops.layernorm_modulate(x, out, gamma=gamma, beta=beta, mod_a=(mods[0], mods[1]), mod_b=(mods[2], mods[3]), split_row=split)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        ops.layernorm_modulate(x, out, gamma=gamma, beta=beta, mod_a=(mods[0], mods[1]), mod_b=(mods[2], mods[3]), split_row=split)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 50 * 1e3
    print(f"layernorm_modulate 17776 x 3072: {us:.1f} us  ({2 * rows * dim * 2 / us / 1e3:.0f} GB/s)  checksum {float(out.float().sum()):.6e}")
