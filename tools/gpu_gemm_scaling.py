"""How the small-K GEMM's time splits into fixed overhead and per-tile time (M sweeps whole waves of 148 tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa
from bya_b200 import ops
def bench(M, N, K, iters=20, **kw):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out, bias=b, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): ops.gemm(a, w, out, bias=b, **kw)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
for N, K in ((512, 512), (1536, 512), (512, 2048), (256, 512)):
    for waves in (1, 2, 4, 8):
        tiles_n = max(N // 256, 1)
        M = 128 * 148 * waves // tiles_n
        us = bench(M, N, K)
        print(f"N={N} K={K} M={M} ({waves} waves of 148 tiles): {us:.1f} us -> {us / waves:.1f} us/wave", flush=True)
