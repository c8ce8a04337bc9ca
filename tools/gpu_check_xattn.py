"""GPU timing of the routed 32-key cross-attention at the c2 sizes (audio: 48 heads x 64, 13 frames; face: 16 x 128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa
from bya_b200 import ops
torch.manual_seed(0)
Nv, C = 17550, 2
for name, heads, hd, kvf in (("audio", 48, 64, 13), ("face", 16, 128, 1)):
    q = torch.randn(Nv, heads * hd, device="cuda").bfloat16()
    K = torch.randn(C * kvf, heads, 32, hd, device="cuda").bfloat16()
    Vt = torch.randn(C * kvf, heads, hd, 32, device="cuda").bfloat16()
    w = torch.rand(Nv, C, device="cuda")
    out = torch.empty_like(q)
    f = lambda: ops.xattn_kv32(q, K, Vt, w, out, heads, hd, C, kvf, hd ** -0.5)
    for _ in range(3): f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(20): f()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    print(f"{name}: {us:.1f} us  ({2 * q.numel() * 2 / us / 1e3:.0f} GB/s of q read + out write)", flush=True)
