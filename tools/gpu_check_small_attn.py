"""GPU check + timing of the router's temporal / multi-ID attention at the c2 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bya_b200  # noqa
from bya_b200 import ops
torch.manual_seed(0)
C, Fr, hw, H = 2, 13, 1350, 8
Nv = Fr * hw
qkv = torch.randn(C * Nv, 1536, device="cuda").bfloat16()
out = torch.zeros(C * Nv, 512, device="cuda", dtype=torch.bfloat16)
def timeit(fn, n=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for name, args, Ls in (("temporal", (C * hw, Fr, H, hw, Nv, hw), Fr), ("multi-id", (Nv, C, H, Nv, 0, Nv), C), ("3 ids", None, 3)):
    if args is None:
        continue
    ops.small_attention(qkv, out, *args)
    if name == "temporal":
        x = qkv.float().view(C, Fr, hw, 3, H, 64)
        q, k, v = (x[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v).permute(0, 3, 1, 2, 4).reshape(C * Nv, 512)
    else:
        x = qkv.float().view(C, Nv, 3, H, 64)
        q, k, v = (x[:, :, i].permute(1, 2, 0, 3) for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(C * Nv, 512)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    us = timeit(lambda: ops.small_attention(qkv, out, *args))
    gb = (qkv.numel() + out.numel()) * 2 / 1e9
    print(f"{name}: rel_err={err:.3e} {'OK' if err < 2e-2 else 'FAIL'}  {us:.1f} us = {gb / us * 1e6:.0f} GB/s", flush=True)
