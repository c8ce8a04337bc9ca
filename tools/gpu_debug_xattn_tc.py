"""Correctness probe of the tensor-memory cross-attention (xattn_tc.cu) against fp32 torch, with error localisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa
from bya_b200 import ops
torch.manual_seed(0)
dev = "cuda"
for heads, hd, chars, kvf, tokens in ((2, 64, 1, 1, 128), (2, 64, 2, 1, 256), (4, 64, 2, 3, 3 * 200), (48, 64, 2, 13, 13 * 1350),
                                      (2, 128, 1, 1, 128), (16, 128, 2, 1, 17550), (3, 128, 2, 2, 2 * 129)):
    q = torch.randn(tokens, heads * hd, device=dev).bfloat16()
    K = torch.randn(chars * kvf, heads, 32, hd, device=dev).bfloat16()
    V = torch.randn(chars * kvf, heads, 32, hd, device=dev).bfloat16()
    w = torch.rand(tokens, chars, device=dev)
    out = torch.full_like(q, float("nan"))
    scale = hd ** -0.5
    ops.xattn_kv32(q, K, V.transpose(-1, -2).contiguous(), w, out, heads, hd, chars, kvf, scale)
    torch.cuda.synchronize()
    tpf = tokens // kvf
    qh = q.float().view(kvf, tpf, heads, hd).permute(0, 2, 1, 3)
    ref = torch.zeros(kvf, heads, tpf, hd, device=dev)
    for c in range(chars):
        kc, vc = K[c * kvf:(c + 1) * kvf].float(), V[c * kvf:(c + 1) * kvf].float()
        p = torch.softmax(qh @ kc.transpose(-1, -2) * scale, -1)
        ref += (p @ vc) * w[:, c].view(kvf, 1, tpf, 1)
    ref = ref.permute(0, 2, 1, 3).reshape(tokens, heads * hd)
    err = (out.float() - ref).abs()
    nan = out.float().isnan()
    rel = float(err[~nan].max() / ref.abs().max()) if (~nan).any() else float("nan")
    print(f"H={heads} d={hd} C={chars} F={kvf} N={tokens}: rel max err {rel:.4f}  nan {int(nan.sum())}/{nan.numel()}", flush=True)
    if rel > 0.02 or nan.any():
        e = torch.where(nan, torch.full_like(err, 9.0), err)
        bycol = e.view(tokens, heads, hd // 32, 32).amax((0, 3))
        print("  max err by (head, 32-col chunk):", [[round(float(x), 3) for x in r] for r in bycol[:4]])
        byrow = e.view(kvf, tpf, -1).amax(2)
        print("  max err by row block of 32 (frame 0):", [round(float(x), 3) for x in byrow[0][: (tpf // 32) * 32].view(-1, 32).amax(1)[:12]])
        print("  sample out[0,:8]", out[0, :8].float().tolist(), "ref", ref[0, :8].tolist())
