"""Times the fused QKV projection (bias + per-head LayerNorm + RoPE + q pre-scale epilogue) against the same GEMM with a
plain store epilogue, at the c2 shape (17 776 x 9 216 x 3 072).  CUDA events, 20 launches each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa: F401
from bya_b200 import ops
dev = "cuda"
torch.manual_seed(0)
M, D, K, T = 17776, 3072, 3072, 226
a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(3 * D, K, device=dev) * 0.05).bfloat16()
b = (torch.randn(3 * D, device=dev) * 0.1).bfloat16()
nq = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
nk = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
ang = torch.rand(M - T, 32, device=dev) * 6.28
cos = ang.cos().repeat_interleave(2, 1).contiguous(); sin = ang.sin().repeat_interleave(2, 1).contiguous()
out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)

def t(fn, n=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

packed = ops.rope_pack(cos, sin, torch.empty_like(cos), torch.zeros(1, dtype=torch.int32, device=dev))
out2 = torch.empty_like(out)
ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, q_premul=0.18)
ops.gemm(a, w, out2, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, q_premul=0.18, rope_packed=packed)
torch.cuda.synchronize()
print("packed rotary table == full tables, bit for bit:", torch.equal(out, out2), "| mismatch flag", int(packed[1]))
fl = 2.0 * M * 3 * D * K / 1e9
t_qkv = t(lambda: ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, q_premul=0.18))
t_st = t(lambda: ops.gemm(a, w, out, bias=b))
t_norope = t(lambda: ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=M, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, q_premul=0.18))
print(f"QKV epilogue with every row a text row (no RoPE loads): {t_norope:.3f} ms")
t_pk = t(lambda: ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, q_premul=0.18, rope_packed=packed))
print(f"QKV epilogue with the packed rotary table: {t_pk:.3f} ms = {fl / t_pk:.0f} TF/s")
print(f"QKV epilogue {t_qkv:.3f} ms = {fl / t_qkv:.0f} TF/s | plain store {t_st:.3f} ms = {fl / t_st:.0f} TF/s | ratio {t_qkv / t_st:.3f}")
