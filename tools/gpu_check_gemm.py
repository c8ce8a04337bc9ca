"""GPU check of the tcgen05 GEMM against torch (run on the B200 box through gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200
from bya_b200 import ops
from bya_b200.lib import lib

torch.manual_seed(0)
dev = "cuda"
print("device", torch.cuda.get_device_name(0), "check", lib().bya_check_device())

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item()

def test_store(M, N, K, bias=True, act=0):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.1).bfloat16() if bias else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, act=act)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b.float()
    if act == 1: ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if act == 2: ref = torch.nn.functional.gelu(ref)
    e = rel(out, ref)
    print(f"store M={M} N={N} K={K} bias={bias} act={act}: rel_err={e:.3e}", "OK" if e < 1e-2 else "FAIL", flush=True)
    return e < 1e-2

def test_resid(M, N, K, split):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.1).bfloat16()
    h = torch.randn(M, N, device=dev).bfloat16()
    ga = torch.randn(N, device=dev); gb = torch.randn(N, device=dev)
    rbs = torch.rand(M, device=dev) * 2
    ref = a.float() @ w.float().t()
    gate = torch.where((torch.arange(M, device=dev) < split)[:, None], ga[None], gb[None])
    ref = h.float() + 0.7 * gate * (ref + b.float()[None] * rbs[:, None])
    out = h.clone()
    ops.gemm(a, w, out, bias=b, mode=ops.EPI_RESIDUAL, resid=out, gate_a=ga, gate_b=gb, split_row=split, alpha=0.7,
             row_bias_scale=rbs)
    torch.cuda.synchronize()
    e = rel(out, ref)
    print(f"resid M={M} N={N} K={K}: rel_err={e:.3e}", "OK" if e < 1e-2 else "FAIL", flush=True)
    return e < 1e-2

def test_qkv(M, D, K, T):
    heads = D // 64
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(3 * D, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(3 * D, device=dev) * 0.1).bfloat16()
    nq = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    nk = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    ang = torch.rand(M - T, 32, device=dev) * 6.28
    cos = ang.cos().repeat_interleave(2, 1).contiguous(); sin = ang.sin().repeat_interleave(2, 1).contiguous()
    out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk)
    torch.cuda.synchronize()
    y = a.float() @ w.float().t() + b.float()
    q, k, v = y[:, :D], y[:, D:2 * D], y[:, 2 * D:]
    def ln_rope(x, g):
        x = x.view(M, heads, 64)
        x = torch.nn.functional.layer_norm(x, (64,), g[0].float(), g[1].float(), 1e-6)
        xv = x[T:]
        xr, xi = xv.reshape(M - T, heads, 32, 2).unbind(-1)
        rot = torch.stack([-xi, xr], -1).flatten(-2)
        xv = xv * cos[:, None] + rot * sin[:, None]
        return torch.cat([x[:T], xv], 0).reshape(M, D)
    ref = torch.cat([ln_rope(q, nq), ln_rope(k, nk), v], 1)
    e = rel(out, ref)
    print(f"qkv M={M} D={D} K={K}: rel_err={e:.3e}", "OK" if e < 1.5e-2 else "FAIL", flush=True)
    return e < 1.5e-2

def bench(M, N, K, act=0, iters=20):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out, bias=b, act=act)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): ops.gemm(a, w, out, bias=b, act=act)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    for _ in range(3): torch.matmul(a, w.t())
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): torch.matmul(a, w.t())
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    tf = 2 * M * N * K / 1e9
    print(f"bench M={M} N={N} K={K}: bya {ms:.3f} ms = {tf/ms:.0f} TF/s | cuBLAS {ms2:.3f} ms = {tf/ms2:.0f} TF/s", flush=True)

ok = True
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "small"):
    ok &= test_store(128, 256, 64, bias=False)
    ok &= test_store(128, 256, 256)
    ok &= test_store(256, 512, 512)
    ok &= test_store(1000, 768, 1024, act=1)
    ok &= test_store(333, 128, 192, act=2)
    ok &= test_store(777, 64, 3072)
    ok &= test_resid(1474, 3072, 3072, 226)
    ok &= test_qkv(1474, 3072, 3072, 226)
if which in ("all", "big"):
    ok &= test_store(17776, 3072, 3072)
    bench(17776, 9216, 3072)
    bench(17776, 3072, 3072)
    bench(17776, 12288, 3072, act=1)
    bench(17776, 3072, 12288)
    bench(35100, 512, 512)
    bench(8192, 8192, 8192)
if which in ("all", "sp"):   # per-rank shapes of 8- and 4-way sequence parallelism (wave-quantised: 256 x 128 pair tiles)
    ok &= test_store(2222, 3072, 12288)
    ok &= test_resid(2222, 3072, 3072, 226)
    for M in (2222, 4444):
        bench(M, 3072, 12288)
        bench(M, 3072, 3072)
        bench(M, 9216, 3072)
        bench(M, 12288, 3072, act=1)
print("ALL OK" if ok else "SOME FAILED")
