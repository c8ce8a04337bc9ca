"""Times one router link fused (bya_gemm_ln_gemm_bf16) against the three kernels it replaces (GEMM + residual, LayerNorm,
GEMM) at the router's shapes.  CUDA events, 20 launches each after 3 warm-ups, operands L2-resident as inside a step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa: F401
from bya_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).bfloat16()


def timeit(fn, n=20):
    """n launches captured into ONE CUDA graph (the host cost of a launch through ctypes, ~20 us, would otherwise hide
    anything shorter), replayed 3 times; the best replay counts."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / n * 1e3)
    return best


for M, N2, act, splits in ((35100, 1536, 0, (1,)), (35100, 512, 2, (1,)), (8775, 1536, 0, (1, 2)), (4388, 1536, 0, (1, 3, 4)),
                           (4388, 512, 2, (1, 2, 4))):
    a1, w1, b1 = rnd(M, 512, s=0.5), rnd(512, 512, s=0.05), rnd(512, s=0.1)
    x, w2, bias2 = rnd(M, 512), rnd(N2, 512, s=0.05), rnd(N2, s=0.1)
    gamma, beta = (1.0 + 0.2 * torch.randn(512, device=dev)).bfloat16(), (0.1 * torch.randn(512, device=dev)).bfloat16()
    wf, csum, b2 = ops.fold_layernorm(w2, bias2, gamma, beta)
    xn = torch.empty_like(x)
    x2 = torch.empty_like(x)
    out2 = torch.empty(M, N2, device=dev, dtype=torch.bfloat16)

    def unfused():
        ops.gemm(a1, w1, x, bias=b1, mode=ops.EPI_RESIDUAL, resid=x)
        ops.layernorm_modulate(x, xn, eps=1e-5, gamma=gamma, beta=beta)
        ops.gemm(xn, w2, out2, bias=bias2, act=act)

    t0 = timeit(unfused)
    line = f"M={M} N2={N2}: unfused (3 kernels) {t0:.1f} us"
    for ns in splits:
        t1 = timeit(lambda: ops.gemm_ln_gemm(a1, w1, b1, x, x2 if ns > 1 else x, wf, csum, b2, out2, ln_eps=1e-5, act=act, n_split=ns))
        flops = 2.0 * M * 512 * (512 + N2)
        line += f" | fused n_split={ns}: {t1:.1f} us ({flops / t1 / 1e6:.0f} TFLOP/s)"
    print(line, flush=True)
