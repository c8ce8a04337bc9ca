"""Debug: clock64 timeline of the GEMM kernel's three roles for CTA 0 (BYA_GEMM_TRACE plumbing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa
trace = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
os.environ["BYA_GEMM_TRACE"] = str(trace.data_ptr())
from bya_b200 import ops
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (35100, 512, 512)))
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    trace.zero_()
    ops.gemm(a, w, out, bias=b)
torch.cuda.synchronize()
tr = trace.cpu().reshape(16, 8)
t0 = int(tr[tr > 0].min())
names = ["p:first", "p:last", "m:tempty", "m:kb0", "m:kbN", "e:start", "e:tfull", "e:done"]
print(f"M={M} N={N} K={K}\ntile " + " ".join(f"{n:>9s}" for n in names))
for i in range(8):
    if (tr[i] > 0).any():
        print(f"{i:4d} " + " ".join(f"{int(tr[i, e]) - t0:9d}" if tr[i, e] > 0 else "        -" for e in range(8)))
