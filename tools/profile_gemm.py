"""ncu target: the DiT GEMM shapes (and the router's 512-wide one) through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa: F401
from bya_b200 import ops
torch.manual_seed(0)
shapes = [(17776, 12288, 3072, 1), (17776, 3072, 12288, 0), (35100, 512, 512, 2), (35100, 1536, 512, 0)]
for M, N, K, act in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, w, out, bias=b, act=act)
torch.cuda.synchronize()
print("done")
