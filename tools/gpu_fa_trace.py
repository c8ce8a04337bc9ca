"""Debug: clock64 timeline of the attention kernel's hand-offs (one CTA, first 32 KV tiles).  BYA_FA_TRACE plumbing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa
trace = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
os.environ["BYA_FA_TRACE"] = str(trace.data_ptr())
from bya_b200 import ops
seq, heads = 17776, 48
D = heads * 64
qkv = (torch.randn(seq, 3 * D, device="cuda") * 0.3).bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
out = torch.zeros(seq, D, device="cuda", dtype=torch.bfloat16)
bounded = os.environ.get("BOUNDED", "1") == "1"
for _ in range(2):
    ops.attention_d64(q, k, v, out, 1, seq, heads, score_bound_log2=64.0 if bounded else None)
torch.cuda.synchronize()
tr = trace.cpu().reshape(32, 16)
t0 = int(tr[tr > 0].min())
names = ["s0:sfull", "s0:sfree", "s0:mid", "s0:pvok", "s1:sfull", "s1:sfree", "s1:mid", "s1:pvok",
         "m:sfree0", "m:sfree1", "m:pfull0", "m:pfull1", "m:end"]
print("j  " + " ".join(f"{n:>9s}" for n in names))
for j in range(4, 24):
    print(f"{j:2d} " + " ".join(f"{int(tr[j, e]) - t0:9d}" if tr[j, e] > 0 else "        -" for e in range(13)))
