"""ncu target: a few launches of the joint self-attention kernel at the c2 shape (17 776 tokens x 48 heads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bya_b200  # noqa: F401
from bya_b200 import ops

seq, heads = int(os.environ.get("SEQ", 17776)), int(os.environ.get("HEADS", 48))
batch = int(os.environ.get("BATCH", 1))
torch.manual_seed(0)
D = heads * 64
qkv = (torch.randn(batch * seq, 3 * D, device="cuda") * float(os.environ.get("QSCALE", "1"))).bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
out = torch.zeros(batch * seq, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    (ops.attention_d64(q, k, v, out, batch, seq, heads, score_bound_log2=64.0) if os.environ.get("BOUNDED") == "1" else ops.attention_d64(q, k, v, out, batch, seq, heads))
torch.cuda.synchronize()
print("done")
