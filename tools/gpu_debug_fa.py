import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bya_b200
from bya_b200 import ops
torch.manual_seed(0)
dev = "cuda"
for seq in (64, 128, 192, 256, 320, 448, 1024):
    heads, batch = 1, 1
    D = 64
    qkv = torch.randn(seq, 3 * D, device=dev).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2*D], qkv[:, 2*D:]
    out = torch.zeros(seq, D, device=dev, dtype=torch.bfloat16)
    ops.attention_d64(q, k, v, out, batch, seq, heads)
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(q.float()[None, None], k.float()[None, None], v.float()[None, None])[0, 0]
    err = (out.float() - ref).abs().amax(1)
    blocks = [f"{float(err[i:i+64].max()):.3f}" for i in range(0, seq, 64)]
    # which prefix of keys reproduces the output? try softmax over only first kk keys
    print(f"seq={seq} n_kv={(seq+63)//64} per-64-row max err: {blocks}", flush=True)
    if float(err.max()) > 0.05:
        s = (q.float() @ k.float().t()) * 0.125
        for drop in range(0, seq, 64):
            mask = torch.ones(seq, dtype=torch.bool, device=dev); mask[drop:drop+64] = False
            p = torch.softmax(s[:, mask], -1) @ v.float()[mask]
            e = float((out.float() - p).abs().max())
            if e < 0.02:
                print(f"   -> output equals attention WITHOUT keys [{drop},{drop+64})", flush=True)
        for dup in range(0, seq, 64):
            pass
