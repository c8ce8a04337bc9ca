"""-m gpu: each sm_100a kernel through the C ABI vs a plain torch fp32 reference of the same op (tolerances stated
per test: bf16 outputs, fp32 accumulation), incl. ragged / tail shapes; mask path bit-exact vs the C oracle and the
reference-produced goldens."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dev = "cuda"


@pytest.fixture(scope="module")
def ops(built):
    import bya_b200  # noqa: F401
    from bya_b200 import ops as o
    from bya_b200.lib import lib

    assert torch.cuda.is_available()
    assert lib().bya_check_device() == 0, "not an sm_100 device / driver entry point missing"
    return o


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9))


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).bfloat16()


# tolerance: outputs are bf16 (rel. spacing 2^-8 = 3.9e-3); inputs bf16-exact, accumulation fp32 -> 1e-2 of abs-max
TOL = 1e-2


@pytest.mark.parametrize("M,N,K,act", [(128, 256, 64, 0), (1000, 768, 1024, 1), (333, 128, 192, 2), (777, 64, 3072, 0),
                                       (1, 256, 64, 0), (17776, 512, 512, 0)])
def test_gemm_store(ops, M, N, K, act):
    torch.manual_seed(1)
    a, w, b = rnd(M, K, s=0.5), rnd(N, K, s=0.05), rnd(N, s=0.1)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, act=act)
    ref = a.float() @ w.float().t() + b.float()
    ref = F.gelu(ref, approximate="tanh") if act == 1 else F.gelu(ref) if act == 2 else ref
    assert rel(out, ref) < TOL


def _chain_ref(a1, w1, b1, resid, w2, bias2, gamma, beta, eps, act):
    """fp32 torch reference of one router link: X rounded to bf16 (it is stored / LayerNorm-ed as bf16 in the reference
    module too), LayerNorm and the second linear in fp32."""
    x = (resid.float() + a1.float() @ w1.float().t() + b1.float()).bfloat16()
    y = F.layer_norm(x.float(), (512,), gamma.float(), beta.float(), eps) @ w2.float().t() + bias2.float()
    return x, (F.gelu(y) if act == 2 else y)


@pytest.mark.parametrize("M,N2,act,n_split", [(256, 512, 0, 1), (1000, 1536, 0, 1), (35100, 1536, 0, 1), (35100, 512, 2, 1),
                                              (4388, 1536, 0, 3), (4388, 512, 2, 2), (77, 1536, 0, 12), (300, 384, 0, 3)])
def test_gemm_ln_gemm_link(ops, M, N2, act, n_split):
    """bya_gemm_ln_gemm_bf16 (the router's out-projection + residual + next LayerNorm + next projection as one kernel,
    LayerNorm folded into the weights) vs the unfused fp32 arithmetic; rows with a large common offset exercise the
    folded mean subtraction.  Tolerance: bf16 outputs, 1e-2 of abs-max (as the plain GEMM); X itself must match the
    bf16-rounded fp32 value to one bf16 ulp."""
    torch.manual_seed(11)
    a1, w1, b1 = rnd(M, 512, s=0.5), rnd(512, 512, s=0.05), rnd(512, s=0.1)
    resid = (torch.randn(M, 512, device=dev) + torch.randn(M, 1, device=dev) * 3.0).bfloat16()   # per-row mean offsets
    w2, bias2 = rnd(N2, 512, s=0.05), rnd(N2, s=0.1)
    gamma, beta = (1.0 + 0.2 * torch.randn(512, device=dev)).bfloat16(), (0.1 * torch.randn(512, device=dev)).bfloat16()
    eps = 1e-5
    x_ref, y_ref = _chain_ref(a1, w1, b1, resid, w2, bias2, gamma, beta, eps, act)
    wf, csum, b2 = ops.fold_layernorm(w2, bias2, gamma, beta)
    x_out = torch.zeros(M, 512, device=dev, dtype=torch.bfloat16)
    out2 = torch.zeros(M, N2, device=dev, dtype=torch.bfloat16)
    ops.gemm_ln_gemm(a1, w1, b1, resid, x_out, wf, csum, b2, out2, ln_eps=eps, act=act, n_split=n_split)
    torch.cuda.synchronize()
    assert rel(x_out, x_ref) < 5e-3
    assert rel(out2, y_ref) < TOL
    if n_split == 1:   # in place (resid aliases x_out), as the single-GPU router runs it: same bits
        x2 = resid.clone()
        out3 = torch.zeros_like(out2)
        ops.gemm_ln_gemm(a1, w1, b1, x2, x2, wf, csum, b2, out3, ln_eps=eps, act=act)
        assert torch.equal(x2, x_out) and torch.equal(out3, out2)


def test_gemm_ln_gemm_blocked_operands(ops):
    """The sequence-parallel forms: A1 given as K-blocks (all-to-all receive buffer), out2 scattered into column blocks
    (send buffer) — same bits as the plain layout."""
    torch.manual_seed(12)
    M, N2, P = 900, 1536, 4
    a1, w1, b1 = rnd(M, 512, s=0.5), rnd(512, 512, s=0.05), rnd(512, s=0.1)
    resid, w2, bias2 = rnd(M, 512), rnd(N2, 512, s=0.05), rnd(N2, s=0.1)
    gamma, beta = (1.0 + 0.2 * torch.randn(512, device=dev)).bfloat16(), (0.1 * torch.randn(512, device=dev)).bfloat16()
    wf, csum, b2 = ops.fold_layernorm(w2, bias2, gamma, beta)
    x0, y0 = torch.zeros(M, 512, device=dev, dtype=torch.bfloat16), torch.zeros(M, N2, device=dev, dtype=torch.bfloat16)
    ops.gemm_ln_gemm(a1, w1, b1, resid, x0, wf, csum, b2, y0, ln_eps=1e-5)
    kb = 512 // P
    a_blk = a1.view(M, P, kb).permute(1, 0, 2).contiguous()          # [P][M][kb]
    cb = N2 // P
    y_blk = torch.zeros(P, M, cb, device=dev, dtype=torch.bfloat16)  # [dest][M][cb]
    x1 = torch.zeros_like(x0)
    ops.gemm_ln_gemm(a_blk[0], w1, b1, resid, x1, wf, csum, b2, y_blk[0], ln_eps=1e-5, n_split=3, a_kblock=kb,
                     a_kblock_stride=M * kb, col_block=cb, col_block_stride=M * cb)
    torch.cuda.synchronize()
    assert torch.equal(x1, x0)
    assert torch.equal(y_blk.permute(1, 0, 2).reshape(M, N2), y0)


def test_gemm_strided_views(ops):
    """A, out given as column-slice views (row stride > width), as the engine uses them."""
    torch.manual_seed(2)
    big = rnd(300, 1024, s=0.5)
    a = big[:, 256:768]
    w = rnd(256, 512, s=0.05)
    outbig = torch.zeros(300, 1024, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, outbig[:, 512:768])
    assert rel(outbig[:, 512:768], a.float() @ w.float().t()) < TOL
    assert float(outbig[:, :512].abs().max()) == 0.0


def test_gemm_gated_residual(ops):
    torch.manual_seed(3)
    M, N, K, split = 1474, 3072, 3072, 226
    a, w, b, h = rnd(M, K, s=0.5), rnd(N, K, s=0.05), rnd(N, s=0.1), rnd(M, N)
    ga, gb = torch.randn(N, device=dev), torch.randn(N, device=dev)
    rbs = torch.rand(M, device=dev) * 2
    gate = torch.where((torch.arange(M, device=dev) < split)[:, None], ga[None], gb[None])
    ref = h.float() + 0.7 * gate * (a.float() @ w.float().t() + b.float()[None] * rbs[:, None])
    out = h.clone()
    ops.gemm(a, w, out, bias=b, mode=ops.EPI_RESIDUAL, resid=out, gate_a=ga, gate_b=gb, split_row=split, alpha=0.7,
             row_bias_scale=rbs)
    assert rel(out, ref) < TOL


def test_gemm_qkv_layernorm_rope(ops):
    torch.manual_seed(4)
    M, D, K, T = 700, 1024, 512, 226
    heads = D // 64
    a, w, b = rnd(M, K, s=0.5), rnd(3 * D, K, s=0.05), rnd(3 * D, s=0.1)
    nq = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    nk = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    ang = torch.rand(M - T, 32, device=dev) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous()
    out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk)
    y = a.float() @ w.float().t() + b.float()

    def ln_rope(x, g):
        x = F.layer_norm(x.view(M, heads, 64), (64,), g[0].float(), g[1].float(), 1e-6)
        xv = x[T:]
        xr, xi = xv.reshape(M - T, heads, 32, 2).unbind(-1)
        rot = torch.stack([-xi, xr], -1).flatten(-2)
        return torch.cat([x[:T], xv * cos[:, None] + rot * sin[:, None]], 0).reshape(M, D)

    ref = torch.cat([ln_rope(y[:, :D], nq), ln_rope(y[:, D:2 * D], nk), y[:, 2 * D:]], 1)
    assert rel(out, ref) < 1.5e-2
    # packed rotary table (bya_rope_pack: every pair once): same arithmetic, same bits — also with a row offset and the
    # q pre-scale; a table whose pairs differ raises the device-side flag and the epilogue reads the full tables again
    packed = ops.rope_pack(cos, sin, torch.empty_like(cos), torch.zeros(1, dtype=torch.int32, device=dev))
    assert int(packed[1]) == 0 and torch.equal(packed[0][:, :32], cos[:, 0::2]) and torch.equal(packed[0][:, 32:], sin[:, 0::2])
    out2 = torch.empty_like(out)
    ops.gemm(a, w, out2, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), nq=nq, nk=nk, rope_packed=packed)
    assert torch.equal(out2, out)
    cos_o, sin_o = torch.cat([cos[:7] * 0 + 1, cos]), torch.cat([sin[:7] * 0, sin])     # 7 leading rows, skipped by rope_row0
    pk_o = ops.rope_pack(cos_o, sin_o, torch.empty_like(cos_o), torch.zeros(1, dtype=torch.int32, device=dev))
    o3, o4 = torch.empty_like(out), torch.empty_like(out)
    ops.gemm(a, w, o3, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos_o, sin_o), rope_row0=7, nq=nq, nk=nk, q_premul=0.18)
    ops.gemm(a, w, o4, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos_o, sin_o), rope_row0=7, nq=nq, nk=nk, q_premul=0.18,
             rope_packed=pk_o)
    assert torch.equal(o3, o4)
    cos_bad = cos.clone()
    cos_bad[5, 11] += 0.25                                                               # pair (10, 11) of row 5 now differs
    pk_bad = ops.rope_pack(cos_bad, sin, torch.empty_like(cos), torch.zeros(1, dtype=torch.int32, device=dev))
    assert int(pk_bad[1]) == 1
    o5, o6 = torch.empty_like(out), torch.empty_like(out)
    ops.gemm(a, w, o5, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos_bad, sin), nq=nq, nk=nk)
    ops.gemm(a, w, o6, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos_bad, sin), nq=nq, nk=nk, rope_packed=pk_bad)
    assert torch.equal(o5, o6) and not torch.equal(o5, out)


def _gemm_ref_sampled(a, w, rows, cols):
    """fp32 reference on a sample of output rows / columns (the full product of the hot shapes is 0.7-1.3 TFLOP)."""
    return a[rows].float() @ w[cols].float().t()


@pytest.mark.parametrize("M,N,K,act", [(17776, 9216, 3072, 0), (17776, 12288, 3072, 1), (17776, 3072, 12288, 0),
                                       (17776, 3072, 3072, 0), (17550, 2048, 3072, 0), (35100, 1536, 512, 0),
                                       (4444, 9216, 3072, 0), (2222, 3072, 12288, 0)])
def test_gemm_hot_shapes(ops, M, N, K, act):
    """The step's own GEMM shapes at configs[1] (QKV, FFN-in + GELU, FFN-out, attention out, face to_q, router QKV) and the
    per-rank row counts of 4- and 8-way sequence parallelism; every output row block and column block is sampled."""
    torch.manual_seed(21)
    a, w, b = rnd(M, K, s=0.5), rnd(N, K, s=0.03), rnd(N, s=0.1)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, act=act)
    rows = torch.cat([torch.arange(0, M, 61, device=dev), torch.arange(max(M - 130, 0), M, device=dev)])
    ref = a[rows].float() @ w.float().t() + b.float()
    ref = F.gelu(ref, approximate="tanh") if act == 1 else ref
    assert rel(out[rows], ref) < TOL
    assert torch.isfinite(out.float()).all()


def test_gemm_hot_shape_gated_residual_in_place(ops):
    """FFN-out at full size exactly as the engine calls it: K = 12 288, in-place gated residual, text / video gate split."""
    torch.manual_seed(22)
    M, N, K, split = 17776, 3072, 12288, 226
    a, w, b, h = rnd(M, K, s=0.3), rnd(N, K, s=0.02), rnd(N, s=0.1), rnd(M, N)
    ga, gb = torch.randn(N, device=dev), torch.randn(N, device=dev)
    out = h.clone()
    ops.gemm(a, w, out, bias=b, mode=ops.EPI_RESIDUAL, resid=out, gate_a=ga, gate_b=gb, split_row=split)
    rows = torch.cat([torch.arange(0, 400, device=dev), torch.arange(400, M, 53, device=dev)])
    gate = torch.where((rows < split)[:, None], ga[None], gb[None])
    ref = h[rows].float() + gate * (a[rows].float() @ w.float().t() + b.float()[None])
    assert rel(out[rows], ref) < TOL


@pytest.mark.parametrize("P,R", [(2, 1000), (4, 4444), (8, 2222)])
def test_gemm_col_block_and_a_kblock(ops, P, R):
    """The two layouts of the sequence-parallel exchange (engine.py): `col_block` = the QKV GEMM writes
    the all-to-all send buffer [dest][row][cols of dest] through a 3-D TMA store map (with the qk-LayerNorm / RoPE
    epilogue on destination-ordered weight rows); `a_kblock` = the out-projection reads the receive buffer
    [src][row][cols of src] as a K-blocked A operand.  Both against the plain layout of the same GEMM."""
    from bya_b200.sp import qkv_rows_by_destination

    torch.manual_seed(23)
    D, K, T = 1536, 1024, 226        # 24 heads
    Dl = D // P
    a, w, b = rnd(R, K, s=0.5), rnd(3 * D, K, s=0.05), rnd(3 * D, s=0.1)
    nq = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    nk = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    ang = torch.rand(R + 77, 32, device=dev) * 6.28
    cos, sin = ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous()
    plain = torch.empty(R, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, plain, bias=b, mode=ops.EPI_QKV, split_row=T, ln_eps=1e-6, rope=(cos, sin), rope_row0=77, nq=nq, nk=nk,
             q_premul=0.18)
    send = torch.full((P, R, 3 * Dl), 9.0, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, qkv_rows_by_destination(w, D, P), send[0], bias=qkv_rows_by_destination(b, D, P), mode=ops.EPI_QKV,
             split_row=T, ln_eps=1e-6, rope=(cos, sin), rope_row0=77, nq=nq, nk=nk, qkv_block=3 * Dl, col_block=3 * Dl,
             col_block_stride=R * 3 * Dl, q_premul=0.18)
    for d in range(P):   # destination d receives [q | k | v] of its heads: bit-identical to the plain layout's columns
        want = torch.cat([plain[:, j * D + d * Dl: j * D + (d + 1) * Dl] for j in range(3)], 1)
        assert torch.equal(send[d], want), d
    # K-blocked A: recv [P][R][Dl] holds column block s of the logical [R, D] operand
    x = rnd(R, D, s=0.5)
    wo, bo, h = rnd(K, D, s=0.05), rnd(K, s=0.1), rnd(R, K)
    g = torch.randn(K, device=dev)
    recv = x.view(R, P, Dl).permute(1, 0, 2).contiguous()
    o1, o2 = h.clone(), h.clone()
    ops.gemm(x, wo, o1, bias=bo, mode=ops.EPI_RESIDUAL, resid=o1, gate_a=g, gate_b=g)
    ops.gemm(recv[0], wo, o2, bias=bo, mode=ops.EPI_RESIDUAL, resid=o2, gate_a=g, gate_b=g, a_kblock=Dl, a_kblock_stride=R * Dl)
    assert torch.equal(o1, o2)
    assert rel(o1, h.float() + g * (x.float() @ wo.float().t() + bo.float())) < TOL


def test_attention_d64_full_sequence_48_heads(ops):
    """The joint self-attention at its real size (17 776 tokens x 48 heads, bounded-score kernel, q pre-scaled as the
    QKV epilogue leaves it) vs fp32 math on sampled heads and query rows; the general kernel on the same inputs."""
    torch.manual_seed(24)
    seq, heads = 17776, 48
    D = heads * 64
    x = torch.randn(seq, 3 * heads, 64, device=dev)
    x[:, :2 * heads] = F.layer_norm(x[:, :2 * heads], (64,)) * 1.2
    x[:, :heads] *= 0.125 * 1.4426950408889634
    qkv = x.reshape(seq, 3 * D).bfloat16()
    del x
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    bound = 1.02 * (8 * 1.2) ** 2 * 0.125 * 1.4426950408889634
    out = torch.zeros(seq, D, device=dev, dtype=torch.bfloat16)
    ops.attention_d64(q, k, v, out, 1, seq, heads, score_bound_log2=bound)
    out2 = torch.zeros_like(out)
    ops.attention_d64(q, k, v, out2, 1, seq, heads, scale=math.log(2.0))
    rows = torch.cat([torch.arange(0, seq, 97, device=dev), torch.arange(seq - 160, seq, device=dev)])
    for h in (0, 1, 17, 31, 47):
        sl = slice(h * 64, (h + 1) * 64)
        p = torch.softmax(q[rows, sl].float() @ k[:, sl].float().t() * math.log(2.0), -1)
        ref = p @ v[:, sl].float()
        assert rel(out[rows, sl], ref) < TOL, h
        assert rel(out2[rows, sl], ref) < TOL, h
    assert torch.isfinite(out.float()).all()


def test_gemm_rejects_bad_shapes(ops):
    a = rnd(128, 100)
    with pytest.raises(RuntimeError):
        ops.gemm(a, rnd(256, 100), torch.empty(128, 256, device=dev, dtype=torch.bfloat16))  # K % 64 != 0
    with pytest.raises(RuntimeError):
        ops.gemm(rnd(128, 64), rnd(100, 64), torch.empty(128, 100, device=dev, dtype=torch.bfloat16))  # N % 64 != 0


@pytest.mark.parametrize("batch,seq,heads,qs", [(1, 128, 1, 1.0), (1, 1000, 3, 1.0), (2, 1350, 8, 1.0), (1, 1474, 48, 3.0),
                                               (3, 70, 2, 1.0), (1, 257, 1, 6.0)])
def test_attention_d64(ops, batch, seq, heads, qs):
    """incl. ragged sequence lengths (tail keys masked, tail queries dropped) and large logits (lazy rescale path)."""
    torch.manual_seed(5)
    D = heads * 64
    qkv = rnd(batch * seq, 3 * D, s=qs)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    out = torch.zeros(batch * seq, D, device=dev, dtype=torch.bfloat16)
    ops.attention_d64(q, k, v, out, batch, seq, heads)

    def hv(x):
        return x.reshape(batch, seq, heads, 64).transpose(1, 2).float()

    ref = F.scaled_dot_product_attention(hv(q), hv(k), hv(v)).transpose(1, 2).reshape(batch * seq, D)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < 2e-2


@pytest.mark.parametrize("batch,seq,heads", [(1, 128, 1), (1, 1000, 3), (2, 1350, 8), (1, 1474, 48), (3, 77, 2)])
def test_attention_d64_bounded(ops, batch, seq, heads):
    """Max-free kernel (`bya_attention_d64_bounded`): q carries scale*log2(e), |q.k| <= bound.  Reference: the same
    softmax in fp32 (softmax_2(s) = softmax(s ln 2)).  Rows/keys past the sequence end and batch boundaries are covered
    by the non-multiple-of-128 lengths."""
    torch.manual_seed(3)
    D = heads * 64
    x = torch.randn(batch * seq, 3 * heads, 64, device=dev)
    x[:, :2 * heads] = F.layer_norm(x[:, :2 * heads], (64,)) * 1.3            # |q| = |k| = 8 * 1.3 like qk-LayerNorm
    x[:, :heads] *= 0.125 * 1.4426950408889634
    qkv = x.reshape(batch * seq, 3 * D).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    bound = 1.02 * (8 * 1.3) ** 2 * 0.125 * 1.4426950408889634
    out = torch.zeros(batch * seq, D, device=dev, dtype=torch.bfloat16)
    ops.attention_d64(q, k, v, out, batch, seq, heads, score_bound_log2=bound)
    hv = lambda t: t.reshape(batch, seq, heads, 64).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(hv(q), hv(k), hv(v), scale=math.log(2.0)).transpose(1, 2).reshape(batch * seq, D)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < TOL
    # the general kernel on the same (pre-scaled) inputs agrees: scale = 1 / log2(e) undoes the folded log2(e)
    out2 = torch.zeros_like(out)
    ops.attention_d64(q, k, v, out2, batch, seq, heads, scale=math.log(2.0))
    assert rel(out2, out) < TOL


def test_attention_d64_strided_batches(ops):
    """Sequences `seq_stride` rows apart with padding rows in between (sequence-parallel router frames): the padding
    must neither be attended to nor written."""
    torch.manual_seed(4)
    batch, seq, stride, heads = 5, 1350, 1352, 2
    D = heads * 64
    qkv = rnd(batch * stride, 3 * D)
    qkv.view(batch, stride, 3 * D)[:, seq:] = 1e4          # poison the padding rows
    out = torch.full((batch * stride, D), 7.0, device=dev, dtype=torch.bfloat16)
    ops.attention_d64(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch, seq, heads, seq_stride=stride)
    x = qkv.view(batch, stride, 3, heads, 64)[:, :seq].float()
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(batch, seq, D)
    o = out.view(batch, stride, D)
    assert rel(o[:, :seq], ref) < 2e-2
    assert bool((o[:, seq:] == 7.0).all())


def test_attention_d64_bounded_rejects_bad_bound(ops):
    qkv = rnd(128, 192)
    out = torch.zeros(128, 64, device=dev, dtype=torch.bfloat16)
    for bad in (0.0, -1.0, 64.5, float("nan")):
        with pytest.raises(RuntimeError):
            ops.attention_d64(qkv[:, :64], qkv[:, 64:128], qkv[:, 128:], out, 1, 128, 1, score_bound_log2=bad)


@pytest.mark.parametrize("dim", [512, 768, 1024, 2048, 3072])
def test_layernorm_modulate(ops, dim):
    torch.manual_seed(6)
    rows, split = 300, 37
    x = rnd(rows, dim, s=2.0) + 0.5
    g, b = rnd(dim), rnd(dim, s=0.1)
    sa, ha, sb, hb = (torch.randn(dim, device=dev) * 0.3 for _ in range(4))
    add = rnd(50, dim)
    out = torch.empty_like(x)
    ops.layernorm_modulate(x, out, eps=1e-5, gamma=g, beta=b, mod_a=(sa, ha), mod_b=(sb, hb), split_row=split, add=add)
    y = F.layer_norm(x.float(), (dim,), g.float(), b.float(), 1e-5)
    cls = (torch.arange(rows, device=dev) < split)[:, None]
    y = y * (1 + torch.where(cls, sa[None], sb[None])) + torch.where(cls, ha[None], hb[None])
    y = y + add.float()[torch.arange(rows, device=dev) % 50]
    assert rel(out, y) < TOL
    out2 = torch.empty_like(x)
    ops.layernorm_modulate(x, out2)  # plain, no affine
    assert rel(out2, F.layer_norm(x.float(), (dim,))) < TOL
    # the operand combinations that have their own compiled variant, and one that takes the run-time-flag path
    ops.layernorm_modulate(x, out2, gamma=g, beta=b)
    assert rel(out2, F.layer_norm(x.float(), (dim,), g.float(), b.float())) < TOL
    ops.layernorm_modulate(x, out2, gamma=g, beta=b, add=add)
    assert rel(out2, F.layer_norm(x.float(), (dim,), g.float(), b.float()) + add.float()[torch.arange(rows, device=dev) % 50]) < TOL
    ops.layernorm_modulate(x, out2, mod_a=(sa, ha), mod_b=(sb, hb), split_row=split)
    assert rel(out2, F.layer_norm(x.float(), (dim,)) * (1 + torch.where(cls, sa[None], sb[None])) + torch.where(cls, ha[None], hb[None])) < TOL
    ops.layernorm_modulate(x, out2, gamma=g, add=add)   # gamma without beta
    assert rel(out2, F.layer_norm(x.float(), (dim,)) * g.float() + add.float()[torch.arange(rows, device=dev) % 50]) < TOL
    if dim >= 2048:   # many rows, ragged row count (not a multiple of the 8 rows per block), strided views
        rows2 = 1031
        big = rnd(rows2, dim + 64, s=2.0) + 0.25
        xv = big[:, :dim]
        outb = torch.zeros(rows2, dim + 64, device=dev, dtype=torch.bfloat16)
        ops.layernorm_modulate(xv, outb[:, :dim], eps=1e-6, gamma=g, beta=b, mod_a=(sa, ha), mod_b=(sb, hb), split_row=226)
        y = F.layer_norm(xv.float(), (dim,), g.float(), b.float(), 1e-6)
        cls = (torch.arange(rows2, device=dev) < 226)[:, None]
        y = y * (1 + torch.where(cls, sa[None], sb[None])) + torch.where(cls, ha[None], hb[None])
        assert rel(outb[:, :dim], y) < TOL
        assert bool((outb[:, dim:] == 0).all())
        ops.layernorm_modulate(xv, outb[:, :dim], mod_b=(sb, hb))          # only the video class is modulated
        assert rel(outb[:, :dim], F.layer_norm(xv.float(), (dim,)) * (1 + sb[None]) + hb[None]) < TOL


def test_gemv_and_timestep_features(ops):
    torch.manual_seed(7)
    for B in (1, 2):
        w, b = rnd(1000, 512, s=0.05), rnd(1000, s=0.1)
        x = torch.randn(B, 512, device=dev)
        y = torch.empty(B, 1000, device=dev)
        ops.gemv(w, b, x, y, in_act=1, out_act=1)
        ref = F.silu(F.silu(x) @ w.float().t() + b.float())
        assert float((y - ref).abs().max()) < 1e-4
    t = torch.tensor([500, 999], device=dev, dtype=torch.int64)
    out = torch.empty(2, 3072, device=dev)
    ops.timestep_features(t, out)
    half = 1536
    fr = torch.exp(-np.log(10000.0) * torch.arange(half, device=dev, dtype=torch.float32) / half)
    ang = t[:, None].float() * fr[None]
    assert float((out - torch.cat([ang.cos(), ang.sin()], -1)).abs().max()) < 2e-4


def test_patchify_unpatchify(ops):
    torch.manual_seed(8)
    Fr, C, H, W = 3, 48, 16, 24
    lat = rnd(Fr, C, H, W)
    out = torch.empty(Fr * (H // 2) * (W // 2), 192, device=dev, dtype=torch.bfloat16)
    ops.patchify(lat, out)
    ref = F.unfold(lat.float(), kernel_size=2, stride=2).transpose(1, 2).reshape(-1, C * 4)  # columns (c, dy, dx)
    assert torch.equal(out.float(), ref)
    y = rnd(Fr * 8 * 12, 64)
    img = torch.empty(Fr, 16, 16, 24, device=dev, dtype=torch.bfloat16)
    ops.unpatchify(y, img)
    ref = y.reshape(1, Fr, 8, 12, 16, 2, 2).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)[0]
    assert torch.equal(img, ref)


@pytest.mark.parametrize("heads,hd,chars,kvf,tokens", [(16, 128, 2, 1, 1248), (48, 64, 2, 13, 13 * 96), (4, 64, 3, 5, 5 * 37),
                                                       (2, 128, 1, 1, 50), (3, 128, 3, 2, 2 * 129)])
def test_routed_cross_attention(ops, heads, hd, chars, kvf, tokens):
    torch.manual_seed(9)
    q = rnd(tokens, heads * hd)
    K = rnd(chars * kvf, heads, 32, hd)
    V = rnd(chars * kvf, heads, 32, hd)
    w = torch.rand(tokens, chars, device=dev)
    out = torch.empty_like(q)
    scale = hd ** -0.5
    ops.xattn_kv32(q, K, V.transpose(-1, -2).contiguous(), w, out, heads, hd, chars, kvf, scale)
    tpf = tokens // kvf
    qh = q.float().view(kvf, tpf, heads, hd).permute(0, 2, 1, 3)  # [f,h,t,d]
    ref = torch.zeros(kvf, heads, tpf, hd, device=dev)
    for c in range(chars):
        kc, vc = K[c * kvf:(c + 1) * kvf].float(), V[c * kvf:(c + 1) * kvf].float()
        p = torch.softmax(qh @ kc.transpose(-1, -2) * scale, -1)
        ref += (p @ vc) * w[:, c].view(kvf, 1, tpf, 1)
    ref = ref.permute(0, 2, 1, 3).reshape(tokens, heads * hd)
    assert rel(out, ref) < 1.5e-2


@pytest.mark.parametrize("heads,hd,chars,kvf,tpf,shards", [(48, 64, 2, 13, 150, 4), (16, 128, 2, 1, 1350, 8), (6, 64, 3, 5, 77, 2),
                                                          (4, 64, 1, 3, 200, 3)])
def test_routed_cross_attention_on_token_shards(ops, heads, hd, chars, kvf, tpf, shards):
    """Sequence-parallel use (`tok_begin`, `total_tokens`): a rank passes only its rows, which start and end in the middle
    of frames.  Every shard must reproduce its rows of the full-clip call BIT FOR BIT (the multi-GPU step is compared with
    the single-GPU step that way), and must not touch a row outside the shard."""
    torch.manual_seed(11)
    tokens = kvf * tpf
    q = rnd(tokens, heads * hd)
    K = rnd(chars * kvf, heads, 32, hd)
    Vt = rnd(chars * kvf, heads, hd, 32)
    w = torch.rand(tokens, chars, device=dev)
    w[::3, 0] = 0.0                      # hard-routed tokens: a character with weight 0 contributes nothing
    full = torch.empty_like(q)
    ops.xattn_kv32(q, K, Vt, w, full, heads, hd, chars, kvf, hd ** -0.5)
    bounds = [round(i * tokens / shards) for i in range(shards + 1)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        guard = 5
        buf = torch.full((b - a + 2 * guard, heads * hd), 7.0, device=dev, dtype=torch.bfloat16)
        part = buf[guard:guard + b - a]
        ops.xattn_kv32(q[a:b].contiguous(), K, Vt, w[a:b].contiguous(), part, heads, hd, chars, kvf, hd ** -0.5,
                       tok_begin=a, total_tokens=tokens)
        assert torch.equal(part, full[a:b])
        assert bool((buf[:guard] == 7.0).all()) and bool((buf[guard + b - a:] == 7.0).all())


@pytest.mark.parametrize("env", [{"BYA_XA_TC": "0"}, {"BYA_XA_TC": "0", "BYA_XA_VAR": "1"}])
def test_routed_cross_attention_mma_sync_forms(env):
    """The mma.sync forms of the cross-attention (3 characters use them by default; BYA_XA_TC=0 forces them for 1-2
    characters, BYA_XA_VAR=1 stages K / V^T in shared memory) against fp32 torch at the probe's shapes.  The knobs are read
    once per process, hence the subprocess."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_debug_xattn_tc.py")], capture_output=True, text=True,
                       env={**os.environ, **env}, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if "rel max err" in l]
    errs = [float(l.split("rel max err")[1].split()[0]) for l in lines]
    nans = [int(l.split(" nan ")[1].split("/")[0]) for l in lines]
    assert len(errs) == 7 and max(errs) < 1.5e-2 and sum(nans) == 0, r.stdout


def test_router_small_attention_and_head(ops):
    torch.manual_seed(10)
    C, Fr, hw, H = 2, 13, 24, 8
    Nv = Fr * hw
    qkv = rnd(C * Nv, 1536)
    out = torch.zeros(C * Nv, 512, device=dev, dtype=torch.bfloat16)
    # temporal: sequences of the Fr tokens at one (c, hw)
    ops.small_attention(qkv, out, C * hw, Fr, H, hw, Nv, hw)
    x = qkv.float().view(C, Fr, hw, 3, H, 64)
    q, k, v = (x[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))  # [C,hw,H,Fr,64]
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 3, 1, 2, 4).reshape(C * Nv, 512)
    assert rel(out, ref) < TOL
    # multi-ID: sequences of the C tokens at one n
    ops.small_attention(qkv, out, Nv, C, H, Nv, 0, Nv)
    x = qkv.float().view(C, Nv, 3, H, 64)
    q, k, v = (x[:, :, i].permute(1, 2, 0, 3) for i in range(3))  # [Nv,H,C,64]
    ref = F.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(C * Nv, 512)
    assert rel(out, ref) < TOL
    # other sequence lengths: 3 characters, a full 16-row tile, 25 latent frames (97-frame clips: two 16-row tiles)
    for L, hw2 in ((3, 50), (16, 7), (17, 5), (25, 11), (32, 3)):
        n = L * hw2
        qkv2 = rnd(n, 1536)
        out2 = torch.zeros(n, 512, device=dev, dtype=torch.bfloat16)
        ops.small_attention(qkv2, out2, hw2, L, H, hw2, 0, hw2)   # sequences of L rows, hw2 apart
        x = qkv2.float().view(L, hw2, 3, H, 64)
        q, k, v = (x[:, :, i].permute(1, 2, 0, 3) for i in range(3))  # [hw2,H,L,64]
        ref = F.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(n, 512)
        assert rel(out2, ref) < TOL, L
    with pytest.raises(RuntimeError):
        ops.small_attention(rnd(33 * 2, 1536), torch.zeros(66, 512, device=dev, dtype=torch.bfloat16), 2, 33, H, 2, 0, 2)
    xh, w, b = rnd(C * Nv, 512), rnd(512, s=0.1), rnd(1)
    r = torch.empty(Nv, C, device=dev)
    ops.router_head(xh, w, b, r, Nv, C)
    ref = torch.sigmoid(xh.float() @ w.float() + b.float()).view(C, Nv).t()
    assert float((r - ref).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------ prologue kernels (N2)
def test_copy2d_cast_strides_broadcast_and_overlapping_windows(ops):
    torch.manual_seed(31)
    src = torch.randn(37, 200, device=dev)
    out = torch.zeros(37, 256, device=dev, dtype=torch.bfloat16)
    ops.copy2d(src[:, 8:136], out[:, 64:192])                       # fp32 -> bf16, strided both sides, vector path
    assert torch.equal(out[:, 64:192], src[:, 8:136].bfloat16()) and float(out[:, :64].abs().max()) == 0
    ops.copy2d(src[:, 3:10].bfloat16(), out[:, 1:8])                # unaligned, scalar path
    assert torch.equal(out[:, 1:8], src[:, 3:10].bfloat16())
    row = rnd(1, 512)
    big = torch.zeros(5, 3, 512, device=dev, dtype=torch.bfloat16)
    ops.copy2d(row.expand(5, 512), big.view(5, 1536)[:, 512:1024])  # broadcast (row stride 0) into a column block
    assert torch.equal(big[:, 1], row.expand(5, 512)) and float(big[:, 0].abs().max()) == 0
    a = rnd(53 * 96)                                                 # sliding windows of 5 frames, stride 1 frame
    win = torch.as_strided(a, (49, 5 * 96), (96, 1))
    w = torch.empty(49, 480, device=dev, dtype=torch.bfloat16)
    ops.copy2d(win, w)
    assert torch.equal(w, a.view(53, 96).unfold(0, 5, 1).permute(0, 2, 1).reshape(49, 480))


@pytest.mark.parametrize("M,N,K,split,act", [(98, 512, 46080, 148, 3), (48, 24576, 4096, 3, 0), (24, 768, 1024, 5, 0), (130, 256, 640, 64, 1)])
def test_gemm_split_k_is_deterministic_and_matches_fp32(ops, M, N, K, split, act):
    """The skinny weight-streaming GEMMs of the prologue (AudioProjModel.proj1 and the Conv1d(k=2,s=2) taken as a GEMM,
    audio_model.py:78-114): k-splits as independent tiles, per-split fp32 slices, ordered reduction in the finalize."""
    torch.manual_seed(32)
    a, w, b = rnd(M, K, s=0.5), rnd(N, K, s=0.02), rnd(N, s=0.1)
    split = min(split, K // 64)
    ws = torch.full((split, M, N), float("nan"), device=dev)
    out = torch.empty(M + 3, N, device=dev, dtype=torch.bfloat16)
    outs = []
    for _ in range(2):
        ops.gemm(a, w, ws, mode=ops.EPI_SPLITK_F32, split_k=split)
        ops.splitk_finalize(ws, b, act, out[3:])
        outs.append(out[3:].clone())
    assert torch.equal(outs[0], outs[1])
    ref = a.float() @ w.float().t() + b.float()
    ref = F.relu(ref) if act == 3 else F.gelu(ref, approximate="tanh") if act == 1 else ref
    assert rel(outs[0], ref) < TOL
    half = torch.empty(M // 2, N, device=dev, dtype=torch.bfloat16)  # a row range of the workspace
    ops.splitk_finalize(ws, b, act, half, row0=M - M // 2)
    assert torch.equal(half, outs[0][M - M // 2:])


def test_gemm_relu_and_layernorm_leakyrelu(ops):
    torch.manual_seed(33)
    a, w, b = rnd(98, 512, s=0.5), rnd(512, 512, s=0.05), rnd(512, s=0.1)
    out = torch.empty(98, 512, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=b, act=ops.ACT_RELU)
    assert rel(out, F.relu(a.float() @ w.float().t() + b.float())) < TOL
    x, g, be = rnd(1155, 1024, s=2.0) + 0.3, rnd(1024), rnd(1024, s=0.2)
    y = torch.empty_like(x)
    ops.layernorm_leakyrelu(x, y, g, be, eps=1e-5, slope=0.01)
    assert rel(y, F.leaky_relu(F.layer_norm(x.float(), (1024,), g.float(), be.float(), 1e-5), 0.01)) < TOL


@pytest.mark.parametrize("G,H,d", [(2, 16, 128), (26, 48, 64), (3, 5, 64)])
def test_kv_pack_and_router_keys_scatter(ops, G, H, d):
    torch.manual_seed(34)
    x = rnd(G * 32, 2 * H * d + 64)
    K = torch.empty(G, H, 32, d, device=dev, dtype=torch.bfloat16)
    Vt = torch.empty(G, H, d, 32, device=dev, dtype=torch.bfloat16)
    ops.kv_pack(x, 64, 64 + H * d, K, Vt)
    k = x[:, 64:64 + H * d].view(G, 32, H, d).permute(0, 2, 1, 3)
    v = x[:, 64 + H * d:].view(G, 32, H, d).permute(0, 2, 3, 1)
    assert torch.equal(K, k) and torch.equal(Vt, v)
    if d == 128:   # routed keys -> block-structured score matrix (bya_b200.modules.MultiIPRouter.router_keys layout)
        kk = x[:, :H * d]
        mat = torch.full((G * 32 * H, H * d), 7.0, device=dev, dtype=torch.bfloat16)
        ops.router_keys_scatter(kk, mat, G, H, d)
        ref = torch.zeros(G, 32, H, H, d, device=dev, dtype=torch.bfloat16)
        idx = torch.arange(H, device=dev)
        ref[:, :, idx, idx] = kk.reshape(G, 32, H, d)
        assert torch.equal(mat, ref.view(G * 32 * H, H * d))


def test_prologue_on_own_kernels_vs_oracle_and_torch_modules(built):
    """LocalFacialExtractor, AudioProjModel (incl. the 1.2 B-parameter conv as a split-K GEMM) and the K/V
    precompute on libbya.so vs the fp32 oracle (`restated.facial_extractor` / `audio_context`) and vs the torch forward of
    the host-side module mirrors; CFG batch 2 x 2 characters through one batch; deterministic."""
    import bya_b200  # noqa: F401
    from bya_b200.synth import CONFIGS, fill_module, make_inputs
    from bya_b200.transformer import BindyouravatarTransformer3DModel
    from oracle import restated
    import dataclasses

    def cosine(a, b):
        a, b = a.flatten().double(), b.flatten().double()
        return float((a @ b) / (a.norm() * b.norm() + 1e-30))

    cfg = dataclasses.replace(CONFIGS["c1"], num_layers=2, batch=2)
    with torch.device("meta"):
        m = BindyouravatarTransformer3DModel(**cfg.ctor_kwargs())
    m = m.to(torch.bfloat16).to_empty(device="cuda").eval()
    m.router.frames, m.router.height, m.router.width = cfg.frames, cfg.grid_w, cfg.grid_h
    m.router.pos_emb = m.router._create_positional_embedding().to("cuda", torch.bfloat16)
    fill_module(m, 0)
    inp = make_inputs(cfg, 77, device="cuda", dtype=torch.bfloat16)
    eng = m.engine()
    pro = eng.prologue(inp["id_cond"], inp["id_vit_hidden"], inp["audio_embeds"], cfg.frames, True)
    snap = {k: (v.clone() if torch.is_tensor(v) else [[t.clone() for t in l] for l in v]) for k, v in pro.items()}
    sd = {k: v.float() for k, v in m.state_dict().items()}
    B, C = 2, cfg.chars
    face_ref = torch.stack([restated.facial_extractor(sd, inp["id_cond"][c].float(), [v.float() for v in inp["id_vit_hidden"][c]])
                            for c in range(C)], 1)
    assert tuple(pro["face_tokens"].shape) == (B, C, 32, 2048)
    assert cosine(pro["face_tokens"], face_ref) >= 0.9995
    assert rel(pro["face_tokens"], face_ref) < 0.03
    a = inp["audio_embeds"].float()
    ctx_ref = restated.audio_context(sd, a.reshape(B * C, *a.shape[2:]), cfg.frames).reshape(B, C, cfg.frames, 32, -1)
    assert cosine(pro["audio_ctx"], ctx_ref) >= 0.9995 and rel(pro["audio_ctx"], ctx_ref) < 0.03
    # K / V^T / routed keys: the torch forward of the module mirrors on the SAME (kernel-produced) tokens
    for b in range(B):
        for j in (0, len(m.perceiver_cross_attention) - 1):
            k, v = m.perceiver_cross_attention[j].face_kv(pro["face_tokens"][b])
            assert rel(pro["face_k"][b][j], k) < TOL and rel(pro["face_vt"][b][j], v.transpose(-1, -2)) < TOL
            assert rel(pro["kmat"][b][j], m.router.router_keys(k, j)) < 1.5e-2
        for l in (0, len(m.audio_model.layers) - 1):
            k, vt = m.audio_model.audio_kv(pro["audio_ctx"][b], l)
            assert rel(pro["aud_k"][b][l], k) < TOL and rel(pro["aud_vt"][b][l], vt) < TOL
    again = eng.prologue(inp["id_cond"], inp["id_vit_hidden"], inp["audio_embeds"], cfg.frames, True)
    assert torch.equal(again["face_tokens"], snap["face_tokens"]) and torch.equal(again["audio_ctx"], snap["audio_ctx"])
    assert all(torch.equal(x, y) for x, y in zip(again["aud_k"][1], snap["aud_k"][1]))


# ------------------------------------------------------------------------------------------------ bit-exact mask path
def _c_oracle(masks, Fr, gh, gw, frame_or=False):
    import ctypes

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libmask_oracle.so"))
    C, T, H, W = masks.shape
    n = Fr * gh * gw
    idx, lg = np.zeros(n, np.int64), np.zeros((n, C), np.float32)
    m = np.ascontiguousarray(masks, np.uint8)
    lib.bya_oracle_masks_to_routing(m.ctypes.data_as(ctypes.c_void_p), C, T, H, W, Fr, gh, gw,
                                    idx.ctypes.data_as(ctypes.c_void_p), lg.ctypes.data_as(ctypes.c_void_p), int(frame_or))
    return idx, lg


@pytest.mark.parametrize("kind", ["moving", "static", "overlap", "speckle"])
def test_masks_to_routing_bit_exact_vs_reference_golden(ops, kind):
    import bya_b200  # noqa: F401
    from bya_b200.synth import tracking_masks

    g = torch.load(os.path.join(GOLD, f"masks_{kind}.pt"))
    masks = tracking_masks(kind)
    idx, lg = ops.masks_to_routing(torch.from_numpy(masks).to(dev), 13, 30, 45)
    assert torch.equal(lg.cpu(), g["routing_logits"][0].float())  # reference's own output, bit for bit
    ci, cl = _c_oracle(masks, 13, 30, 45)
    assert np.array_equal(idx.cpu().numpy(), ci) and np.array_equal(lg.cpu().numpy(), cl)
    ored = ops.routing_frame_or(lg, torch.empty_like(lg), 13)
    assert np.array_equal(ored.cpu().numpy(), _c_oracle(masks, 13, 30, 45, frame_or=True)[1])


@pytest.mark.parametrize("kind", ["moving", "overlap"])
def test_mask_png_directories_to_routing_logits(ops, kind, tmp_path):
    """The file-level entry (`bya_b200.masks.process_masks_to_routing_logits`, reference util/utils.py:871-936): PNG
    directories written the way the SAM-2 stage does -> logits equal to the reference's output on the same files."""
    from PIL import Image

    from bya_b200 import masks as bm
    from bya_b200.synth import tracking_masks

    m = tracking_masks(kind)
    for c in range(2):
        d = tmp_path / str(c + 1)
        d.mkdir()
        for t in range(m.shape[1]):
            Image.fromarray(m[c, t] * (200 if c else 1)).save(str(d / f"annotated_frame_{t:05d}.png"))   # any value > 0 counts
    lg = bm.process_masks_to_routing_logits(str(tmp_path))
    g = torch.load(os.path.join(GOLD, f"masks_{kind}.pt"))
    assert lg.shape == (1, 17550, 2) and lg.dtype == torch.float32
    assert torch.equal(lg.cpu(), g["routing_logits"].float())
    (tmp_path / "2").rename(tmp_path / "two")
    with pytest.raises(ValueError):
        bm.process_masks_to_routing_logits(str(tmp_path))


@pytest.mark.parametrize("geom", [(3, 97, 200, 300, 25, 10, 15), (2, 49, 480, 720, 13, 30, 45), (1, 13, 30, 45, 13, 30, 45),
                                  (3, 10, 33, 47, 13, 8, 12), (2, 1, 4, 4, 1, 1, 1)])
def test_masks_to_routing_bit_exact_random_geometries(ops, geom):
    """3 characters, 97 frames, non-integer scales, identity and degenerate sizes; random speckle maximises ties."""
    C, T, H, W, Fr, gh, gw = geom
    rng = np.random.RandomState(11)
    masks = (rng.rand(C, T, H, W) > 0.5).astype(np.uint8) * rng.randint(1, 255, size=(C, 1, 1, 1)).astype(np.uint8)
    idx, lg = ops.masks_to_routing(torch.from_numpy(masks).to(dev), Fr, gh, gw)
    ci, cl = _c_oracle(masks, Fr, gh, gw)
    assert np.array_equal(idx.cpu().numpy(), ci) and np.array_equal(lg.cpu().numpy(), cl)
    # empty masks -> all background
    idx0, lg0 = ops.masks_to_routing(torch.zeros_like(torch.from_numpy(masks)).to(dev), Fr, gh, gw)
    assert int((idx0 != -1).sum()) == 0 and float(lg0.abs().sum()) == 0.0


def test_audio_weights_exact_for_hard_masks(ops):
    from oracle import restated

    torch.manual_seed(12)
    for C in (2, 3):
        lab = torch.randint(-1, C, (500,))
        r = torch.zeros(500, C)
        for c in range(C):
            r[lab == c, c] = 1
        for af in (torch.eye(C), (1 - torch.eye(C)) if C == 2 else torch.eye(C)[[1, 2, 0]]):
            w = torch.empty(500, C, device=dev)
            ws = torch.empty(500, device=dev)
            ops.audio_weights(af.to(dev).contiguous(), r.to(dev), w, ws)
            ref = restated.audio_weights(af, r)
            assert torch.equal(w.cpu(), ref) and torch.equal(ws.cpu(), ref.sum(1))
    r = torch.rand(300, 2)
    w = torch.empty(300, 2, device=dev)
    ops.audio_weights(torch.eye(2, device=dev), r.to(dev), w)
    assert float((w.cpu() - (1 - r[:, [1, 0]])).abs().max()) < 1e-6
