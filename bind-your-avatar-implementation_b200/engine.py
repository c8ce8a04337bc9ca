"""The denoising-step engine: packs the model's weights once for the sm_100a kernels and runs one step.

What one step launches (per CFG batch element; SURVEY.md §3.3/§3.4, reference models/transformer.py:615-964):

  temb          timestep_features -> gemv(linear_1, SiLU) -> gemv(linear_2) -> ONE gemv over all 2L+1 adaLN linears
  embed         text GEMM; patchify -> GEMM (the k=2,s=2 conv as [Nv,192]x[192,D])
  per layer     LN+modulate -> QKV GEMM (bias + qk-LayerNorm + RoPE epilogue) -> tcgen05 flash attention ->
                out GEMM (gate*x + residual epilogue) -> LN+modulate -> FFN1 GEMM (GELU epilogue) -> FFN2 GEMM (gated
                residual epilogue)
  face layers   LN -> to_q GEMM (once, not per character: §0.11) -> [router | forced masks] -> routed 32-key
                cross-attention that blends the characters on the fly -> ONE to_out GEMM (scale*x + residual)
  router        LN(perm-folded) -> to_q GEMM -> score GEMM against block-structured keys -> LN+pos-emb ->
                4 x {spatial FA, temporal attn, multi-ID attn, MLP} -> head
  audio layers  audio weights -> LN -> to_q GEMM -> routed per-frame 32-key cross-attention -> to_out GEMM
  head          LN -> LN+modulate -> proj_out GEMM -> unpatchify

Everything timestep-invariant (face tokens, face K/V, router keys, audio context, audio K/V) is computed once per
generation in `prologue()` with torch on the GPU and cached.
There is no fallback path: every op above is a libbya.so kernel.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import ops


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def router_feature_perm(heads: int = 16, dh: int = 128, device=None) -> torch.Tensor:
    """f(g): position in the reference's router feature order (d*heads + h) of natural feature g = h*dh + d
    (router.py:375-378)."""
    g = torch.arange(heads * dh, device=device)
    return (g % dh) * heads + g // dh


class RouterPack:
    """Router weights in kernel layout (input permutation folded into norm_q / to_q, QKV fused per attention)."""

    def __init__(self, router):
        dev = router.norm.weight.device
        perm = router_feature_perm(router.heads, 2048 // router.heads, dev)
        self.eps = router.norm.eps
        self.nq_w, self.nq_b = _bf(router.norm_q.weight[perm]), _bf(router.norm_q.bias[perm])
        self.to_q = [_bf(l.weight[:, perm]) for l in router.to_q]
        self.n_w, self.n_b = _bf(router.norm.weight), _bf(router.norm.bias)
        self.pos = _bf(router.pos_emb.reshape(-1, router.pos_emb.shape[-1]))
        self.blocks = []
        for blk in router.spatial_temporal_layers:
            d = {}
            for name, attn in (("s", blk.spatial_attn), ("t", blk.temporal_attn), ("i", blk.multi_id_attn)):
                d[f"{name}_qkv_w"] = _bf(torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight], 0))
                d[f"{name}_qkv_b"] = _bf(torch.cat([attn.to_q.bias, attn.to_k.bias, attn.to_v.bias], 0))
                d[f"{name}_o_w"], d[f"{name}_o_b"] = _bf(attn.to_out[0].weight), _bf(attn.to_out[0].bias)
            for k, n in (("n1", blk.norm1), ("n2", blk.norm2), ("n3", blk.norm3), ("n4", blk.norm4)):
                d[k] = (_bf(n.weight), _bf(n.bias), n.eps)
            d["m0_w"], d["m0_b"] = _bf(blk.mlp[0].weight), _bf(blk.mlp[0].bias)
            d["m2_w"], d["m2_b"] = _bf(blk.mlp[2].weight), _bf(blk.mlp[2].bias)
            # LayerNorm folded into the linear that follows it, for the fused links (`ops.gemm_ln_gemm`):
            # f_s = norm1 -> spatial qkv, f_t = norm2 -> temporal qkv, f_i = norm3 -> multi-ID qkv, f_m = norm4 -> mlp[0]
            for k, n, w, b in (("f_s", "n1", "s_qkv_w", "s_qkv_b"), ("f_t", "n2", "t_qkv_w", "t_qkv_b"),
                               ("f_i", "n3", "i_qkv_w", "i_qkv_b"), ("f_m", "n4", "m0_w", "m0_b")):
                d[k] = ops.fold_layernorm(d[w], d[b], d[n][0], d[n][1])
            self.blocks.append(d)
        self.head_w, self.head_b = _bf(router.final_proj[0].weight.reshape(-1)), _bf(router.final_proj[0].bias)


def poison_scratch() -> bool:
    """BYA_POISON_SCRATCH=1: every scratch / exchange buffer is NaN-filled when it is allocated.  A kernel that consumes
    rows nobody wrote (even with weight 0: 0 x NaN = NaN, while 0 x finite garbage is exactly 0 and goes unnoticed) then
    shows up as NaNs in the result — the parity tests and bench.py's sp_check run with it."""
    return os.environ.get("BYA_POISON_SCRATCH", "0") == "1"


class _Workspace:
    """Named scratch buffers, allocated once per geometry (the C ABI never allocates)."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}

    def get(self, name, shape, dtype=torch.bfloat16):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            if poison_scratch() and t.dtype.is_floating_point:
                t.fill_(float("nan"))
            self.bufs[name] = t
        return t

    def get_zeroed(self, name, shape, dtype=torch.bfloat16):
        """Like get(), zero-filled (a memset node) when it is (re)allocated — split-K accumulators, which their
        finalize kernel leaves clean, and operand regions that are never written."""
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = ops.memset_zero(self.get(name, shape, dtype))
        return t


def _tree_values(x):
    """Nested dict / list of tensors -> nested lists in a deterministic order."""
    if isinstance(x, dict):
        return [_tree_values(x[k]) for k in sorted(x)]
    if isinstance(x, (list, tuple)):
        return [_tree_values(y) for y in x]
    return x


def run_router(rp: RouterPack, ws: _Workspace, qf: torch.Tensor, kmat: torch.Tensor, layer: int, chars: int, frames: int,
               hw: int, out: torch.Tensor, fused: bool = True) -> torch.Tensor:
    """qf [Nv,2048] (natural head-major face queries) , kmat [C*512,2048] -> out [Nv,C] fp32 soft routing.
    MultiIPRouter.forward (router.py:364-411) + SpatialTemporalAttentionBlock.forward (:468-493)."""
    Nv = qf.shape[0]
    C, M = chars, chars * Nv
    rq = ws.get("r_q", (Nv, 2048))
    ops.layernorm_modulate(qf, rq, eps=rp.eps, gamma=rp.nq_w, beta=rp.nq_b)
    rq2 = ws.get("r_q2", (Nv, 2048))
    ops.gemm(rq, rp.to_q[layer], rq2)
    sc = ws.get("r_scores", (Nv, C * 512))
    ops.gemm(rq2, kmat, sc)
    x = ws.get("r_x", (M, 512))
    for c in range(C):
        ops.layernorm_modulate(sc[:, c * 512:(c + 1) * 512], x[c * Nv:(c + 1) * Nv], eps=rp.eps, gamma=rp.n_w,
                               beta=rp.n_b, add=rp.pos)
    xn = ws.get("r_xn", (M, 512))
    qkv = ws.get("r_qkv", (M, 1536))
    att = ws.get("r_att", (M, 512))
    if fused:
        # every `x += proj(...)` + next LayerNorm + next projection is one kernel (`bya_gemm_ln_gemm_bf16`)
        nb = len(rp.blocks)
        for bi, d in enumerate(rp.blocks):
            if bi == 0:
                ops.layernorm_modulate(x, xn, eps=d["n1"][2], gamma=d["n1"][0], beta=d["n1"][1])
                ops.gemm(xn, d["s_qkv_w"], qkv, bias=d["s_qkv_b"])
            ops.attention_d64(qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:], att, C * frames, hw, 8)
            ops.gemm_ln_gemm(att, d["s_o_w"], d["s_o_b"], x, x, *d["f_t"], qkv, ln_eps=d["n2"][2])
            ops.small_attention(qkv, att, C * hw, frames, 8, hw, Nv, hw)
            ops.gemm_ln_gemm(att, d["t_o_w"], d["t_o_b"], x, x, *d["f_i"], qkv, ln_eps=d["n3"][2])
            ops.small_attention(qkv, att, Nv, C, 8, Nv, 0, Nv)
            ops.gemm_ln_gemm(att, d["i_o_w"], d["i_o_b"], x, x, *d["f_m"], xn, ln_eps=d["n4"][2], act=ops.ACT_GELU_ERF)
            if bi + 1 < nb:
                dn = rp.blocks[bi + 1]
                ops.gemm_ln_gemm(xn, d["m2_w"], d["m2_b"], x, x, *dn["f_s"], qkv, ln_eps=dn["n1"][2])
            else:
                ops.gemm(xn, d["m2_w"], x, bias=d["m2_b"], mode=ops.EPI_RESIDUAL, resid=x)
        ops.router_head(x, rp.head_w, rp.head_b, out, Nv, C)
        return out
    for d in rp.blocks:
        # spatial: all H*W tokens of one (character, frame)
        ops.layernorm_modulate(x, xn, eps=d["n1"][2], gamma=d["n1"][0], beta=d["n1"][1])
        ops.gemm(xn, d["s_qkv_w"], qkv, bias=d["s_qkv_b"])
        ops.attention_d64(qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:], att, C * frames, hw, 8)
        ops.gemm(att, d["s_o_w"], x, bias=d["s_o_b"], mode=ops.EPI_RESIDUAL, resid=x)
        # temporal: the F tokens at one (character, h, w)
        ops.layernorm_modulate(x, xn, eps=d["n2"][2], gamma=d["n2"][0], beta=d["n2"][1])
        ops.gemm(xn, d["t_qkv_w"], qkv, bias=d["t_qkv_b"])
        ops.small_attention(qkv, att, C * hw, frames, 8, hw, Nv, hw)
        ops.gemm(att, d["t_o_w"], x, bias=d["t_o_b"], mode=ops.EPI_RESIDUAL, resid=x)
        # multi-ID: the C tokens at one (t, h, w)
        ops.layernorm_modulate(x, xn, eps=d["n3"][2], gamma=d["n3"][0], beta=d["n3"][1])
        ops.gemm(xn, d["i_qkv_w"], qkv, bias=d["i_qkv_b"])
        ops.small_attention(qkv, att, Nv, C, 8, Nv, 0, Nv)
        ops.gemm(att, d["i_o_w"], x, bias=d["i_o_b"], mode=ops.EPI_RESIDUAL, resid=x)
        # MLP (exact-erf GELU, ratio 1)
        ops.layernorm_modulate(x, xn, eps=d["n4"][2], gamma=d["n4"][0], beta=d["n4"][1])
        ops.gemm(xn, d["m0_w"], att, bias=d["m0_b"], act=ops.ACT_GELU_ERF)
        ops.gemm(att, d["m2_w"], x, bias=d["m2_b"], mode=ops.EPI_RESIDUAL, resid=x)
    ops.router_head(x, rp.head_w, rp.head_b, out, Nv, C)
    return out


class RouterShard:
    """Sequence-parallel layout of the router: rank r owns the spatial positions
    [r*hwl, (r+1)*hwl) of EVERY (character, frame) — hwl = ceil(hw / P), the last rank's slice is padded with copies of
    the last real position — so the temporal, multi-ID and row-local parts are local, and only the spatial attention
    (all positions of one frame) needs an exchange: Ulysses over its 8 heads, q|k|v out / attention output back."""

    def __init__(self, rp: RouterPack, frames: int, hw: int, world: int, rank: int, device):
        from .sp import qkv_rows_by_destination, router_local_tokens

        if 8 % world:
            raise RuntimeError(f"bya_b200: the router's 8 heads are not divisible by the sequence-parallel size {world}")
        self.P, self.rank, self.frames, self.hw = world, rank, frames, hw
        self.hwl = (hw + world - 1) // world
        self.hw_pad = self.hwl * world
        self.hl = 8 // world                      # spatial-attention heads per rank
        self.idx = router_local_tokens(frames, hw, world, rank, device)    # [F*hwl] global token of each local row
        self.pos = rp.pos.index_select(0, self.idx).contiguous()
        self.blocks = [dict(w=qkv_rows_by_destination(d["s_qkv_w"], 512, world),
                            b=qkv_rows_by_destination(d["s_qkv_b"], 512, world),
                            f_s=tuple(qkv_rows_by_destination(t, 512, world) for t in d["f_s"])) for d in rp.blocks]


def run_router_sp(rp: RouterPack, rs: RouterShard, ws: _Workspace, q_all, kmat: torch.Tensor, layer: int,
                  chars: int, out: torch.Tensor, group, peer=None, q_local=None, text_len: int = 0, rows_per_rank: int = 0,
                  fused: bool = True) -> torch.Tensor:
    """`run_router` sharded over the sequence-parallel group: q_all [Nv,2048] (every rank holds all face queries; None with
    the peer exchange, where every owner pushes the rows of `q_local` [rows_per_rank, 2048] the other ranks' routers need)
    -> out [Nv,C] fp32 on every rank.  Same kernels, same per-row arithmetic as the single-GPU router.  Exchanges: NCCL
    all-to-all + permuting copies (peer=None), or one strided push into the peers' buffers + one device barrier each
    (`peer.py`; the position gather / scatter permutations are folded into the segment strides)."""
    import torch.distributed as dist

    from . import peer as pk
    from .sp import router_gather_positions, router_scatter_positions

    C, Fr, P, hwl, hl = chars, rs.frames, rs.P, rs.hwl, rs.hl
    R = Fr * hwl                 # local (frame, position) rows
    M = C * R                    # local router rows, ordered (character, frame, local position)
    CF = C * Fr
    Ws, Wo = 3 * hl * 64, hl * 64
    if peer is None:
        qf = ws.get("rs_qsel", (R, 2048))
        torch.index_select(q_all, 0, rs.idx, out=qf)
        s_full = ws.get("rs_s_full", (CF * rs.hw_pad, Ws))   # [(c,f)][all positions (padded)][my heads]
        o_recv = ws.get("rs_o_recv", (P, M, Wo))          # [src heads][local row] = K-blocked A of the out-projection
        s_recv = ws.get("rs_s_recv", (P, M, Ws))          # [src][(c,f), src's positions][my heads]
        o_send = ws.get("rs_o_send", (P, M, Wo))
    else:
        qf, _, qf_ptrs = peer.get("rs_qsel", (R, 2048))
        s_full, _, s_full_ptrs = peer.get("rs_s_full", (CF * rs.hw_pad, Ws))
        o_recv, _, o_recv_ptrs = peer.get("rs_o_recv", (P, M, Wo))
        peer.push(("face_q", Fr, rs.hw, text_len, rows_per_rank), lambda d: pk.face_query_segments(
            P, d, Fr, rs.hw, text_len, rows_per_rank, 2048), q_local, qf_ptrs)
        peer.barrier()           # the queries of my router positions have landed
    s_send = ws.get("rs_s_send", (P, M, Ws))              # [dest][local row][q|k|v heads of dest]
    # zeroed once: the attention writes the hw real positions of every (c,f); the padding rows of the last shard are
    # exchanged and projected like any row, and feed the NEXT block's K/V rows hw.., which the attention kernel loads
    # with its last key tile (probability 0 — but 0 x NaN garbage is NaN)
    s_att = ws.get_zeroed("rs_s_att", (CF * rs.hw_pad, Wo))
    rq = ws.get("rs_q", (R, 2048))
    ops.layernorm_modulate(qf, rq, eps=rp.eps, gamma=rp.nq_w, beta=rp.nq_b)
    rq2 = ws.get("rs_q2", (R, 2048))
    ops.gemm(rq, rp.to_q[layer], rq2)
    sc = ws.get("rs_scores", (R, C * 512))
    ops.gemm(rq2, kmat, sc)
    x = ws.get("rs_x", (M, 512))
    for c in range(C):
        ops.layernorm_modulate(sc[:, c * 512:(c + 1) * 512], x[c * R:(c + 1) * R], eps=rp.eps, gamma=rp.n_w, beta=rp.n_b,
                               add=rs.pos)
    xn = ws.get("rs_xn", (M, 512))
    qkv = ws.get("rs_qkv", (M, 1536))
    att = ws.get("rs_att", (M, 512))
    x2 = ws.get("rs_x2", (M, 512)) if fused else None     # fused links with column slices write X out of place
    nb = len(rp.blocks)
    for bi, (d, dsp) in enumerate(zip(rp.blocks, rs.blocks)):
        # spatial: all H*W tokens of one (character, frame) — heads sharded, positions gathered
        if not fused or bi == 0:
            ops.layernorm_modulate(x, xn, eps=d["n1"][2], gamma=d["n1"][0], beta=d["n1"][1])
            ops.gemm(xn, dsp["w"], s_send[0], bias=dsp["b"], col_block=Ws, col_block_stride=M * Ws)
        if peer is None:
            dist.all_to_all_single(s_recv, s_send, group=group)
            s_full.view(CF, P, hwl, Ws).copy_(router_gather_positions(s_recv, CF, P, hwl))
        else:
            peer.push(("rs_gather", CF, hwl, M, Ws), lambda r: pk.router_gather_segments(P, r, CF, hwl, M, Ws), s_send, s_full_ptrs)
            peer.barrier()
        ops.attention_d64(s_full[:, :Wo], s_full[:, Wo:2 * Wo], s_full[:, 2 * Wo:], s_att, CF, rs.hw, hl,
                          seq_stride=rs.hw_pad)
        if peer is None:
            o_send.view(P, CF, hwl, Wo).copy_(router_scatter_positions(s_att, CF, P, hwl))
            dist.all_to_all_single(o_recv, o_send, group=group)
        else:
            peer.push(("rs_scatter", CF, hwl, M, Wo), lambda r: pk.router_scatter_segments(P, r, CF, hwl, M, Wo), s_att, o_recv_ptrs)
            peer.barrier()
        if fused:
            # `x += proj(...)` + next LayerNorm + next projection as one kernel each (`bya_gemm_ln_gemm_bf16`)
            ops.gemm_ln_gemm(o_recv[0], d["s_o_w"], d["s_o_b"], x, x2, *d["f_t"], qkv, ln_eps=d["n2"][2],
                             n_split=ops.chain_n_split(M, 1536), a_kblock=Wo, a_kblock_stride=M * Wo)
            x, x2 = x2, x
            ops.small_attention(qkv, att, C * hwl, Fr, 8, hwl, R, hwl)
            ops.gemm_ln_gemm(att, d["t_o_w"], d["t_o_b"], x, x2, *d["f_i"], qkv, ln_eps=d["n3"][2],
                             n_split=ops.chain_n_split(M, 1536))
            x, x2 = x2, x
            ops.small_attention(qkv, att, R, C, 8, R, 0, R)
            ops.gemm_ln_gemm(att, d["i_o_w"], d["i_o_b"], x, x2, *d["f_m"], xn, ln_eps=d["n4"][2], act=ops.ACT_GELU_ERF,
                             n_split=ops.chain_n_split(M, 512))
            x, x2 = x2, x
            if bi + 1 < nb:
                dn, dspn = rp.blocks[bi + 1], rs.blocks[bi + 1]
                ops.gemm_ln_gemm(xn, d["m2_w"], d["m2_b"], x, x2, *dspn["f_s"], s_send[0], ln_eps=dn["n1"][2],
                                 n_split=ops.chain_n_split(M, 1536), col_block=Ws, col_block_stride=M * Ws)
                x, x2 = x2, x
            else:
                ops.gemm(xn, d["m2_w"], x, bias=d["m2_b"], mode=ops.EPI_RESIDUAL, resid=x)
            continue
        ops.gemm(o_recv[0], d["s_o_w"], x, bias=d["s_o_b"], mode=ops.EPI_RESIDUAL, resid=x, a_kblock=Wo,
                 a_kblock_stride=M * Wo)
        # temporal: the F tokens at one (character, local position)
        ops.layernorm_modulate(x, xn, eps=d["n2"][2], gamma=d["n2"][0], beta=d["n2"][1])
        ops.gemm(xn, d["t_qkv_w"], qkv, bias=d["t_qkv_b"])
        ops.small_attention(qkv, att, C * hwl, Fr, 8, hwl, R, hwl)
        ops.gemm(att, d["t_o_w"], x, bias=d["t_o_b"], mode=ops.EPI_RESIDUAL, resid=x)
        # multi-ID: the C tokens at one (frame, local position)
        ops.layernorm_modulate(x, xn, eps=d["n3"][2], gamma=d["n3"][0], beta=d["n3"][1])
        ops.gemm(xn, d["i_qkv_w"], qkv, bias=d["i_qkv_b"])
        ops.small_attention(qkv, att, R, C, 8, R, 0, R)
        ops.gemm(att, d["i_o_w"], x, bias=d["i_o_b"], mode=ops.EPI_RESIDUAL, resid=x)
        # MLP (exact-erf GELU, ratio 1)
        ops.layernorm_modulate(x, xn, eps=d["n4"][2], gamma=d["n4"][0], beta=d["n4"][1])
        ops.gemm(xn, d["m0_w"], att, bias=d["m0_b"], act=ops.ACT_GELU_ERF)
        ops.gemm(att, d["m2_w"], x, bias=d["m2_b"], mode=ops.EPI_RESIDUAL, resid=x)
    r_loc = ws.get("rs_r_loc", (R, C), torch.float32)
    ops.router_head(x, rp.head_w, rp.head_b, r_loc, R, C)
    if peer is None:
        r_all = ws.get("rs_r_all", (P, Fr, hwl, C), torch.float32)
        dist.all_gather_into_tensor(r_all, r_loc, group=group)
        out.view(Fr, rs.hw, C).copy_(r_all.permute(1, 0, 2, 3).reshape(Fr, rs.hw_pad, C)[:, :rs.hw])
    else:   # `out` is the symmetric routing buffer: every rank writes its positions into everyone's copy
        out_ptrs = peer.get("routing", tuple(out.shape), torch.float32)[2]
        peer.push(("rs_routing", Fr, rs.hw, C), lambda r: pk.routing_gather_segments(P, Fr, rs.hw, C), r_loc, out_ptrs)
        peer.barrier()
    return out


def router_forward_standalone(router, q_out, k_out, layer_idx):
    """Module-level entry (`MultiIPRouter.forward` signature): q_out [C,16,Nv,128] -> [1,Nv,C]."""
    C = k_out.shape[0]
    Nv = q_out.shape[2]
    rp = RouterPack(router)
    ws = _Workspace(q_out.device)
    qf = q_out[0].transpose(0, 1).reshape(Nv, -1).to(torch.bfloat16).contiguous()  # q is identical for every character
    kmat = _bf(router.router_keys(k_out.to(router.norm_k.weight.dtype), layer_idx))
    out = torch.empty(Nv, C, device=q_out.device, dtype=torch.float32)
    hw = router.height * router.width
    run_router(rp, ws, qf, kmat, layer_idx, C, router.frames, hw, out)
    return out[None].to(q_out.dtype)


class StepEngine:
    def __init__(self, model):
        self.model = model
        cfg = model.config
        p = next(model.transformer_blocks.parameters())
        if p.device.type != "cuda" or p.dtype != torch.bfloat16:
            raise RuntimeError("bya_b200: the denoiser must live on a CUDA device in bfloat16 "
                               "(`model.to('cuda', torch.bfloat16)`); there is no CPU path")
        ops.lib()  # fail loudly if libbya.so is missing
        self.device = p.device
        self.D = cfg.num_attention_heads * cfg.attention_head_dim
        self.heads = cfg.num_attention_heads
        if cfg.attention_head_dim != 64:
            raise RuntimeError("bya_b200: attention_head_dim must be 64")
        if cfg.patch_size != 2:
            raise RuntimeError("bya_b200: patchify / unpatchify kernels are built for patch_size == 2")
        self.L = cfg.num_layers
        self.ws = _Workspace(self.device)
        self._pack()
        self._prologue_key = None
        self._prologue = None
        self._graphs: Dict[tuple, dict] = {}
        self._router_shard = None

    # ------------------------------------------------------------------------------------------------ packing
    def _pack(self):
        m, D = self.model, self.D
        self.layers = []
        ada_w, ada_b = [], []
        # softmax scale * log2(e), folded into q by the QKV epilogue when the bounded attention kernel is used
        hd = m.config.attention_head_dim
        self.q_premul = hd ** -0.5 * math.log2(math.e)
        for blk in m.transformer_blocks:
            a = blk.attn1
            # |q.k| bound from the per-head LayerNorm (|LN(x)|_2 <= sqrt(64); RoPE is a rotation): lets the joint
            # self-attention run without a running max (bya_attention_d64_bounded); 2 % margin for bf16 rounding
            qn = math.sqrt(hd) * float(a.norm_q.weight.float().abs().max()) + float(a.norm_q.bias.float().norm())
            kn = math.sqrt(hd) * float(a.norm_k.weight.float().abs().max()) + float(a.norm_k.bias.float().norm())
            bound = 1.02 * qn * kn * self.q_premul + 1e-3
            self.layers.append(dict(
                score_bound=bound if (bound <= 64.0 and getattr(m, "bounded_attention", True)) else None,
                w_qkv=_bf(torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight], 0)),
                b_qkv=_bf(torch.cat([a.to_q.bias, a.to_k.bias, a.to_v.bias], 0)) if a.to_q.bias is not None else None,
                nq=(_bf(a.norm_q.weight), _bf(a.norm_q.bias)), nk=(_bf(a.norm_k.weight), _bf(a.norm_k.bias)),
                qk_eps=a.norm_q.eps,
                w_o=_bf(a.to_out[0].weight), b_o=_bf(a.to_out[0].bias),
                w_f1=_bf(blk.ff.net[0].proj.weight), b_f1=_bf(blk.ff.net[0].proj.bias),
                w_f2=_bf(blk.ff.net[2].weight), b_f2=_bf(blk.ff.net[2].bias),
                ln1=(_bf(blk.norm1.norm.weight), _bf(blk.norm1.norm.bias), blk.norm1.norm.eps),
                ln2=(_bf(blk.norm2.norm.weight), _bf(blk.norm2.norm.bias), blk.norm2.norm.eps),
            ))
            for n in (blk.norm1, blk.norm2):
                ada_w.append(n.linear.weight)
                ada_b.append(n.linear.bias)
        ada_w.append(m.norm_out.linear.weight)
        ada_b.append(m.norm_out.linear.bias)
        self.ada_w, self.ada_b = _bf(torch.cat(ada_w, 0)), _bf(torch.cat(ada_b, 0))
        self.te = (_bf(m.time_embedding.linear_1.weight), _bf(m.time_embedding.linear_1.bias),
                   _bf(m.time_embedding.linear_2.weight), _bf(m.time_embedding.linear_2.bias))
        pw = m.patch_embed.proj.weight
        kp = pw[0].numel()
        self.patch_k = (kp + 63) // 64 * 64
        w = torch.zeros(D, self.patch_k, device=self.device, dtype=torch.bfloat16)
        w[:, :kp] = pw.reshape(D, kp)
        self.patch_w, self.patch_b = w, _bf(m.patch_embed.proj.bias)
        self.text_w, self.text_b = _bf(m.patch_embed.text_proj.weight), _bf(m.patch_embed.text_proj.bias)
        self.nf = (_bf(m.norm_final.weight), _bf(m.norm_final.bias), m.norm_final.eps)
        self.no = (_bf(m.norm_out.norm.weight), _bf(m.norm_out.norm.bias), m.norm_out.norm.eps)
        po = m.proj_out.weight
        self.out_cols = po.shape[0]
        n_pad = (self.out_cols + 63) // 64 * 64
        self.proj_w = torch.zeros(n_pad, D, device=self.device, dtype=torch.bfloat16)
        self.proj_w[: self.out_cols] = po
        self.proj_b = torch.zeros(n_pad, device=self.device, dtype=torch.bfloat16)
        self.proj_b[: self.out_cols] = m.proj_out.bias
        self.face = []
        if m.is_train_face:
            for ca in m.perceiver_cross_attention:
                self.face.append(dict(ln=(_bf(ca.norm2.weight), _bf(ca.norm2.bias), ca.norm2.eps), w_q=_bf(ca.to_q.weight),
                                      w_o=_bf(ca.to_out.weight)))
            self.router = RouterPack(m.router)
            from .prologue import ProloguePack

            self.pro_pack = ProloguePack(m, router_feature_perm(m.router.heads, 2048 // m.router.heads, self.device))
        self.audio = []
        if getattr(m, "is_train_audio", False):
            for lyr in m.audio_model.layers:
                a = lyr["attn"]
                self.audio.append(dict(ln=(_bf(lyr["norm_q"].weight), _bf(lyr["norm_q"].bias), lyr["norm_q"].eps),
                                       w_q=_bf(a.to_q.weight), b_q=_bf(a.to_q.bias), w_o=_bf(a.to_out[0].weight),
                                       b_o=_bf(a.to_out[0].bias)))

    # ------------------------------------------------------------------------------------------------ prologue
    @torch.no_grad()
    def prologue(self, id_cond, id_vit_hidden, audio_embeds, frames, use_router: bool, ws: Optional[_Workspace] = None):
        """Timestep-invariant tensors of one generation on libbya.so kernels (`prologue.py`).
        The results live in `ws` (named buffers): a holder that caches them (the eager per-generation cache, every
        captured graph with its prologue outside the graph) passes a workspace of its own, so that recomputing for one
        holder never rewrites what another one still reads."""
        from . import prologue as pk

        if not hasattr(self, "pro_pack"):
            raise RuntimeError("bya_b200: the model was built without the face branch (is_train_face=False)")
        P, dev = self.pro_pack, self.device
        ws = ws if ws is not None else self.ws
        C, B = len(id_cond), id_cond[0].shape[0]
        face = pk.face_tokens(P, ws, id_cond, id_vit_hidden, dev)
        out = pk.face_kv_and_router_keys(P, ws, face, use_router)
        out["face_tokens"] = face                                  # [B,C,32,2048]
        out["aud_k"], out["aud_vt"] = [], []
        if audio_embeds is not None and P.audio is not None:
            a = audio_embeds.to(dev, torch.bfloat16)
            if a.ndim != 5:
                raise NotImplementedError("bya_b200: the single-speaker + mute-audio path (audio_embeds.ndim == 4) needs "
                                          "tests/input/ae_mute.pt, which the reference does not ship (audio_model.py:203)")
            ctx = pk.audio_context(P, ws, a.reshape(B * C, *a.shape[2:]), frames)
            out["audio_ctx"] = ctx.view(B, C, *ctx.shape[1:])
            out["aud_k"], out["aud_vt"] = pk.audio_kv(P, ws, ctx, B, C)
        return out

    # ------------------------------------------------------------------------------------------------ sequence parallel
    def enable_sequence_parallel(self, group):
        """Ulysses-style sharding of the [text; video] token axis over `group`: every rank keeps
        N/P rows and all weights; around the joint self-attention the fused QKV GEMM writes the all-to-all send buffer
        directly ([dest][rows][q|k|v heads of dest]) and the out-projection reads the returned buffer as a K-blocked A
        operand, so the two exchanges per layer need no pack / unpack kernels."""
        import torch.distributed as dist

        self.sp_group = group
        self.sp_rank = dist.get_rank(group)
        self.sp_size = P = dist.get_world_size(group)
        if self.heads % P:
            raise RuntimeError(f"bya_b200: {self.heads} heads are not divisible by the sequence-parallel size {P}")
        from .sp import qkv_rows_by_destination

        for L_ in self.layers:
            L_["w_qkv_sp"] = qkv_rows_by_destination(L_["w_qkv"], self.D, P)
            L_["b_qkv_sp"] = qkv_rows_by_destination(L_["b_qkv"], self.D, P)
        # exchange mechanism: "peer" = NVLink peer memory (push epilogues + pull kernel + device barrier, peer.py),
        # "nccl" = all_to_all_single (the baseline); "auto" tries peer memory and says so loudly if it cannot be set up
        import os
        import sys

        want = os.environ.get("BYA_SP_EXCHANGE", getattr(self.model, "sp_exchange", "auto"))
        self.peer = None
        if want in ("auto", "peer") and P > 1:
            try:
                from .peer import PeerGroup

                self.peer = PeerGroup(group, self.device)
            except Exception as e:   # noqa: BLE001
                if want == "peer":
                    raise
                print(f"bya_b200: NVLink peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-to-all",
                      file=sys.stderr, flush=True)
            # every rank must take the same path
            ok = torch.tensor([1 if self.peer is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.peer = None
        self.sp_exchange = "peer" if self.peer is not None else "nccl"

    def _prologue_cache_key(self, id_cond, id_vit_hidden, audio_embeds, frames, use_router):
        """Identity of the timestep-invariant inputs (the pipeline passes the same tensors on all 50 steps).  The key
        is (address, version, shape) per tensor; `_prologue_refs` keeps the keyed tensors of the cached generation
        alive, so the caching allocator cannot hand their addresses (with a fresh version counter) to the next
        generation's inputs while the cache still answers for them."""
        ts = list(id_cond) + [v for l in id_vit_hidden for v in l] + ([audio_embeds] if audio_embeds is not None else [])
        self._prologue_refs_pending = ts
        return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts) + (frames, use_router)

    def new_generation(self):
        """`denoise_step == 0` (transformer.py forward kwarg): drop every cached prologue, eager and graphed."""
        self._prologue_key = None
        for g in self._graphs.values():
            g["pro_key"] = None

    # ------------------------------------------------------------------------------------------------ CUDA-graph replay
    @torch.no_grad()
    def step_graphed(self, hidden_states, encoder_hidden_states, timestep, image_rotary_emb, id_cond, id_vit_hidden,
                     audio_embeds, af_matrix, routing_logits_forcing=None, per_frame_forcing=False, cache_prologue=True):
        """`step()` captured once per input geometry into a CUDA graph and replayed: the ~2 000 kernel launches of a
        step cost ~0.4 s of host time through ctypes, which bounds the step as soon as the kernels take less.  Inputs
        are copied into the graph's static buffers (device or pinned-host sources, asynchronously on the current
        stream); the returned tensor is the graph's static output (valid until the next call)."""
        dev = self.device
        nested = dict(hidden_states=hidden_states, encoder_hidden_states=encoder_hidden_states, timestep=timestep,
                      image_rotary_emb=None if image_rotary_emb is None else list(image_rotary_emb), id_cond=list(id_cond),
                      id_vit_hidden=[list(l) for l in id_vit_hidden], audio_embeds=audio_embeds, af_matrix=af_matrix,
                      routing_logits_forcing=routing_logits_forcing)

        def flat(x):
            if isinstance(x, (list, tuple)):
                return [t for y in x for t in flat(y)]
            return [x]

        def like(x):
            if isinstance(x, (list, tuple)):
                return [like(y) for y in x]
            return None if x is None else x.detach().to(dev, copy=True)

        sig = tuple(None if t is None else (tuple(t.shape), t.dtype) for t in flat(list(nested.values())))
        sig += (per_frame_forcing, cache_prologue)
        Fr = hidden_states.shape[1]
        use_router = routing_logits_forcing is None
        key = self._prologue_cache_key(id_cond, id_vit_hidden, audio_embeds, Fr, use_router) if cache_prologue else None
        g = self._graphs.get(sig)
        if g is None:
            static = {k: like(v) for k, v in nested.items()}
            rope = static["image_rotary_emb"]
            kw = dict(static, image_rotary_emb=None if rope is None else tuple(rope), per_frame_forcing=per_frame_forcing,
                      cache_prologue=False)
            pro_ws = _Workspace(dev) if cache_prologue else None
            if cache_prologue:   # the prologue stays outside the graph: it is recomputed only when its inputs change
                kw["_pro"] = self.prologue(static["id_cond"], static["id_vit_hidden"], static["audio_embeds"], Fr, use_router,
                                           ws=pro_ws)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture: workspaces, tensor maps, lazy module state
                self.step(**kw)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = ops.LAUNCHES
            # thread_local: the NCCL watchdog thread of a sequence-parallel run may query events while we capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self.step(**kw)
            g = dict(graph=graph, static=static, out=out, launches=ops.LAUNCHES - l0, pro=kw.get("_pro"), pro_key=key,
                     pro_refs=self._prologue_refs_pending if cache_prologue else None, pro_ws=pro_ws)
            self._graphs[sig] = g
        else:
            for dst, src in zip(flat(list(g["static"].values())), flat(list(nested.values()))):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            if cache_prologue and key != g["pro_key"]:
                # same workspace, same shapes -> the results land in the buffers the captured graph reads
                new = self.prologue(g["static"]["id_cond"], g["static"]["id_vit_hidden"], g["static"]["audio_embeds"], Fr,
                                    use_router, ws=g["pro_ws"])
                for dst, src in zip(flat(_tree_values(g["pro"])), flat(_tree_values(new))):
                    if dst is not None and dst.data_ptr() != src.data_ptr():
                        dst.copy_(src)
                g["pro_key"], g["pro_refs"] = key, self._prologue_refs_pending
        g["graph"].replay()
        ops.LAUNCHES += g["launches"]
        return g["out"]

    # ------------------------------------------------------------------------------------------------ one step
    @torch.no_grad()
    def step(self, hidden_states, encoder_hidden_states, timestep, image_rotary_emb, id_cond, id_vit_hidden,
             audio_embeds, af_matrix, routing_logits_forcing=None, per_frame_forcing=False, cache_prologue=True,
             taps: Optional[dict] = None, _pro: Optional[dict] = None):
        m, cfg, D, ws = self.model, self.model.config, self.D, self.ws
        dev, bf = self.device, torch.bfloat16
        B, Fr, Cin, Hl, Wl = hidden_states.shape
        p = cfg.patch_size
        gh, gw = Hl // p, Wl // p
        hw = gh * gw
        Nv = Fr * hw
        T = encoder_hidden_states.shape[1]
        N = T + Nv
        C = len(id_cond)
        use_router = routing_logits_forcing is None
        has_audio = audio_embeds is not None and len(self.audio) > 0
        fused_links = bool(getattr(m, "fused_router_links", True))   # router block chain on bya_gemm_ln_gemm_bf16
        # taps: a dict (filled with fp32 CPU copies) or a callable(name, device tensor) — single-GPU debugging aid
        tap = taps if callable(taps) else (
            (lambda k, v: taps.__setitem__(k, v.detach().float().cpu().clone())) if taps is not None else None)
        # ---- this rank's slice of the [text; video] token axis
        P, rank = getattr(self, "sp_size", 1), getattr(self, "sp_rank", 0)
        if P > 1:
            import torch.distributed as dist

            if N % P:
                raise RuntimeError(f"bya_b200: {N} tokens are not divisible by the sequence-parallel size {P}")
            if taps is not None and not callable(taps):
                raise RuntimeError("bya_b200: dict taps are a single-GPU debugging aid (a callable tap sees this rank's rows)")
        R = N // P                       # local rows
        n0 = rank * R                    # first local row (global index)
        Tl = min(max(T - n0, 0), R)      # local text rows
        Vl = R - Tl                      # local video rows
        v0 = max(n0 - T, 0)              # global index of the first local video token
        Dl, Hl_ = D // P, self.heads // P
        if m.is_train_face:
            m.router.set_grid(Fr, gh, gw)
            if getattr(self, "_router_grid", None) != (Fr, gh, gw):   # also catches 30x45 -> 45x30 (same row count)
                self.router = RouterPack(m.router)
                self._router_grid = (Fr, gh, gw)

        # ---- prologue (cached per generation)
        if _pro is not None:
            pro = _pro
        else:
            key = self._prologue_cache_key(id_cond, id_vit_hidden, audio_embeds, Fr, use_router) if cache_prologue else None
            if key is None or key != self._prologue_key:
                if cache_prologue and getattr(self, "_pro_ws", None) is None:
                    self._pro_ws = _Workspace(dev)
                self._prologue = self.prologue(id_cond, id_vit_hidden, audio_embeds, Fr, use_router,
                                               ws=self._pro_ws if cache_prologue else None)
                self._prologue_key = key
                self._prologue_refs = self._prologue_refs_pending if cache_prologue else None
            pro = self._prologue
        if tap:
            tap("face_tokens", pro["face_tokens"])
            if has_audio:
                tap("audio_ctx", pro["audio_ctx"])

        # ---- conditioning vectors
        ts_ = timestep.to(dev)
        if ts_.ndim == 0:
            ts_ = ts_[None].expand(B)
        ts_ = ts_.to(torch.int64).contiguous()
        tf = ws.get("t_feat", (B, D), torch.float32)
        ops.timestep_features(ts_, tf)
        t1 = ws.get("t_h", (B, cfg.time_embed_dim), torch.float32)
        ops.gemv(self.te[0], self.te[1], tf, t1, out_act=1)
        temb = ws.get("temb", (B, cfg.time_embed_dim), torch.float32)
        ops.gemv(self.te[2], self.te[3], t1, temb)
        ada = ws.get("ada", (B, self.ada_w.shape[0]), torch.float32)
        ops.gemv(self.ada_w, self.ada_b, temb, ada, in_act=1)
        if tap:
            tap("temb", temb)

        if image_rotary_emb is None:   # non-RoPE configuration: the QKV epilogue rotates by the identity
            idt = getattr(self, "_rope_identity", None)
            if idt is None or idt[0].shape[0] != Nv:
                idt = self._rope_identity = (torch.ones(Nv, 64, device=dev), torch.zeros(Nv, 64, device=dev))
            image_rotary_emb = idt
        cos, sin = (t.to(dev, torch.float32).contiguous() for t in image_rotary_emb)
        # every (cos, sin) pair once per row for the QKV epilogue (diffusers' tables repeat each value twice; a table that
        # does not is detected on the device and the epilogue then reads the full tables)
        rope_packed = None
        if cos.shape[1] == 64 and tuple(cos.shape) == tuple(sin.shape) and getattr(m, "packed_rope", True):
            rope_packed = ops.rope_pack(cos, sin, ws.get("rope_cs", tuple(cos.shape), torch.float32),
                                        ws.get("rope_mismatch", (1,), torch.int32))
        # joint positional table of the sincos / learned (CogVideoX-5B-I2V) configurations (transformer.py:370-392)
        pos_tab = m.patch_embed.table_for(Fr, Hl, Wl)
        if pos_tab is not None:
            if pos_tab.shape[1] != N:
                raise RuntimeError(f"bya_b200: positional table has {pos_tab.shape[1]} rows, the input {N} tokens "
                                   "(text length must equal max_text_seq_length)")
            pk = (pos_tab.data_ptr(), pos_tab._version, tuple(pos_tab.shape))
            if getattr(self, "_pos_key", None) != pk:
                self._pos_bf, self._pos_key = _bf(pos_tab[0].to(dev)), pk
            pos_tab = self._pos_bf
        forced = None
        if not use_router:
            forced = routing_logits_forcing.to(dev, torch.float32).reshape(Nv, C).contiguous()
            if not per_frame_forcing:
                forced = ops.routing_frame_or(forced, torch.empty_like(forced), Fr)

        lat = hidden_states.to(dev, bf).contiguous()
        txt = encoder_hidden_states.to(dev, bf).contiguous()
        af = af_matrix.to(dev, torch.float32).contiguous() if af_matrix is not None else None
        out = torch.empty(B, Fr, self.out_cols // (p * p), Hl, Wl, device=dev, dtype=bf)

        x_all = ws.get("x", (B, R, D))
        xn = ws.get("xn", (R, D))
        att = ws.get("att", (R, D))
        ffh = ws.get("ffh", (R, 4 * D))
        patches = ws.get("patches", (Nv, self.patch_k))
        pg0 = getattr(self, "peer", None) if P > 1 else None
        # the routing result: with the peer exchange every rank's router pushes its positions into everyone's copy
        routing = pg0.get("routing", (Nv, C), torch.float32)[0] if pg0 is not None else ws.get("routing", (Nv, C), torch.float32)
        aw = ws.get("aud_w", (max(Vl, 1), C), torch.float32)
        awsum = ws.get("aud_wsum", (max(Vl, 1),), torch.float32)
        pg = getattr(self, "peer", None) if P > 1 else None
        if P == 1:
            qkv = ws.get("qkv", (N, 3 * D))
        elif pg is None:
            qkv_send = ws.get("qkv_send", (P, R, 3 * Dl))   # [dest][local row][q|k|v heads of dest]
            qkv = ws.get("qkv", (N, 3 * Dl))                # after the exchange: every row, this rank's heads
            o_send = ws.get("o_send", (N, Dl))              # attention output of this rank's heads, every row
            o_recv = ws.get("o_recv", (P, R, Dl))           # after the exchange: local rows, [src heads]
        else:
            # PUSH exchange: the QKV epilogue stores destination d's [q|k|v] block into rank d's `qkv` rows
            # [rank*R, (rank+1)*R); the attention epilogue stores row n into its owner's o_recv[rank] — no send buffers
            qkv, qkv_peers, _ = pg.get("qkv", (N, 3 * Dl))
            o_recv, o_peers, _ = pg.get("o_recv", (P, R, Dl))
            qkv_dst = [t[rank * R:(rank + 1) * R] for t in qkv_peers]
            o_dst = [t[rank] for t in o_peers]

        for b in range(B):
            x = x_all[b]
            xv = x[Tl:]
            mod = ada[b]
            # ---- embed (transformer.py:690-695)
            if Tl:
                ops.gemm(txt[b][n0:n0 + Tl], self.text_w, x[:Tl], bias=self.text_b)
            ops.patchify(lat[b], patches)
            if Vl and pos_tab is None:
                ops.gemm(patches[v0:v0 + Vl], self.patch_w, xv, bias=self.patch_b)
            elif Vl:   # text rows of the table are zero; the video rows ride in as the epilogue's residual operand
                ops.gemm(patches[v0:v0 + Vl], self.patch_w, xv, bias=self.patch_b, mode=ops.EPI_RESIDUAL,
                         resid=pos_tab[T + v0: T + v0 + Vl])
            if tap and b == 0:
                tap("embed_video", xv)
            routing.zero_()
            rt = routing if use_router else forced  # forced masks replace the router output (transformer.py:813-819)
            ca = 0
            for i, L in enumerate(self.layers):
                o = i * 12 * D
                sh, sc, g, esh, esc, eg = (mod[o + k * D: o + (k + 1) * D] for k in range(6))
                sh2, sc2, g2, esh2, esc2, eg2 = (mod[o + (6 + k) * D: o + (7 + k) * D] for k in range(6))
                # ---- DiT block (transformer.py:223-262)
                ops.layernorm_modulate(x, xn, eps=L["ln1"][2], gamma=L["ln1"][0], beta=L["ln1"][1], mod_a=(esc, esh),
                                       mod_b=(sc, sh), split_row=Tl)
                sb = L["score_bound"]
                qpm = self.q_premul if sb is not None else 0.0
                if P == 1:
                    ops.gemm(xn, L["w_qkv"], qkv, bias=L["b_qkv"], mode=ops.EPI_QKV, split_row=Tl, ln_eps=L["qk_eps"],
                             rope=(cos, sin), nq=L["nq"], nk=L["nk"], q_premul=qpm, rope_packed=rope_packed)
                    ops.attention_d64(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, 1, N, self.heads, tag="self_attention",
                                      score_bound_log2=sb)
                    ops.gemm(att, L["w_o"], x, bias=L["b_o"], mode=ops.EPI_RESIDUAL, resid=x, gate_a=eg, gate_b=g, split_row=Tl)
                elif pg is None:
                    ops.gemm(xn, L["w_qkv_sp"], qkv_send[0], bias=L["b_qkv_sp"], mode=ops.EPI_QKV, split_row=Tl,
                             ln_eps=L["qk_eps"], rope=(cos, sin), rope_row0=v0, nq=L["nq"], nk=L["nk"], qkv_block=3 * Dl,
                             col_block=3 * Dl, col_block_stride=R * 3 * Dl, q_premul=qpm, rope_packed=rope_packed, tag="qkv_gemm")
                    dist.all_to_all_single(qkv.view(P, R, 3 * Dl), qkv_send, group=self.sp_group)
                    ops.attention_d64(qkv[:, :Dl], qkv[:, Dl:2 * Dl], qkv[:, 2 * Dl:], o_send, 1, N, Hl_, tag="self_attention",
                                      score_bound_log2=sb)
                    dist.all_to_all_single(o_recv, o_send.view(P, R, Dl), group=self.sp_group)
                    ops.gemm(o_recv[0], L["w_o"], x, bias=L["b_o"], mode=ops.EPI_RESIDUAL, resid=x, gate_a=eg, gate_b=g,
                             split_row=Tl, a_kblock=Dl, a_kblock_stride=R * Dl, tag="out_gemm")
                else:
                    ops.gemm(xn, L["w_qkv_sp"], qkv_dst[0], bias=L["b_qkv_sp"], mode=ops.EPI_QKV, split_row=Tl,
                             ln_eps=L["qk_eps"], rope=(cos, sin), rope_row0=v0, nq=L["nq"], nk=L["nk"], qkv_block=3 * Dl,
                             col_block=3 * Dl, q_premul=qpm, peer_out=qkv_dst, rope_packed=rope_packed, tag="qkv_gemm")
                    pg.barrier()      # every rank's q|k|v block has landed in my `qkv`
                    ops.attention_d64_scatter(qkv[:, :Dl], qkv[:, Dl:2 * Dl], qkv[:, 2 * Dl:], o_dst, R, N, Hl_,
                                              tag="self_attention", score_bound_log2=sb)
                    pg.barrier()      # every rank's heads have landed in my o_recv
                    ops.gemm(o_recv[0], L["w_o"], x, bias=L["b_o"], mode=ops.EPI_RESIDUAL, resid=x, gate_a=eg, gate_b=g,
                             split_row=Tl, a_kblock=Dl, a_kblock_stride=R * Dl, tag="out_gemm")
                ops.layernorm_modulate(x, xn, eps=L["ln2"][2], gamma=L["ln2"][0], beta=L["ln2"][1], mod_a=(esc2, esh2),
                                       mod_b=(sc2, sh2), split_row=Tl)
                ops.gemm(xn, L["w_f1"], ffh, bias=L["b_f1"], act=ops.ACT_GELU_TANH)
                ops.gemm(ffh, L["w_f2"], x, bias=L["b_f2"], mode=ops.EPI_RESIDUAL, resid=x, gate_a=eg2, gate_b=g2,
                         split_row=Tl)
                if tap and b == 0:
                    tap(f"block{i}.video", xv)
                    tap(f"block{i}.text", x[:T])
                # ---- face cross-attention + routing (transformer.py:737-833)
                if m.is_train_face and i % m.cross_attn_interval == 0 and ca < len(self.face):
                    Fc = self.face[ca]
                    dq = Fc["w_q"].shape[0]
                    qpad = ws.get("face_q", (R, dq))         # rows [Tl:] hold the local video queries
                    qf = qpad[Tl:]
                    peer_router = pg is not None and use_router and getattr(m, "sp_shard_router", True)
                    if Vl:
                        xnv = xn[:Vl]
                        ops.layernorm_modulate(xv, xnv, eps=Fc["ln"][2], gamma=Fc["ln"][0], beta=Fc["ln"][1])
                        ops.gemm(xnv, Fc["w_q"], qf)
                    if use_router:
                        if P == 1:
                            q_all = qf
                        elif not peer_router:  # every rank needs the queries of its router positions in all frames: gather them
                            qg = ws.get("face_q_all", (N, dq))
                            dist.all_gather_into_tensor(qg, qpad, group=self.sp_group)
                            q_all = qg[T:]
                        else:
                            q_all = None         # every owner pushes the rows a rank's router needs (run_router_sp)
                        if P > 1 and getattr(m, "sp_shard_router", True):
                            rs = self._router_shard
                            if rs is None or (rs.frames, rs.hw, rs.P) != (Fr, hw, P):
                                rs = self._router_shard = RouterShard(self.router, Fr, hw, P, rank, dev)
                            run_router_sp(self.router, rs, ws, q_all, pro["kmat"][b][ca], ca, C, routing, self.sp_group,
                                          peer=pg if peer_router else None, q_local=qpad, text_len=T, rows_per_rank=R,
                                          fused=fused_links)
                        else:
                            run_router(self.router, ws, q_all, pro["kmat"][b][ca], ca, C, Fr, hw, routing, fused=fused_links)
                        if tap and b == 0:
                            tap(f"ca{ca}.router", routing)
                    if Vl:
                        fa = ws.get("face_a", (Vl, dq))
                        k = pro["face_k"][b][ca]
                        ops.xattn_kv32(qf, k, pro["face_vt"][b][ca], rt[v0:v0 + Vl], fa, k.shape[1], k.shape[3], C, 1,
                                       k.shape[3] ** -0.5)
                        ops.gemm(fa, Fc["w_o"], xv, mode=ops.EPI_RESIDUAL, resid=xv, alpha=float(m.local_face_scale))
                    if tap and b == 0:
                        tap(f"ca{ca}.video", xv)
                    ca += 1
                # ---- audio cross-attention (transformer.py:858-936)
                if has_audio and i % m.audio_attn_interval == 0 and Vl:
                    la = i // m.audio_attn_interval
                    A = self.audio[la]
                    ops.audio_weights(af[b], rt[v0:v0 + Vl], aw, awsum)
                    xnv = xn[:Vl]
                    ops.layernorm_modulate(xv, xnv, eps=A["ln"][2], gamma=A["ln"][0], beta=A["ln"][1])
                    qa = att[:Vl]
                    ops.gemm(xnv, A["w_q"], qa, bias=A["b_q"])
                    aa = ffh[:Vl, :D]
                    k = pro["aud_k"][b][la]
                    ops.xattn_kv32(qa, k, pro["aud_vt"][b][la], aw, aa, k.shape[1], k.shape[3], C, Fr, k.shape[3] ** -0.5,
                                   tok_begin=v0, total_tokens=Nv)
                    ops.gemm(aa, A["w_o"], xv, bias=A["b_o"], mode=ops.EPI_RESIDUAL, resid=xv, row_bias_scale=awsum)
                    if tap and b == 0:
                        tap(f"audio{i}.weights", aw)
                        tap(f"audio{i}.video", xv)
            # ---- head (transformer.py:938-957)
            o = self.L * 12 * D
            shift, scale = ada[b][o: o + D], ada[b][o + D: o + 2 * D]
            ypad = ws.get("proj", (R, self.proj_w.shape[0]))
            if Vl:
                xnv = xn[:Vl]
                ops.layernorm_modulate(xv, xnv, eps=self.nf[2], gamma=self.nf[0], beta=self.nf[1])
                ops.layernorm_modulate(xnv, xnv, eps=self.no[2], gamma=self.no[0], beta=self.no[1], mod_b=(scale, shift))
                ops.gemm(xnv, self.proj_w, ypad[Tl:], bias=self.proj_b)
            if P == 1:
                y = ypad[Tl:]
            else:
                yg = ws.get("proj_all", (N, self.proj_w.shape[0]))
                dist.all_gather_into_tensor(yg, ypad, group=self.sp_group)
                y = yg[T:]
            ops.unpatchify(y, out[b])
        return out
