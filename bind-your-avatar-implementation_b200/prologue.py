"""The per-generation prologue on the package's own kernels.

Everything the reference recomputes inside every `forward` although it does not depend on the timestep or the latents
(SURVEY.md §0.10, Appendix A.4):

  face tokens     LocalFacialExtractor            models/router.py:157-193 (+ PerceiverAttention :46-75)
  face K / V^T    to_kv(norm1(face)) per layer    models/router.py:247-254
  router keys     to_k[l](norm_k(k)) per layer    models/router.py:377-383  (block-structured for the dense score GEMM)
  audio context   sliding windows + AudioProjModel models/audio_model.py:188-193, :78-114
  audio K / V^T   to_k / to_v per layer           models/audio_model.py:241-256 (diffusers Attention)

Round 1 ran this through torch (cuBLAS / ATen: 914 library launches, 6.8 ms).  Here it is libbya.so only: tcgen05 GEMMs
(the two skinny weight-streaming ones — AudioProjModel.proj1 and its 1.2 B-parameter Conv1d(k=2, s=2) taken as a
[rows, 2*C_in] x [C_out, 2*C_in] product — as split-K GEMMs that put every SM on the 2.4 GB weight stream), the LayerNorm
kernels, the tcgen05 attention for LocalFacialExtractor's 37-query x 619-key perceiver attention, and a handful of
data-movement kernels (`csrc/prologue.cu`).  All characters and CFG batch elements go through one batch.
"""
from __future__ import annotations

import torch

from . import ops


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _split_k(n_tiles: int, num_kb: int, sms: int = 148) -> int:
    """k-splits so that tiles x splits is about two waves of CTAs, each split keeping >= 4 k-blocks."""
    return max(1, min(num_kb // 4, round(2 * sms / max(n_tiles, 1))))


class ProloguePack:
    """Prologue weights in kernel layout (built once per weight version by `StepEngine._pack`)."""

    def __init__(self, model, router_perm):
        m = model
        lfe = m.local_facial_extractor
        self.lfe = lfe

        def mlp(seq):
            return dict(w0=_bf(seq[0].weight), b0=_bf(seq[0].bias), g1=_bf(seq[1].weight), h1=_bf(seq[1].bias), e1=seq[1].eps,
                        w3=_bf(seq[3].weight), b3=_bf(seq[3].bias), g4=_bf(seq[4].weight), h4=_bf(seq[4].bias), e4=seq[4].eps,
                        w6=_bf(seq[6].weight), b6=_bf(seq[6].bias), slope=seq[2].negative_slope)

        self.id_map = mlp(lfe.id_embedding_mapping)
        self.vit_maps = [mlp(getattr(lfe, f"mapping_{i}")) for i in range(5)]
        self.latents = _bf(lfe.latents[0])                       # [32, 1024]
        self.proj_out_t = _bf(lfe.proj_out.t())                  # [2048, 1024]: W of the final projection GEMM
        self.layers = []
        for attn, ff in lfe.layers:
            self.layers.append(dict(
                n1=(_bf(attn.norm1.weight), _bf(attn.norm1.bias), attn.norm1.eps),
                n2=(_bf(attn.norm2.weight), _bf(attn.norm2.bias), attn.norm2.eps),
                w_q=_bf(attn.to_q.weight), w_kv=_bf(attn.to_kv.weight), w_o=_bf(attn.to_out.weight),
                heads=attn.heads, dh=attn.dim_head,
                nf=(_bf(ff[0].weight), _bf(ff[0].bias), ff[0].eps), w1=_bf(ff[1].weight), w3=_bf(ff[3].weight)))
        # face K/V and router keys per cross-attention layer
        self.face = []
        for j, ca in enumerate(m.perceiver_cross_attention):
            self.face.append(dict(n1=(_bf(ca.norm1.weight), _bf(ca.norm1.bias), ca.norm1.eps), w_kv=_bf(ca.to_kv.weight),
                                  heads=ca.heads, dh=ca.dim_head,
                                  w_rk=_bf(m.router.to_k[j].weight[:, router_perm])))
        r = m.router
        self.rk_norm = (_bf(r.norm_k.weight[router_perm]), _bf(r.norm_k.bias[router_perm]), r.norm_k.eps)
        # audio
        self.audio = None
        if getattr(m, "is_train_audio", False):
            am, ap = m.audio_model, m.audio_model.audio_proj_model
            self.audio = dict(
                w1=_bf(ap.proj1.weight), b1=_bf(ap.proj1.bias), w2=_bf(ap.proj2.weight), b2=_bf(ap.proj2.bias),
                w3=_bf(ap.proj3.weight), b3=_bf(ap.proj3.bias),
                # Conv1d(k=2, s=2) over time-major pairs: Wc[o, k*C_in + c] = W[o, c, k]
                wc=_bf(ap.conv1.weight.permute(0, 2, 1).reshape(ap.conv1.weight.shape[0], -1)), bc=_bf(ap.conv1.bias),
                norm=(_bf(ap.norm.weight), _bf(ap.norm.bias), ap.norm.eps), tokens=ap.context_tokens, dim=ap.output_dim,
                window=am.window_size, stride=am.window_stride, heads=am.heads, dh=am.head_dim,
                kv=[(_bf(torch.cat([l["attn"].to_k.weight, l["attn"].to_v.weight], 0)),
                     _bf(torch.cat([l["attn"].to_k.bias, l["attn"].to_v.bias], 0))) for l in am.layers])


def _mlp(P, ws, tag, x, out_views):
    """Linear -> LayerNorm -> LeakyReLU -> Linear -> LayerNorm -> LeakyReLU -> Linear (router.py:118-154 mapping MLPs).
    x [M, K]; the last Linear is evaluated once per entry of `out_views` = [(row slice of x's rows, out view)]."""
    M = x.shape[0]
    t0 = ws.get(f"{tag}_t0", (M, 1024))
    t1 = ws.get(f"{tag}_t1", (M, 1024))
    ops.gemm(x, P["w0"], t0, bias=P["b0"])
    ops.layernorm_leakyrelu(t0, t1, P["g1"], P["h1"], eps=P["e1"], slope=P["slope"])
    ops.gemm(t1, P["w3"], t0, bias=P["b3"])
    ops.layernorm_leakyrelu(t0, t1, P["g4"], P["h4"], eps=P["e4"], slope=P["slope"])
    for rows, out in out_views:
        ops.gemm(t1[rows], P["w6"], out, bias=P["b6"])


def face_tokens(P: ProloguePack, ws, id_cond, id_vit_hidden, dev):
    """LocalFacialExtractor.forward for every (batch element, character) at once -> [B, C, 32, 2048]."""
    C, B = len(id_cond), id_cond[0].shape[0]
    S = B * C                      # sample s = b*C + c
    NQ, NI, NV, D = 32, 5, id_vit_hidden[0][0].shape[1], 1024
    NL, NC = NQ + NI, NI + NV      # latents rows (37), context rows (582)
    NK = NC + NL                   # keys of the perceiver attention (619)
    xid = ws.get("lfe_xid_in", (S, id_cond[0].shape[1]))
    for c in range(C):
        ops.copy2d(id_cond[c].to(dev), xid[c::C])
    xe = ws.get("lfe_xid", (S, NI * D))
    _mlp(P.id_map, ws, "lfe_id", xid, [(slice(0, S), xe)])
    lat = ws.get("lfe_lat", (S * NL, D))
    ops.copy2d(P.latents.view(1, NQ * D).expand(S, NQ * D), lat.view(S, NL * D)[:, :NQ * D])
    ops.copy2d(xe, lat.view(S, NL * D)[:, NQ * D:])
    ctx = ws.get("lfe_ctx", (S * NC, D))
    ops.copy2d(xe, ctx.view(S, NC * D)[:, :NI * D])
    y = ws.get("lfe_y", (S * NV, D))
    kvin = ws.get("lfe_kvin", (S * NK, D))
    qkv = ws.get_zeroed("lfe_qkv", (S * NK, 3 * D))     # q of the context rows stays zero: their outputs are never read
    att = ws.get("lfe_att", (S * NK, D))
    ffn_n = ws.get("lfe_ffn_n", (S * NL, D))
    ffn_h = ws.get("lfe_ffn_h", (S * NL, 4 * D))
    depth = P.lfe.depth
    for i in range(5):
        for c in range(C):
            v = id_vit_hidden[c][i].to(dev)
            for b in range(B):
                s = b * C + c
                ops.copy2d(v[b], y[s * NV:(s + 1) * NV])
        _mlp(P.vit_maps[i], ws, "lfe_vit", y,
             [(slice(s * NV, (s + 1) * NV), ctx[s * NC + NI:(s + 1) * NC]) for s in range(S)])
        for L in P.layers[i * depth:(i + 1) * depth]:
            H = L["heads"]
            for s in range(S):   # keys / values = [norm1(ctx) ; norm2(latents)] (router.py:58-64)
                ops.layernorm_modulate(ctx[s * NC:(s + 1) * NC], kvin[s * NK:s * NK + NC], eps=L["n1"][2], gamma=L["n1"][0], beta=L["n1"][1])
                ops.layernorm_modulate(lat[s * NL:(s + 1) * NL], kvin[s * NK + NC:(s + 1) * NK], eps=L["n2"][2], gamma=L["n2"][0], beta=L["n2"][1])
            ops.gemm(kvin, L["w_kv"], qkv[:, D:])
            for s in range(S):
                ops.gemm(kvin[s * NK + NC:(s + 1) * NK], L["w_q"], qkv[s * NK + NC:(s + 1) * NK, :D])
            # (q s)(k s)^T with s = dh^-1/4  ==  q k^T / sqrt(dh); softmax in fp32 (router.py:66-71)
            ops.attention_d64(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, S, NK, H, scale=L["dh"] ** -0.5)
            for s in range(S):
                ls = lat[s * NL:(s + 1) * NL]
                ops.gemm(att[s * NK + NC:(s + 1) * NK], L["w_o"], ls, mode=ops.EPI_RESIDUAL, resid=ls)
            ops.layernorm_modulate(lat, ffn_n, eps=L["nf"][2], gamma=L["nf"][0], beta=L["nf"][1])
            ops.gemm(ffn_n, L["w1"], ffn_h, act=ops.ACT_GELU_ERF)
            ops.gemm(ffn_h, L["w3"], lat, mode=ops.EPI_RESIDUAL, resid=lat)
    face = ws.get("face_tokens", (S * NQ, P.proj_out_t.shape[0]))
    for s in range(S):
        ops.gemm(lat[s * NL:s * NL + NQ], P.proj_out_t, face[s * NQ:(s + 1) * NQ])
    return face.view(B, C, NQ, -1)


def face_kv_and_router_keys(P: ProloguePack, ws, face, use_router: bool):
    """face [B, C, 32, 2048] -> per batch element and cross-attention layer: K [C,16,32,128], V^T [C,16,128,32] and
    (when the learned router runs) the block-structured routed keys [C*512, 2048]."""
    B, C, NQ, KD = face.shape
    S = B * C
    f2 = face.view(S * NQ, KD)
    fn = ws.get("face_n", (S * NQ, KD))
    out = dict(face_k=[[] for _ in range(B)], face_vt=[[] for _ in range(B)], kmat=[[] for _ in range(B)])
    for j, L in enumerate(P.face):
        H, dh = L["heads"], L["dh"]
        kv = ws.get("face_kv", (S * NQ, 2 * H * dh))
        ops.layernorm_modulate(f2, fn, eps=L["n1"][2], gamma=L["n1"][0], beta=L["n1"][1])
        ops.gemm(fn, L["w_kv"], kv)
        K = ws.get(f"face_K{j}", (S, H, NQ, dh))
        Vt = ws.get(f"face_Vt{j}", (S, H, dh, NQ))
        ops.kv_pack(kv, 0, H * dh, K, Vt)
        if use_router:
            kn = ws.get("face_rkn", (S * NQ, H * dh))
            rk = ws.get("face_rk", (S * NQ, H * dh))
            ops.layernorm_modulate(kv[:, :H * dh], kn, eps=P.rk_norm[2], gamma=P.rk_norm[0], beta=P.rk_norm[1])
            ops.gemm(kn, L["w_rk"], rk)
        for b in range(B):
            out["face_k"][b].append(K[b * C:(b + 1) * C])
            out["face_vt"][b].append(Vt[b * C:(b + 1) * C])
            if use_router:
                mat = ws.get(f"face_kmat{j}_{b}", (C * NQ * H, H * dh))
                ops.router_keys_scatter(rk[b * C * NQ:(b + 1) * C * NQ], mat, C, H, dh)
                out["kmat"][b].append(mat)
            else:
                out["kmat"][b].append(None)
    return out


def audio_context(P: ProloguePack, ws, audio, frames: int):
    """audio [R, 4(F-1)+5, 12, 768] bf16 -> context tokens [R, F, 32, 768] (sliding_windows + AudioProjModel)."""
    A = P.audio
    R, T = audio.shape[:2]
    win, st = A["window"], A["stride"]
    want = 1 + (frames - 1) * 4 + (win - st)
    assert want == T, (f"hidden_states_num_frames: {frames}, window_size: {win}, window_stride: {st}, "
                       f"audio_embeds.shape[1]: {T}")           # audio_model.py:190
    per = audio.shape[2] * audio.shape[3]                        # 12 * 768 elements per audio frame
    Lw = (T - win) // st + 1                                     # windows (49)
    a2 = audio.contiguous().view(R, T * per)
    # window l of sample r = `win` consecutive audio frames = one contiguous run starting at frame l*st
    x1 = ws.get("aud_win", (R * Lw, win * per))
    for r in range(R):
        ops.copy2d(torch.as_strided(a2[r], (Lw, win * per), (st * per, 1)), x1[r * Lw:(r + 1) * Lw])
    inter = A["w1"].shape[0]
    # (k-splits chosen from the weight shape only: the summation order, hence the result, must not depend on how many
    # characters / CFG branches share the batch)
    sk = _split_k(inter // 256 if inter % 256 == 0 else inter // 64, x1.shape[1] // 64)
    acc1 = ws.get("aud_acc1", (sk, R * Lw, inter), torch.float32)
    ops.gemm(x1, A["w1"], acc1, mode=ops.EPI_SPLITK_F32, split_k=sk)
    h1 = ws.get("aud_h1", (R * Lw, inter))
    ops.splitk_finalize(acc1, A["b1"], ops.ACT_RELU, h1)
    h2 = ws.get("aud_h2", (R * Lw, inter))
    ops.gemm(h1, A["w2"], h2, bias=A["b2"], act=ops.ACT_RELU)
    CD = A["w3"].shape[0]                                        # 32 * 768
    x = ws.get("aud_x0", (R * Lw, CD))
    ops.gemm(h2, A["w3"], x, bias=A["b3"])
    L = Lw
    level = 0
    for _ in range(2):      # 49 -> 25 -> 13: keep frame 0 of an odd-length sequence, Conv1d(k=2, s=2) on the rest
        keep = L % 2
        pairs = (L - keep) // 2
        if pairs == 0:
            continue
        level += 1
        Ln = keep + pairs
        pin = ws.get(f"aud_pairs{level}", (R * pairs, 2 * CD))
        xv = x.view(R, L * CD)
        ops.copy2d(xv[:, keep * CD:], pin.view(R, pairs * 2 * CD))   # frames (keep + 2i, keep + 2i + 1) side by side
        sk = _split_k(CD // 256, 2 * CD // 64)
        acc = ws.get(f"aud_acc_c{level}", (sk, R * pairs, CD), torch.float32)
        ops.gemm(pin, A["wc"], acc, mode=ops.EPI_SPLITK_F32, split_k=sk)
        xn = ws.get(f"aud_x{level}", (R * Ln, CD))
        if keep:
            ops.copy2d(xv[:, :CD], xn.view(R, Ln * CD)[:, :CD])
        for r in range(R):
            ops.splitk_finalize(acc, A["bc"], ops.ACT_NONE, xn[r * Ln + keep:(r + 1) * Ln], row0=r * pairs)
        x, L = xn, Ln
    tokens, dim = A["tokens"], A["dim"]
    ctx = ws.get("audio_ctx", (R * L * tokens, dim))
    ops.layernorm_modulate(x.view(R * L * tokens, dim), ctx, eps=A["norm"][2], gamma=A["norm"][0], beta=A["norm"][1])
    return ctx.view(R, L, tokens, dim)


def audio_kv(P: ProloguePack, ws, ctx, B: int, C: int):
    """ctx [B*C, F, 32, 768] -> per batch element and audio layer: K [C*F,48,32,64], V^T [C*F,48,64,32]."""
    A = P.audio
    R, Fr, tokens, dim = ctx.shape
    H, dh = A["heads"], A["dh"]
    c2 = ctx.view(R * Fr * tokens, dim)
    ks = [[] for _ in range(B)]
    vs = [[] for _ in range(B)]
    kv = ws.get("aud_kv", (R * Fr * tokens, 2 * H * dh))
    for l, (w, b) in enumerate(A["kv"]):
        ops.gemm(c2, w, kv, bias=b)
        K = ws.get(f"aud_K{l}", (R * Fr, H, tokens, dh))
        Vt = ws.get(f"aud_Vt{l}", (R * Fr, H, dh, tokens))
        ops.kv_pack(kv, 0, H * dh, K, Vt)
        for bb in range(B):
            ks[bb].append(K[bb * C * Fr:(bb + 1) * C * Fr])
            vs[bb].append(Vt[bb * C * Fr:(bb + 1) * C * Fr])
    return ks, vs
