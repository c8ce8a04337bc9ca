"""Python-friendly signatures over the custom-op layer: tensor / shape checks here, then `torch.ops.bya.<name>`
(`custom_ops.py`: one torch.library op per C-ABI entry point of `include/bya.h`, CUDA implementation only, raw pointers
+ the current stream into libbya.so).  torch is used only for device memory, streams and op dispatch."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import custom_ops  # noqa: F401  (defines torch.ops.bya.*)
from .lib import check, lib

_bya = torch.ops.bya

EPI_STORE, EPI_RESIDUAL, EPI_QKV, EPI_SPLITK_F32 = 0, 1, 2, 3
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_RELU = 0, 1, 2, 3

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)

# Optional live kernel timing (bench.py roofline): PROFILE = {"tag": [(start_event, end_event), ...]}.  Events are
# recorded on the launching (current) stream, immediately around the launch.
PROFILE = None


def _prof(tag):
    if PROFILE is None or tag is None or tag not in PROFILE:
        return None
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    PROFILE[tag].append((s, e))
    return e


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _bf16_2d(t: torch.Tensor, name: str):
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
        raise RuntimeError(f"bya_b200: {name} must be a CUDA bf16 matrix with unit inner stride")
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias=None, act=ACT_NONE, mode=EPI_STORE,
         resid=None, gate_a=None, gate_b=None, split_row=0, alpha=1.0, row_bias_scale=None,
         qkv_block=0, ln_eps=1e-6, rope=None, rope_row0=0, nq=None, nk=None, group_m=0, col_block=0,
         col_block_stride=0, a_kblock=0, a_kblock_stride=0, q_premul=0.0, split_k=0, peer_out=None, rope_packed=None,
         tag=None) -> torch.Tensor:
    """out = epilogue(a @ w.T); a [M,K], w [N,K], out [M,N] (row strides may exceed the width).
    With peer_out (a list of N/col_block [M, col_block] tensors, possibly views of OTHER ranks' memory): column block d is
    stored to peer_out[d]; pass out=peer_out[0].
    With a_kblock: `a` is the first [M, a_kblock] block of K/a_kblock blocks a_kblock_stride elements apart.
    With col_block: `out` is the first [M, col_block] block of N/col_block blocks col_block_stride elements apart."""
    global LAUNCHES
    _bf16_2d(a, "a"), _bf16_2d(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if mode == EPI_SPLITK_F32:
        split_k = max(1, min(int(split_k), K // 64))
        if out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or tuple(out.shape) != (split_k, M, N):
            raise RuntimeError(f"bya_b200.gemm: the split-K workspace must be a contiguous CUDA fp32 [{split_k}, {M}, {N}] tensor")
    else:
        _bf16_2d(out, "out")
    if a_kblock:
        K = w.shape[1]
    if w.shape[1] != K or out.shape[-2] != M or (out.shape[-1] != (col_block if col_block else N)):
        raise RuntimeError(f"bya_b200.gemm: shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    rope_cos = rope_sin = None
    nq_w = nq_b = nk_w = nk_b = None
    if mode == EPI_RESIDUAL:
        _bf16_2d(resid, "resid")
    else:
        resid = gate_a = gate_b = row_bias_scale = None
    if mode == EPI_QKV:
        rope_cos, rope_sin = rope
        (nq_w, nq_b), (nk_w, nk_b) = nq, nk
    rope_cs, rope_mis = rope_packed if (rope_packed is not None and mode == EPI_QKV) else (None, None)
    ev = _prof(tag)
    _bya.gemm_bf16(a, w, out, bias, act, mode, resid, gate_a, gate_b, split_row, float(alpha), row_bias_scale, qkv_block,
                   float(ln_eps), rope_cos, rope_sin, rope_row0, nq_w, nq_b, nk_w, nk_b, group_m, col_block, col_block_stride,
                   a_kblock, a_kblock_stride, float(q_premul), split_k, peer_out, rope_cs, rope_mis)
    if ev is not None:
        ev.record()
    LAUNCHES += 1
    return out


def fold_layernorm(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm affine folded into the following linear: LN(x)·W^T + b = rstd * (x·W'^T - mean * csum) + b' with
    W' = bf16(W diag(gamma)), csum = rowsum(W'), b' = b + W·beta (`bya_gemm_ln_gemm_bf16`, include/bya.h)."""
    wf = (w.float() * gamma.float()[None, :]).to(torch.bfloat16).contiguous()
    csum = wf.float().sum(1).contiguous()
    b2 = w.float() @ beta.float()
    if bias is not None:
        b2 = b2 + bias.float()
    return wf, csum, b2.contiguous()


def chain_n_split(rows: int, n2: int) -> int:
    """Column slices for `gemm_ln_gemm`: with few rows one CTA pair per 256-row tile leaves most SMs idle, so the N2
    columns are cut into slices that run on different pairs (the small first GEMM is recomputed per slice)."""
    pairs = 74
    tiles = (rows + 255) // 256
    best = 1
    for s in (2, 3, 4, 6, 12):
        if (n2 // 128) % s == 0 and tiles * s <= pairs:
            best = s
    return best


def gemm_ln_gemm(a1: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], resid: torch.Tensor, x_out: torch.Tensor,
                 w2f: torch.Tensor, csum: torch.Tensor, b2: torch.Tensor, out2: torch.Tensor, *, ln_eps: float,
                 act=ACT_NONE, n_split: int = 1, store_x: bool = True, col_block=0, col_block_stride=0, a_kblock=0,
                 a_kblock_stride=0, tag=None) -> torch.Tensor:
    """x_out = resid + a1 @ w1.T + b1 ; out2 = act(LayerNorm(x_out) @ W2.T + b2) with the LayerNorm folded into
    (w2f, csum, b2) by `fold_layernorm` — one link of the router's block chain as one kernel (include/bya.h).
    a1 [M,512] (or K-blocked, see `gemm`), w1 [512,512], w2f [N2,512], out2 [M,N2] (or column-blocked)."""
    global LAUNCHES
    for t, n in ((a1, "a1"), (w1, "w1"), (w2f, "w2f"), (resid, "resid"), (x_out, "x_out"), (out2, "out2")):
        _bf16_2d(t, n)
    M = a1.shape[0]
    N2 = w2f.shape[0]
    if tuple(w1.shape) != (512, 512) or w2f.shape[1] != 512 or (a1.shape[1] != (a_kblock if a_kblock else 512)) or \
            tuple(resid.shape) != (M, 512) or tuple(x_out.shape) != (M, 512) or out2.shape[0] != M or \
            out2.shape[1] != (col_block if col_block else N2):
        raise RuntimeError(f"bya_b200.gemm_ln_gemm: shape mismatch a1{tuple(a1.shape)} w2f{tuple(w2f.shape)} out2{tuple(out2.shape)}")
    _f32(csum, "csum"), _f32(b2, "b2")
    if csum.numel() != N2 or b2.numel() != N2:
        raise RuntimeError("bya_b200.gemm_ln_gemm: csum / b2 must have N2 elements")
    ev = _prof(tag)
    _bya.gemm_ln_gemm_bf16(a1, w1, b1, resid, x_out, int(store_x), w2f, csum, b2, float(ln_eps), act, out2, n_split, col_block,
                           col_block_stride, a_kblock, a_kblock_stride)
    if ev is not None:
        ev.record()
    LAUNCHES += 1
    return out2


def attention_d64(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, batch: int, seq: int,
                  heads: int, scale: float = 0.125, tag=None, score_bound_log2: Optional[float] = None,
                  seq_stride: Optional[int] = None) -> torch.Tensor:
    """q/k/v/out: [batch*seq, heads*64] column-slice views (shared row stride for q,k,v) of bf16 matrices.
    With `score_bound_log2` (<= 64): q is pre-scaled so that q.k is in log2 units and |q.k| <= the bound
    (`bya_attention_d64_bounded`); `scale` is then ignored."""
    global LAUNCHES
    stride = seq if seq_stride is None else seq_stride
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _bf16_2d(t, n)
        if t.shape[0] < (batch - 1) * stride + seq or t.shape[1] != heads * 64 or (seq_stride is None and t.shape[0] != batch * seq):
            raise RuntimeError(f"bya_b200.attention_d64: {n} has shape {tuple(t.shape)}")
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise RuntimeError("bya_b200.attention_d64: q, k, v must share a row stride")
    ev = _prof(tag)
    variant = 2 if score_bound_log2 is not None else (1 if seq_stride is not None else 0)
    _bya.attention_d64(q, k, v, out, batch, seq, stride, heads, float(scale),
                       float(score_bound_log2) if score_bound_log2 is not None else 0.0, variant)
    if ev is not None:
        ev.record()
    LAUNCHES += 1
    return out


def _f32(t: Optional[torch.Tensor], name: str):
    if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
        raise RuntimeError(f"bya_b200: {name} must be a contiguous CUDA fp32 tensor")
    return t


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def layernorm_modulate(x, out, *, eps=1e-5, gamma=None, beta=None, mod_a=None, mod_b=None, split_row=0, add=None):
    """out = (LN(x)*gamma+beta)*(1+scale)+shift (+add[row % add_rows]); mod_* = (scale, shift) fp32 [dim] for rows
    < split_row (a) and >= split_row (b)."""
    _bf16_2d(x, "x"), _bf16_2d(out, "out")
    rows, dim = x.shape
    sa, ha = (mod_a if mod_a is not None else (None, None))
    sb, hb = (mod_b if mod_b is not None else (None, None))
    for t in (sa, ha, sb, hb):
        _f32(t, "modulation")
    if add is not None:
        _bf16_2d(add, "add")
        if not add.is_contiguous() or add.shape[1] != dim:
            raise RuntimeError("bya_b200.layernorm_modulate: add must be contiguous [rows, dim]")
    _bya.layernorm_modulate(x, out, float(eps), gamma, beta, sa, ha, sb, hb, split_row, add)
    _count()
    return out


def gemv(w, bias, x, y, in_act=0, out_act=0):
    """y[b] = out_act(W @ in_act(x[b]) + bias); W bf16 [N,K], x fp32 [B,K], y fp32 [B,N]."""
    _bf16_2d(w, "w"), _f32(x, "x"), _f32(y, "y")
    if not w.is_contiguous():
        raise RuntimeError("bya_b200.gemv: W must be contiguous")
    B, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K or tuple(y.shape) != (B, N):
        raise RuntimeError("bya_b200.gemv: shape mismatch")
    _bya.gemv(w, bias, x, y, in_act, out_act)
    _count()
    return y


def rope_pack(cos: torch.Tensor, sin: torch.Tensor, packed: torch.Tensor, mismatch: torch.Tensor):
    """cos / sin fp32 [rows, 64] -> packed fp32 [rows, 64] = [cos of the 32 pairs | sin of the 32 pairs]; `mismatch` (int32
    [1], zeroed here) becomes 1 if the two values of some pair differ, in which case `gemm(..., rope_packed=...)` falls back
    to the full tables on its own (device-side flag: no host synchronisation, safe inside a CUDA graph)."""
    _f32(cos, "cos"), _f32(sin, "sin"), _f32(packed, "packed")
    if tuple(cos.shape) != tuple(sin.shape) or cos.shape[1] != 64 or tuple(packed.shape) != tuple(cos.shape) or \
            mismatch.dtype != torch.int32 or not mismatch.is_cuda:
        raise RuntimeError("bya_b200.rope_pack: cos / sin / packed must be fp32 [rows, 64], mismatch CUDA int32")
    memset_zero(mismatch)
    _bya.rope_pack(cos, sin, packed, mismatch)
    _count()
    return packed, mismatch


def timestep_features(t, out):
    if t.dtype != torch.int64 or not t.is_cuda:
        raise RuntimeError("bya_b200.timestep_features: timestep must be CUDA int64")
    _f32(out, "out")
    _bya.timestep_features(t, out)
    _count()
    return out


def patchify(latents, out):
    """latents bf16 [F,C,H,W] contiguous -> out bf16 [F*H/2*W/2, ld] (zero padded columns)."""
    F, C, H, W = latents.shape
    if latents.dtype != torch.bfloat16 or not latents.is_contiguous():
        raise RuntimeError("bya_b200.patchify: latents must be contiguous bf16")
    _bf16_2d(out, "out")
    _bya.patchify(latents, out)
    _count()
    return out


def unpatchify(y, out):
    """y bf16 [F*gh*gw, >=C*4] -> out bf16 [F,C,2gh,2gw]."""
    F, C, H, W = out.shape
    _bf16_2d(y, "y")
    if out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise RuntimeError("bya_b200.unpatchify: out must be contiguous bf16")
    _bya.unpatchify(y, out)
    _count()
    return out


def router_head(x, w, b, r, rows, chars):
    _bf16_2d(x, "x"), _f32(r, "r")
    _bya.router_head(x, w, b, r, rows, chars)
    _count()
    return r


def xattn_kv32(q, K, Vt, w, out, heads, head_dim, chars, kv_frames, scale, tok_begin=0, total_tokens=0):
    """Routed 32-key cross-attention; q/out bf16 [tokens, heads*head_dim] views, K [G,H,32,d], Vt [G,H,d,32]."""
    _bf16_2d(q, "q"), _bf16_2d(out, "out"), _f32(w, "w")
    tokens = q.shape[0]
    G = chars * kv_frames
    if tuple(K.shape) != (G, heads, 32, head_dim) or tuple(Vt.shape) != (G, heads, head_dim, 32):
        raise RuntimeError(f"bya_b200.xattn_kv32: K{tuple(K.shape)} / Vt{tuple(Vt.shape)} do not match")
    if not (K.is_contiguous() and Vt.is_contiguous() and K.dtype == torch.bfloat16 and Vt.dtype == torch.bfloat16):
        raise RuntimeError("bya_b200.xattn_kv32: K / Vt must be contiguous bf16")
    if w is not None and tuple(w.shape) != (tokens, chars):
        raise RuntimeError("bya_b200.xattn_kv32: w must be [tokens, chars]")
    _bya.xattn_kv32(q, K, Vt, w, out, heads, head_dim, chars, kv_frames, float(scale), int(tok_begin), int(total_tokens))
    _count()
    return out


def small_attention(qkv, out, n_seq, seq_len, heads, inner, outer_stride, tok_stride, scale=0.125):
    _bf16_2d(qkv, "qkv"), _bf16_2d(out, "out")
    _bya.small_attention(qkv, out, n_seq, seq_len, heads, inner, int(outer_stride), int(tok_stride), float(scale))
    _count()
    return out


def masks_to_routing(masks, frames, grid_h, grid_w, index_mask=None, logits=None):
    """masks uint8 [C,T,H,W] (CUDA) -> (index_mask int64 [Nv], logits fp32 [Nv,C]); bit-exact vs the reference."""
    if masks.dtype != torch.uint8 or not masks.is_cuda or not masks.is_contiguous() or masks.dim() != 4:
        raise RuntimeError("bya_b200.masks_to_routing: masks must be contiguous CUDA uint8 [C,T,H,W]")
    C, T, H, W = masks.shape
    n = frames * grid_h * grid_w
    if index_mask is None:
        index_mask = torch.empty(n, dtype=torch.int64, device=masks.device)
    if logits is None:
        logits = torch.empty(n, C, dtype=torch.float32, device=masks.device)
    _bya.masks_to_routing(masks, frames, grid_h, grid_w, index_mask, logits)
    _count()
    return index_mask, logits


def routing_frame_or(logits, out, frames):
    _f32(logits, "logits"), _f32(out, "out")
    n, C = logits.shape
    _bya.routing_frame_or(logits, out, frames)
    _count()
    return out


def audio_weights(af, routing, w, wsum=None):
    _f32(af, "af"), _f32(routing, "routing"), _f32(w, "w"), _f32(wsum, "wsum")
    n, C = routing.shape
    _bya.audio_weights(af, routing, w, wsum)
    _count()
    return w


PRED_EPSILON, PRED_SAMPLE, PRED_V = 0, 1, 2
DPM_NCOEF = 12   # include/bya.h BYA_DPM_*: guidance, sqrt_alpha, sqrt_beta, mult0..3, mult_noise, second_order, 1/sqrt_alpha


def cfg_dpm_step(model_out, sample, prev_sample, old_pred, pred_out, noise, coef, *, prediction_type=PRED_V,
                 step_index=None, model_input=None):
    """Guidance combine + CogVideoXDPMScheduler.step + write of x_{t-1} into the next model input (one kernel).
    model_out: bf16 [B, F, C, H, W] with B in {1, 2} (B = 2: [uncond | cond]) or fp32 [1, F, C, H, W] / [F, C, H, W];
    sample / prev_sample: bf16 [.., F, C, H, W] (may alias); old_pred / pred_out: fp32 same numel (may alias);
    noise: bf16 [steps, 2, numel]; coef: fp32 [steps, DPM_NCOEF]; step_index: int32 [1] on the device or None;
    model_input: optional bf16 [Bin, F, Cin, H, W] whose channels [0, C) receive x_{t-1}."""
    F, C, H, W = sample.shape[-4:]
    n = F * C * H * W

    def chk(t, name, dtype, numel=None):
        if t.dtype != dtype or not t.is_cuda or not t.is_contiguous() or (numel is not None and t.numel() != numel):
            raise RuntimeError(f"bya_b200.cfg_dpm_step: {name} must be contiguous CUDA {dtype}"
                               + (f" with {numel} elements" if numel is not None else ""))

    chk(sample, "sample", torch.bfloat16, n), chk(prev_sample, "prev_sample", torch.bfloat16, n)
    chk(old_pred, "old_pred", torch.float32, n), chk(pred_out, "pred_out", torch.float32, n)
    chk(noise, "noise", torch.bfloat16), chk(coef, "coef", torch.float32)
    if noise.dim() != 3 or noise.shape[1] != 2 or noise.shape[2] != n or coef.dim() != 2 or coef.shape[1] != DPM_NCOEF \
            or coef.shape[0] != noise.shape[0]:
        raise RuntimeError("bya_b200.cfg_dpm_step: noise must be [steps, 2, numel] and coef [steps, 12]")
    if model_out.dtype == torch.float32:
        chk(model_out, "model_out", torch.float32, n)
    else:
        chk(model_out, "model_out", torch.bfloat16)
        if model_out.numel() not in (n, 2 * n):
            raise RuntimeError("bya_b200.cfg_dpm_step: model_out must hold one or two (CFG) predictions")
    if step_index is not None:
        if step_index.dtype != torch.int32 or not step_index.is_cuda:
            raise RuntimeError("bya_b200.cfg_dpm_step: step_index must be a CUDA int32 tensor")
    elif coef.shape[0] != 1:
        raise RuntimeError("bya_b200.cfg_dpm_step: without step_index pass exactly this step's coef row and noise pair")
    if model_input is not None:
        chk(model_input, "model_input", torch.bfloat16)
        Bi, Fi, Ci, Hi, Wi = model_input.shape
        if (Fi, Hi, Wi) != (F, H, W) or Ci < C:
            raise RuntimeError("bya_b200.cfg_dpm_step: model_input must be [B, F, >=C, H, W]")
    _bya.cfg_dpm_step(model_out, sample, prev_sample, old_pred, pred_out, noise, coef, prediction_type, step_index, model_input)
    _count()
    return prev_sample, pred_out


def denoise_select_step(timesteps, timestep_out, counter, step_index):
    """timestep_out[:] = timesteps[*counter]; *step_index = *counter; *counter += 1 (all on the device)."""
    if timesteps.dtype != torch.int64 or timestep_out.dtype != torch.int64 or counter.dtype != torch.int32 \
            or step_index.dtype != torch.int32 or not (timesteps.is_cuda and timestep_out.is_cuda and counter.is_cuda
                                                       and step_index.is_cuda):
        raise RuntimeError("bya_b200.denoise_select_step: int64 timestep tensors and int32 counters on the device")
    _bya.denoise_select_step(timesteps, timestep_out, counter, step_index)
    _count()
    return timestep_out


# ---------------------------------------------------------------- per-generation prologue helpers
def copy2d(src, out):
    """out[r, c] = bf16(src[r, c]); 2-D views with unit inner stride (src bf16 or fp32), row strides free."""
    _bf16_2d(out, "out")
    if src.dim() != 2 or src.stride(1) != 1 or not src.is_cuda or src.dtype not in (torch.bfloat16, torch.float32) \
            or tuple(src.shape) != tuple(out.shape):
        raise RuntimeError(f"bya_b200.copy2d: src {tuple(src.shape)} {src.dtype} vs out {tuple(out.shape)}")
    _bya.copy2d(src, out)
    _count()
    return out


def memset_zero(t):
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError("bya_b200.memset_zero: contiguous CUDA tensor expected")
    _bya.memset_zero(t)
    return t


def splitk_finalize(ws, bias, act, out, row0=0):
    """out = act(sum_s ws[s, row0:row0+rows] + bias) as bf16; ws fp32 [splits, M, N] filled by gemm(mode=EPI_SPLITK_F32)."""
    _bf16_2d(out, "out")
    if ws.dtype != torch.float32 or ws.dim() != 3 or not ws.is_contiguous() or ws.shape[2] != out.shape[1] \
            or row0 + out.shape[0] > ws.shape[1]:
        raise RuntimeError("bya_b200.splitk_finalize: ws must be contiguous fp32 [splits, M, N] covering the rows of out")
    _bya.splitk_finalize(ws, row0, bias, act, out)
    _count()
    return out


def layernorm_leakyrelu(x, out, gamma, beta, eps=1e-5, slope=0.01):
    _bf16_2d(x, "x"), _bf16_2d(out, "out")
    _bya.layernorm_leakyrelu(x, out, float(eps), gamma, beta, float(slope))
    _count()
    return out


def kv_pack(x, k_off, v_off, K, Vt):
    """x [G*32, ld] -> K [G,H,32,d], Vt [G,H,d,32] (contiguous bf16)."""
    _bf16_2d(x, "x")
    G, H, n, d = K.shape
    if n != 32 or tuple(Vt.shape) != (G, H, d, 32) or x.shape[0] != G * 32 or not (K.is_contiguous() and Vt.is_contiguous()) \
            or K.dtype != torch.bfloat16 or Vt.dtype != torch.bfloat16:
        raise RuntimeError("bya_b200.kv_pack: shape mismatch")
    _bya.kv_pack(x, k_off, v_off, K, Vt)
    _count()
    return K, Vt


def router_keys_scatter(k, mat, chars, heads, head_dim):
    _bf16_2d(k, "k"), _bf16_2d(mat, "mat")
    if tuple(mat.shape) != (chars * 32 * heads, heads * head_dim) or not mat.is_contiguous() or k.shape[0] != chars * 32:
        raise RuntimeError("bya_b200.router_keys_scatter: shape mismatch")
    _bya.router_keys_scatter(k, mat, chars, heads, head_dim)
    _count()
    return mat


# ---------------------------------------------------------------- exchanges over NVLink peer memory
def attention_d64_scatter(q, k, v, out_peers, rows_per_peer, seq, heads, scale=0.125, tag=None, score_bound_log2=None):
    """Attention of this rank's heads over all `seq` rows; row n is stored to out_peers[n // rows_per_peer][n % rows_per_peer]
    (tensors [rows_per_peer, heads*64], possibly views of other ranks' memory)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _bf16_2d(t, n)
        if t.shape[0] != seq or t.shape[1] != heads * 64:
            raise RuntimeError(f"bya_b200.attention_d64_scatter: {n} has shape {tuple(t.shape)}")
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise RuntimeError("bya_b200.attention_d64_scatter: q, k, v must share a row stride")
    ld = out_peers[0].stride(0)
    for t in out_peers:
        _bf16_2d(t, "out_peers")
        if tuple(t.shape) != (rows_per_peer, heads * 64) or t.stride(0) != ld:
            raise RuntimeError("bya_b200.attention_d64_scatter: out_peers must be [rows_per_peer, heads*64] with one row stride")
    if len(out_peers) * rows_per_peer < seq:
        raise RuntimeError("bya_b200.attention_d64_scatter: the peers do not cover the sequence")
    ev = _prof(tag)
    _bya.attention_d64_scatter(q, k, v, list(out_peers), rows_per_peer, seq, heads, float(scale),
                               float(score_bound_log2) if score_bound_log2 is not None else 0.0)
    if ev is not None:
        ev.record()
    _count()


def peer_barrier(counter, flag_ptrs, my_rank, n_ranks):
    if counter.dtype != torch.int32 or flag_ptrs.dtype != torch.int64 or not counter.is_cuda or not flag_ptrs.is_cuda \
            or flag_ptrs.numel() < n_ranks:
        raise RuntimeError("bya_b200.peer_barrier: int32 counter and an int64 pointer table on the device")
    ev = _prof("peer_barrier")
    _bya.peer_barrier(counter, flag_ptrs, my_rank, n_ranks)
    if ev is not None:
        ev.record()
    _count()


def peer_copy(segs, n_segs, peer_ptrs, local, push, vec_bytes, blocks_per_seg=8):
    """Strided segments between `local` and the peers' buffers (`peer_ptrs`: device table of base pointers); push: local ->
    peers (posted writes), else peers -> local."""
    if segs.dtype != torch.uint8 or segs.numel() != n_segs * 64 or not segs.is_cuda or peer_ptrs.dtype != torch.int64 \
            or not local.is_cuda or not local.is_contiguous():
        raise RuntimeError("bya_b200.peer_copy: bad segment table / pointer table / local buffer")
    ev = _prof("peer_copy")
    _bya.peer_copy(segs, n_segs, peer_ptrs, local, int(bool(push)), vec_bytes, blocks_per_seg)
    if ev is not None:
        ev.record()
    _count()
