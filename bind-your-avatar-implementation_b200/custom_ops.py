"""The `torch.library` custom-op layer over the C ABI (`include/bya.h`): one op `torch.ops.bya.<name>` per entry point
of libbya.so, CUDA implementation only.

Each implementation turns its tensor arguments into raw device pointers, takes torch's current stream and calls the
`extern "C"` function through ctypes; outputs are caller-allocated and declared as mutated arguments in the schema, so
the ops never allocate.  There is deliberately no CPU / Meta / CompositeImplicit registration: dispatching one of these
ops on anything but CUDA tensors raises (NotImplementedError from the dispatcher) — the hot path has no fallback
(BASELINE.json north_star: "thin C-ABI torch custom-op layer ... no multi-backend dispatch and no CPU fallback").
`bya_b200.ops` holds the argument checking and the Python-friendly signatures and calls these ops.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .lib import ByaChainArgs, ByaDpmStepArgs, ByaGemmArgs, check, lib

_LIB = torch.library.Library("bya", "DEF")
OP_NAMES = []


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _op(schema: str):
    """Defines `bya::<schema>` and registers the decorated function as its CUDA kernel."""
    name = schema.split("(", 1)[0]

    def deco(fn):
        _LIB.define(schema)
        _LIB.impl(name, fn, "CUDA")
        OP_NAMES.append(name)
        return fn

    return deco


@_op("gemm_bf16(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias, int act, int mode, Tensor? resid, Tensor? gate_a, "
     "Tensor? gate_b, int split_row, float alpha, Tensor? row_bias_scale, int qkv_block, float ln_eps, Tensor? rope_cos, "
     "Tensor? rope_sin, int rope_row0, Tensor? nq_w, Tensor? nq_b, Tensor? nk_w, Tensor? nk_b, int group_m, int col_block, "
     "int col_block_stride, int a_kblock, int a_kblock_stride, float q_premul, int split_k, Tensor[]? peer_out, "
     "Tensor? rope_cs, Tensor? rope_mismatch) -> ()")
def _gemm_bf16(a, w, out, bias, act, mode, resid, gate_a, gate_b, split_row, alpha, row_bias_scale, qkv_block, ln_eps,
               rope_cos, rope_sin, rope_row0, nq_w, nq_b, nk_w, nk_b, group_m, col_block, col_block_stride, a_kblock,
               a_kblock_stride, q_premul, split_k, peer_out, rope_cs, rope_mismatch):
    g = ByaGemmArgs()
    g.M, g.N, g.K = a.shape[0], w.shape[0], w.shape[1]
    g.mode, g.act, g.group_m = mode, act, group_m
    g.bias, g.out, g.ldc = _ptr(bias), _ptr(out), out.stride(-2)
    if resid is not None:
        g.resid, g.ldr = _ptr(resid), resid.stride(0)
    g.gate_a, g.gate_b, g.row_bias_scale = _ptr(gate_a), _ptr(gate_b), _ptr(row_bias_scale)
    g.split_row, g.alpha = split_row, alpha
    g.qkv_block, g.ln_eps = qkv_block, ln_eps
    g.rope_cos, g.rope_sin, g.rope_row0 = _ptr(rope_cos), _ptr(rope_sin), rope_row0
    g.nq_w, g.nq_b, g.nk_w, g.nk_b = _ptr(nq_w), _ptr(nq_b), _ptr(nk_w), _ptr(nk_b)
    g.col_block, g.col_block_stride = col_block, col_block_stride
    g.a_kblock, g.a_kblock_stride = a_kblock, a_kblock_stride
    g.q_premul = q_premul
    g.split_k = split_k
    g.rope_cs, g.rope_mismatch = _ptr(rope_cs), _ptr(rope_mismatch)
    if peer_out:   # push exchange: column block d goes to peer_out[d] (another GPU's memory), see include/bya.h
        for d, t in enumerate(peer_out):
            g.peer_out[d] = t.data_ptr()
    check(lib().bya_gemm_bf16(_stream(), _ptr(a), a.stride(0), _ptr(w), w.stride(0), ctypes.byref(g)), "gemm")


@_op("gemm_ln_gemm_bf16(Tensor a1, Tensor w1, Tensor? b1, Tensor resid, Tensor(a!) x_out, int store_x, Tensor w2f, "
     "Tensor csum, Tensor b2, float ln_eps, int act, Tensor(b!) out2, int n_split, int col_block, int col_block_stride, "
     "int a_kblock, int a_kblock_stride) -> ()")
def _gemm_ln_gemm_bf16(a1, w1, b1, resid, x_out, store_x, w2f, csum, b2, ln_eps, act, out2, n_split, col_block,
                       col_block_stride, a_kblock, a_kblock_stride):
    g = ByaChainArgs()
    g.M, g.N2, g.act, g.n_split = a1.shape[0], w2f.shape[0], act, n_split
    g.b1, g.resid, g.ldr = _ptr(b1), _ptr(resid), resid.stride(0)
    g.x_out, g.ldx, g.store_x, g.ln_eps = _ptr(x_out), x_out.stride(0), store_x, ln_eps
    g.csum, g.b2, g.out2, g.ldc = _ptr(csum), _ptr(b2), _ptr(out2), out2.stride(-2)
    g.col_block, g.col_block_stride = col_block, col_block_stride
    g.a_kblock, g.a_kblock_stride = a_kblock, a_kblock_stride
    check(lib().bya_gemm_ln_gemm_bf16(_stream(), _ptr(a1), a1.stride(0), _ptr(w1), w1.stride(0), _ptr(w2f), w2f.stride(0),
                                      ctypes.byref(g)), "gemm_ln_gemm")


@_op("attention_d64(Tensor q, Tensor k, Tensor v, Tensor(a!) out, int batch, int seq, int seq_stride, int heads, "
     "float scale, float score_bound_log2, int variant) -> ()")
def _attention_d64(q, k, v, out, batch, seq, seq_stride, heads, scale, score_bound_log2, variant):
    """variant 0: bya_attention_d64, 1: _strided, 2: _bounded."""
    L = lib()
    if variant == 2:
        rc = L.bya_attention_d64_bounded(_stream(), _ptr(q), _ptr(k), _ptr(v), q.stride(0), _ptr(out), out.stride(0), batch, seq,
                                         heads, ctypes.c_float(score_bound_log2))
    elif variant == 1:
        rc = L.bya_attention_d64_strided(_stream(), _ptr(q), _ptr(k), _ptr(v), q.stride(0), _ptr(out), out.stride(0), batch, seq,
                                         seq_stride, heads, ctypes.c_float(scale))
    else:
        rc = L.bya_attention_d64(_stream(), _ptr(q), _ptr(k), _ptr(v), q.stride(0), _ptr(out), out.stride(0), batch, seq, heads,
                                 ctypes.c_float(scale))
    check(rc, "attention_d64")


@_op("attention_d64_scatter(Tensor q, Tensor k, Tensor v, Tensor[] out_peers, int rows_per_peer, int seq, int heads, float scale, "
     "float score_bound_log2) -> ()")
def _attention_d64_scatter(q, k, v, out_peers, rows_per_peer, seq, heads, scale, score_bound_log2):
    n = len(out_peers)
    arr = (ctypes.c_void_p * n)(*[t.data_ptr() for t in out_peers])
    rc = lib().bya_attention_d64_scatter(_stream(), _ptr(q), _ptr(k), _ptr(v), q.stride(0), arr, n, rows_per_peer,
                                         out_peers[0].stride(0), seq, heads, ctypes.c_float(scale), ctypes.c_float(score_bound_log2))
    check(rc, "attention_d64_scatter")


@_op("layernorm_modulate(Tensor x, Tensor(a!) out, float eps, Tensor? gamma, Tensor? beta, Tensor? scale_a, Tensor? shift_a, "
     "Tensor? scale_b, Tensor? shift_b, int split_row, Tensor? add) -> ()")
def _layernorm_modulate(x, out, eps, gamma, beta, scale_a, shift_a, scale_b, shift_b, split_row, add):
    rows, dim = x.shape
    rc = lib().bya_layernorm_modulate(_stream(), _ptr(x), x.stride(0), _ptr(out), out.stride(0), rows, dim, ctypes.c_float(eps),
                                      _ptr(gamma), _ptr(beta), _ptr(scale_a), _ptr(shift_a), _ptr(scale_b), _ptr(shift_b),
                                      split_row, _ptr(add), 0 if add is None else add.shape[0])
    check(rc, "layernorm_modulate")


@_op("gemv(Tensor w, Tensor? bias, Tensor x, Tensor(a!) y, int in_act, int out_act) -> ()")
def _gemv(w, bias, x, y, in_act, out_act):
    B, K = x.shape
    check(lib().bya_gemv(_stream(), _ptr(w), _ptr(bias), _ptr(x), _ptr(y), B, w.shape[0], K, in_act, out_act), "gemv")


@_op("rope_pack(Tensor cos, Tensor sin, Tensor(a!) packed, Tensor(b!) mismatch) -> ()")
def _rope_pack(cos, sin, packed, mismatch):
    check(lib().bya_rope_pack(_stream(), _ptr(cos), _ptr(sin), _ptr(packed), _ptr(mismatch), cos.shape[0]), "rope_pack")


@_op("timestep_features(Tensor t, Tensor(a!) out) -> ()")
def _timestep_features(t, out):
    check(lib().bya_timestep_features(_stream(), _ptr(t), _ptr(out), out.shape[0], out.shape[1]), "timestep_features")


@_op("patchify(Tensor latents, Tensor(a!) out) -> ()")
def _patchify(latents, out):
    F, C, H, W = latents.shape
    check(lib().bya_patchify(_stream(), _ptr(latents), _ptr(out), F, C, H, W, out.stride(0)), "patchify")


@_op("unpatchify(Tensor y, Tensor(a!) out) -> ()")
def _unpatchify(y, out):
    F, C, H, W = out.shape
    check(lib().bya_unpatchify(_stream(), _ptr(y), y.stride(0), _ptr(out), F, C, H // 2, W // 2), "unpatchify")


@_op("router_head(Tensor x, Tensor w, Tensor b, Tensor(a!) r, int rows, int chars) -> ()")
def _router_head(x, w, b, r, rows, chars):
    check(lib().bya_router_head(_stream(), _ptr(x), _ptr(w), _ptr(b), _ptr(r), rows, chars, x.shape[1]), "router_head")


@_op("xattn_kv32(Tensor q, Tensor K, Tensor Vt, Tensor? w, Tensor(a!) out, int heads, int head_dim, int chars, int kv_frames, "
     "float scale, int tok_begin, int total_tokens) -> ()")
def _xattn_kv32(q, K, Vt, w, out, heads, head_dim, chars, kv_frames, scale, tok_begin, total_tokens):
    rc = lib().bya_xattn_kv32(_stream(), _ptr(q), q.stride(0), _ptr(K), _ptr(Vt), _ptr(w), _ptr(out), out.stride(0), q.shape[0],
                              heads, head_dim, chars, kv_frames, ctypes.c_float(scale), ctypes.c_longlong(tok_begin),
                              ctypes.c_longlong(total_tokens))
    check(rc, "xattn_kv32")


@_op("small_attention(Tensor qkv, Tensor(a!) out, int n_seq, int seq_len, int heads, int inner, int outer_stride, "
     "int tok_stride, float scale) -> ()")
def _small_attention(qkv, out, n_seq, seq_len, heads, inner, outer_stride, tok_stride, scale):
    rc = lib().bya_small_attention(_stream(), _ptr(qkv), qkv.stride(0), _ptr(out), out.stride(0), n_seq, seq_len, heads, inner,
                                   ctypes.c_longlong(outer_stride), ctypes.c_longlong(tok_stride), ctypes.c_float(scale))
    check(rc, "small_attention")


@_op("masks_to_routing(Tensor masks, int frames, int grid_h, int grid_w, Tensor(a!) index_mask, Tensor(b!) logits) -> ()")
def _masks_to_routing(masks, frames, grid_h, grid_w, index_mask, logits):
    C, T, H, W = masks.shape
    rc = lib().bya_masks_to_routing(_stream(), _ptr(masks), C, T, H, W, frames, grid_h, grid_w, _ptr(index_mask), _ptr(logits))
    check(rc, "masks_to_routing")


@_op("routing_frame_or(Tensor logits, Tensor(a!) out, int frames) -> ()")
def _routing_frame_or(logits, out, frames):
    n, C = logits.shape
    check(lib().bya_routing_frame_or(_stream(), _ptr(logits), _ptr(out), frames, n // frames, C), "routing_frame_or")


@_op("audio_weights(Tensor af, Tensor routing, Tensor(a!) w, Tensor(b!)? wsum) -> ()")
def _audio_weights(af, routing, w, wsum):
    n, C = routing.shape
    check(lib().bya_audio_weights(_stream(), _ptr(af), _ptr(routing), _ptr(w), _ptr(wsum), n, C), "audio_weights")


@_op("cfg_dpm_step(Tensor model_out, Tensor sample, Tensor(a!) prev_sample, Tensor old_pred, Tensor(b!) pred_out, Tensor noise, "
     "Tensor coef, int prediction_type, Tensor? step_index, Tensor(c!)? model_input) -> ()")
def _cfg_dpm_step(model_out, sample, prev_sample, old_pred, pred_out, noise, coef, prediction_type, step_index, model_input):
    F, C, H, W = sample.shape[-4:]
    n = F * C * H * W
    a = ByaDpmStepArgs()
    a.frames, a.channels, a.hw, a.prediction_type = F, C, H * W, prediction_type
    if model_out.dtype == torch.float32:
        a.cfg_batch, a.model_out, a.model_out_f32 = 1, None, model_out.data_ptr()
    else:
        a.cfg_batch, a.model_out, a.model_out_f32 = model_out.numel() // n, model_out.data_ptr(), None
    a.sample, a.prev_sample = sample.data_ptr(), prev_sample.data_ptr()
    a.old_pred, a.pred_out, a.noise, a.coef = old_pred.data_ptr(), pred_out.data_ptr(), noise.data_ptr(), coef.data_ptr()
    if step_index is not None:
        a.step_index = step_index.data_ptr()
    if model_input is not None:
        a.model_input, a.in_batch, a.in_channels = model_input.data_ptr(), model_input.shape[0], model_input.shape[2]
    check(lib().bya_cfg_dpm_step(_stream(), ctypes.byref(a)), "cfg_dpm_step")


@_op("denoise_select_step(Tensor timesteps, Tensor(a!) timestep_out, Tensor(b!) counter, Tensor(c!) step_index) -> ()")
def _denoise_select_step(timesteps, timestep_out, counter, step_index):
    check(lib().bya_denoise_select_step(_stream(), _ptr(timesteps), timesteps.numel(), _ptr(timestep_out), timestep_out.numel(),
                                        _ptr(counter), _ptr(step_index)), "denoise_select_step")


# ---------------------------------------------------------------- per-generation prologue helpers
@_op("copy2d(Tensor src, Tensor(a!) out) -> ()")
def _copy2d(src, out):
    rows, cols = out.shape
    check(lib().bya_copy2d(_stream(), _ptr(src), ctypes.c_longlong(src.stride(0)), int(src.dtype == torch.float32), _ptr(out),
                           ctypes.c_longlong(out.stride(0)), rows, cols), "copy2d")


@_op("memset_zero(Tensor(a!) t) -> ()")
def _memset_zero(t):
    check(lib().bya_memset_zero(_stream(), _ptr(t), ctypes.c_longlong(t.numel() * t.element_size())), "memset_zero")


@_op("splitk_finalize(Tensor ws, int row0, Tensor? bias, int act, Tensor(a!) out) -> ()")
def _splitk_finalize(ws, row0, bias, act, out):
    rows, cols = out.shape
    splits, ws_rows, _ = ws.shape
    check(lib().bya_splitk_finalize(_stream(), _ptr(ws), splits, ws_rows, ctypes.c_longlong(ws.stride(1)), row0, _ptr(bias), act,
                                    _ptr(out), ctypes.c_longlong(out.stride(0)), rows, cols), "splitk_finalize")


@_op("layernorm_leakyrelu(Tensor x, Tensor(a!) out, float eps, Tensor gamma, Tensor beta, float slope) -> ()")
def _layernorm_leakyrelu(x, out, eps, gamma, beta, slope):
    rows, dim = x.shape
    check(lib().bya_layernorm_leakyrelu(_stream(), _ptr(x), x.stride(0), _ptr(out), out.stride(0), rows, dim, ctypes.c_float(eps),
                                        _ptr(gamma), _ptr(beta), ctypes.c_float(slope)), "layernorm_leakyrelu")


@_op("kv_pack(Tensor x, int k_off, int v_off, Tensor(a!) K, Tensor(b!) Vt) -> ()")
def _kv_pack(x, k_off, v_off, K, Vt):
    G, H, _, d = K.shape
    check(lib().bya_kv_pack(_stream(), _ptr(x), ctypes.c_longlong(x.stride(0)), k_off, v_off, _ptr(K), _ptr(Vt), G, H, d), "kv_pack")


@_op("router_keys_scatter(Tensor k, Tensor(a!) mat, int chars, int heads, int head_dim) -> ()")
def _router_keys_scatter(k, mat, chars, heads, head_dim):
    check(lib().bya_router_keys_scatter(_stream(), _ptr(k), ctypes.c_longlong(k.stride(0)), _ptr(mat), chars, heads, head_dim),
          "router_keys_scatter")


# ---------------------------------------------------------------- exchanges over NVLink peer memory
@_op("peer_barrier(Tensor(a!) counter, Tensor flag_ptrs, int my_rank, int n_ranks) -> ()")
def _peer_barrier(counter, flag_ptrs, my_rank, n_ranks):
    check(lib().bya_peer_barrier(_stream(), _ptr(counter), _ptr(flag_ptrs), my_rank, n_ranks), "peer_barrier")


@_op("peer_copy(Tensor segs, int n_segs, Tensor peer_ptrs, Tensor(a!) local, int push, int vec_bytes, int blocks_per_seg) -> ()")
def _peer_copy(segs, n_segs, peer_ptrs, local, push, vec_bytes, blocks_per_seg):
    check(lib().bya_peer_copy(_stream(), _ptr(segs), n_segs, _ptr(peer_ptrs), _ptr(local), push, vec_bytes, blocks_per_seg), "peer_copy")
