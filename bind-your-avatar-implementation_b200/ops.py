"""Thin Python wrappers over the C ABI (`include/bya.h`): tensor checks -> raw pointers -> libbya.so.
torch is used only for device memory and the current stream."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .lib import ByaGemmArgs, check, lib

EPI_STORE, EPI_RESIDUAL, EPI_QKV = 0, 1, 2
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


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
         qk_cols=0, ln_eps=1e-6, rope=None, nq=None, nk=None, group_m=0) -> torch.Tensor:
    """out = epilogue(a @ w.T); a [M,K], w [N,K], out [M,N] (row strides may exceed the width)."""
    global LAUNCHES
    _bf16_2d(a, "a"), _bf16_2d(w, "w"), _bf16_2d(out, "out")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or out.shape[0] != M or out.shape[1] != N:
        raise RuntimeError(f"bya_b200.gemm: shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    args = ByaGemmArgs()
    args.M, args.N, args.K = M, N, K
    args.mode, args.act, args.group_m = mode, act, group_m
    args.bias = _ptr(bias)
    args.out, args.ldc = _ptr(out), out.stride(0)
    if mode == EPI_RESIDUAL:
        _bf16_2d(resid, "resid")
        args.resid, args.ldr = _ptr(resid), resid.stride(0)
        args.gate_a, args.gate_b = _ptr(gate_a), _ptr(gate_b)
        args.row_bias_scale = _ptr(row_bias_scale)
    args.split_row = split_row
    args.alpha = alpha
    if mode == EPI_QKV:
        cos, sin = rope
        args.qk_cols, args.ln_eps = qk_cols, ln_eps
        args.rope_cos, args.rope_sin = _ptr(cos), _ptr(sin)
        args.nq_w, args.nq_b, args.nk_w, args.nk_b = _ptr(nq[0]), _ptr(nq[1]), _ptr(nk[0]), _ptr(nk[1])
    rc = lib().bya_gemm_bf16(_stream(), _ptr(a), a.stride(0), _ptr(w), w.stride(0), ctypes.byref(args))
    check(rc, "gemm")
    LAUNCHES += 1
    return out


def attention_d64(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, batch: int, seq: int,
                  heads: int, scale: float = 0.125) -> torch.Tensor:
    """q/k/v/out: [batch*seq, heads*64] column-slice views (shared row stride for q,k,v) of bf16 matrices."""
    global LAUNCHES
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _bf16_2d(t, n)
        if t.shape[0] != batch * seq or t.shape[1] != heads * 64:
            raise RuntimeError(f"bya_b200.attention_d64: {n} has shape {tuple(t.shape)}")
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise RuntimeError("bya_b200.attention_d64: q, k, v must share a row stride")
    rc = lib().bya_attention_d64(_stream(), _ptr(q), _ptr(k), _ptr(v), q.stride(0), _ptr(out), out.stride(0),
                                 batch, seq, heads, ctypes.c_float(scale))
    check(rc, "attention_d64")
    LAUNCHES += 1
    return out
