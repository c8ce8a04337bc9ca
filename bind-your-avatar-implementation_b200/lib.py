"""Loads libbya.so (the C-ABI library of sm_100a kernels) with ctypes.  There is NO fallback: if the library is
missing or the device is not a B200, every op raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbya.so")
_lib = None


class ByaGemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("mode", ctypes.c_int), ("act", ctypes.c_int), ("group_m", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("out", ctypes.c_void_p), ("ldc", ctypes.c_int),
        ("resid", ctypes.c_void_p), ("ldr", ctypes.c_int),
        ("gate_a", ctypes.c_void_p), ("gate_b", ctypes.c_void_p), ("split_row", ctypes.c_int),
        ("alpha", ctypes.c_float), ("row_bias_scale", ctypes.c_void_p),
        ("qkv_block", ctypes.c_int), ("ln_eps", ctypes.c_float),
        ("rope_cos", ctypes.c_void_p), ("rope_sin", ctypes.c_void_p), ("rope_row0", ctypes.c_int),
        ("nq_w", ctypes.c_void_p), ("nq_b", ctypes.c_void_p), ("nk_w", ctypes.c_void_p), ("nk_b", ctypes.c_void_p),
        ("col_block", ctypes.c_int), ("col_block_stride", ctypes.c_longlong),
        ("a_kblock", ctypes.c_int), ("a_kblock_stride", ctypes.c_longlong),
        ("q_premul", ctypes.c_float), ("split_k", ctypes.c_int),
        ("peer_out", ctypes.c_void_p * 8),
        ("rope_cs", ctypes.c_void_p), ("rope_mismatch", ctypes.c_void_p),
    ]


class ByaChainArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("N2", ctypes.c_int), ("act", ctypes.c_int), ("n_split", ctypes.c_int),
        ("b1", ctypes.c_void_p), ("resid", ctypes.c_void_p), ("ldr", ctypes.c_int),
        ("x_out", ctypes.c_void_p), ("ldx", ctypes.c_int), ("store_x", ctypes.c_int), ("ln_eps", ctypes.c_float),
        ("csum", ctypes.c_void_p), ("b2", ctypes.c_void_p), ("out2", ctypes.c_void_p), ("ldc", ctypes.c_int),
        ("col_block", ctypes.c_int), ("col_block_stride", ctypes.c_longlong),
        ("a_kblock", ctypes.c_int), ("a_kblock_stride", ctypes.c_longlong),
    ]


class ByaDpmStepArgs(ctypes.Structure):
    _fields_ = [
        ("frames", ctypes.c_int), ("channels", ctypes.c_int), ("hw", ctypes.c_int),
        ("cfg_batch", ctypes.c_int), ("prediction_type", ctypes.c_int),
        ("model_out", ctypes.c_void_p), ("model_out_f32", ctypes.c_void_p),
        ("sample", ctypes.c_void_p), ("prev_sample", ctypes.c_void_p),
        ("old_pred", ctypes.c_void_p), ("pred_out", ctypes.c_void_p),
        ("noise", ctypes.c_void_p), ("model_input", ctypes.c_void_p),
        ("in_batch", ctypes.c_int), ("in_channels", ctypes.c_int),
        ("coef", ctypes.c_void_p), ("step_index", ctypes.c_void_p),
    ]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"bya_b200: {LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.bya_abi_version.restype = ctypes.c_int
        if _lib.bya_abi_version() != 1:
            raise RuntimeError("bya_b200: libbya.so ABI version mismatch; rebuild")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        names = {-1: "bad shape", -2: "bad alignment", -3: "unsupported arch (need sm_100a)", -4: "CUDA launch error",
                 -5: "driver entry point missing"}
        raise RuntimeError(f"bya_b200.{what} failed: {names.get(rc, rc)}")
