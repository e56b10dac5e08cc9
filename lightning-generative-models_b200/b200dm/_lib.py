"""ctypes binding of libb200dm.so (the C ABI declared in include/b200dm.h).

There is no CPU fallback: if the shared library is missing or a call returns an error code the
caller gets an exception.  Loading the library does not need a GPU (CPU tests check the symbols).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200DM_LIB selects another build of the same ABI (the phase-timing debug build of scripts/phase_timing.py)
LIB_PATH = os.environ.get("B200DM_LIB") or os.path.join(_HERE, "libb200dm.so")

F32, BF16 = 0, 1
PRED_NOISE, PRED_X0, PRED_V = 0, 1, 2
OBJECTIVES = {"pred_noise": PRED_NOISE, "pred_x0": PRED_X0, "pred_v": PRED_V}


class B200dmError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32), ("ksize", C.c_int32), ("impl", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("x", C.c_void_p), ("x_ld", C.c_int32),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("y", C.c_void_p), ("y_ld", C.c_int32),
        ("res", C.c_void_p), ("res_ld", C.c_int32),
        ("accumulate", C.c_int32),
        ("gn_part", C.c_void_p), ("gn_groups", C.c_int32),
    ]


class GnDesc(C.Structure):
    """b200dm_gn_desc: GroupNorm + FiLM + SiLU fused behind a 3x3 conv (b200dm_conv_gn_fwd)."""
    _fields_ = [
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("film", C.c_void_p),
        ("film_ld", C.c_int32), ("groups", C.c_int32), ("eps", C.c_float), ("raw_ld", C.c_int32),
        ("stats", C.c_void_p), ("raw", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32), ("ksize", C.c_int32), ("impl", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("x", C.c_void_p), ("x_ld", C.c_int32),
        ("dy", C.c_void_p), ("dy_ld", C.c_int32),
        ("dw", C.c_void_p),
        ("accumulate", C.c_int32), ("cin_valid", C.c_int32),
        ("s_tap", C.c_int64), ("s_co", C.c_int64), ("s_ci", C.c_int64),
    ]


class ColsumItem(C.Structure):
    """b200dm_colsum_item: one column sum of b200dm_colsum_batched."""
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int64), ("ld", C.c_int32), ("C", C.c_int32)]


class LinAttnBlockDesc(C.Structure):
    """b200dm_linattn_block_desc: the fused inference LinearAttention block (csrc/linattn_tc.cu)."""
    _fields_ = [
        ("B", C.c_int32), ("n", C.c_int32), ("C", C.c_int32), ("x_ld", C.c_int32), ("y_ld", C.c_int32),
        ("reserved", C.c_int32),
        ("x", C.c_void_p), ("y", C.c_void_p), ("wqkv", C.c_void_p), ("wout", C.c_void_p),
        ("bout", C.c_void_p), ("gout", C.c_void_p), ("mem_kv", C.c_void_p), ("ws", C.c_void_p),
    ]


class NoiseDesc(C.Structure):
    """b200dm_noise_desc: inputs of the forward-noising step shared by q_sample and the loss."""
    _fields_ = [
        ("img", C.c_void_p), ("t", C.c_void_p), ("noise", C.c_void_p), ("offset", C.c_void_p),
        ("sqrt_ac", C.c_void_p), ("sqrt_1mac", C.c_void_p),
        ("offset_strength", C.c_float), ("normalize", C.c_int32), ("B", C.c_int32), ("reserved", C.c_int32),
        ("chw", C.c_int64), ("hw", C.c_int64),
        ("seed", C.c_uint64), ("stream_id", C.c_uint64), ("elem_offset", C.c_uint64),
    ]


class PackEntry(C.Structure):
    _fields_ = [
        ("w", C.c_void_p), ("wf", C.c_void_p), ("wt", C.c_void_p),
        ("taps", C.c_int32), ("Cout", C.c_int32), ("Cin", C.c_int32), ("flip", C.c_int32),
        ("s_tap", C.c_int64), ("s_co", C.c_int64), ("s_ci", C.c_int64),
        ("tile_begin", C.c_int32), ("tiles_ci", C.c_int32), ("tiles_co", C.c_int32), ("reserved", C.c_int32),
    ]


_P, _I, _L, _F, _U = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64

# name -> argtypes (restype is int unless listed in _SPECIAL)
PROTOTYPES = {
    "b200dm_q_sample": [C.POINTER(NoiseDesc), _P, _P, _P, _P],
    "b200dm_loss_fwd_bwd": [C.POINTER(NoiseDesc), _P, _P, _P, _P, _I, _P],
    "b200dm_ddim_step": [_P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _F, _I, _I, _L, _U, _U, _U, _P],
    "b200dm_ddpm_step": [_P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _F, _I, _I, _L, _U, _U, _U, _P],
    "b200dm_randn": [_P, _L, _U, _U, _U, _P],
    "b200dm_unnormalize": [_P, _P, _L, _P],
    "b200dm_conv_fwd": [C.POINTER(ConvDesc), _P],
    "b200dm_conv_gn_fwd": [C.POINTER(ConvDesc), C.POINTER(GnDesc), _P],
    "b200dm_conv_wgrad": [C.POINTER(WgradDesc), _P],
    "b200dm_colsum": [_I, _P, _I, _L, _I, _P, _I, _P],
    "b200dm_colsum_batched": [_I, _P, _I, _P],
    "b200dm_concat2_nchw": [_P, _P, _P, _I, _L, _L, _P],
    "b200dm_stem7_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200dm_pack_stem_rows": [_P, _P, _I, _I, _I, _P],
    "b200dm_pack_linattn_qkv": [_P, _P, _P, _I, _P],
    "b200dm_linattn_block_fwd": [_P, _P],
    "b200dm_im2col7": [_P, _P, _I, _I, _I, _I, _I, _P],
    "b200dm_pack_stem_weight": [_P, _P, _I, _I, _I, _P],
    "b200dm_pack_upconv_weight": [_P, _P, _I, _I, _P],
    "b200dm_init_conv_fwd": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200dm_init_conv_wgrad": [_I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "b200dm_final_conv_fwd": [_I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200dm_final_conv_bwd": [_I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "b200dm_upsample2x_fwd": [_I, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "b200dm_upsample2x_bwd": [_I, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "b200dm_gn_stats": [_I, _P, _I, _P, _I, _I, _I, _I, _F, _P],
    "b200dm_gn_fwd": [_I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P],
    "b200dm_gn_fwd_pre": [_I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P],
    "b200dm_gn_apply_fwd": [_I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "b200dm_gn_apply_bwd": [_I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P,
                            _I, _I, _I, _I, _P],
    "b200dm_rmsnorm_fwd": [_I, _P, _I, _P, _P, _I, _P, _I, _L, _I, _P],
    "b200dm_rmsnorm_bwd": [_I, _P, _I, _P, _I, _P, _P, _I, _P, _I, _P, _L, _I, _P],
    "b200dm_linattn_fwd": [_I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P],
    "b200dm_linattn_bwd": [_I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _P, _I, _I, _P],
    "b200dm_attn_fwd": [_I, _P, _I, _P, _P, _I, _I, _I, _P],
    "b200dm_attn_bwd": [_I, _P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _P],
    "b200dm_sinusoidal": [_P, _P, _I, _I, _F, _P],
    "b200dm_linear_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200dm_linear_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200dm_linear_bwd_cols": [_P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    "b200dm_pack_conv_weight": [_I, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _P],
    "b200dm_pack_conv_weights_batched": [_I, _P, _I, _I, _P],
    "b200dm_pack_conv_weights_range": [_I, _P, _I, _I, _I, _P],
    "b200dm_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "b200dm_adam_step_bg": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _I, _P],
    "b200dm_ema_update": [_P, _P, _L, _F, _P],
    "b200dm_fill_f32": [_P, _L, _F, _P],
    "b200dm_cast_f32_bf16": [_P, _P, _L, _P],
    "b200dm_cast_bf16_f32": [_P, _P, _L, _P],
    "b200dm_debug_umma_rate": [_I, _I, _I, _I, _P, _P],
}
_SPECIAL = {
    "b200dm_version": ([], C.c_int),
    "b200dm_last_error": ([], C.c_char_p),
    "b200dm_launch_count": ([], C.c_int64),
    "b200dm_gn_bwd_ws_floats": ([_I, _I, _I], C.c_int64),
    "b200dm_reset_launch_count": ([], None),
    "b200dm_tc_available": ([], C.c_int),
    "b200dm_set_reserved_sms": ([C.c_int32], C.c_int),
    "b200dm_conv_gn_supported": ([C.POINTER(ConvDesc), C.POINTER(GnDesc)], C.c_int),
    "b200dm_linattn_block_ws_floats": ([_I, _I, _I], C.c_int64),
    "b200dm_linattn_block_supported": ([_P], C.c_int),
    "b200dm_stem7_supported": ([_I, _I, _I, _I, _I, _I], C.c_int),
}
ALL_SYMBOLS = sorted(list(PROTOTYPES) + list(_SPECIAL))

_lib = None


def load():
    """Load (once) and return the ctypes library handle.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B200dmError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lightning-generative-models_b200/csrc`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200dm_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise B200dmError(f"libb200dm {what} failed with code {rc}: {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke an entry point on torch's current CUDA stream and raise on a non-zero return code."""
    lib = load()
    rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        raise B200dmError(f"{name} failed with code {rc}: {last_error()}")
