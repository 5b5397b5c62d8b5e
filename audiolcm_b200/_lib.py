"""ctypes binding of csrc/libaudiolcm_b200.so (the C-ABI in include/audiolcm_b200.h).

There is no fallback: if the library is missing, cannot be loaded, or no sm_100 GPU is present,
the product raises.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libaudiolcm_b200.so")

PREC = {"fp32": 0, "tf32": 1, "bf16": 2, "fp16": 3}
CLASSES = ("conv", "act", "norm", "attn", "misc")


class BigVGANCfg(C.Structure):
    _fields_ = [
        ("num_mels", C.c_int),
        ("upsample_initial_channel", C.c_int),
        ("num_upsamples", C.c_int),
        ("num_kernels", C.c_int),
        ("upsample_rates", C.c_int * 8),
        ("upsample_kernel_sizes", C.c_int * 8),
        ("resblock_kernel_sizes", C.c_int * 4),
        ("resblock_dilation_sizes", (C.c_int * 3) * 4),
        ("resblock2", C.c_int),
        ("snake_linear", C.c_int),
    ]


class VAECfg(C.Structure):
    _fields_ = [
        ("ch", C.c_int),
        ("out_ch", C.c_int),
        ("z_channels", C.c_int),
        ("embed_dim", C.c_int),
        ("kernel_size", C.c_int),
        ("num_res_blocks", C.c_int),
        ("n_levels", C.c_int),
        ("ch_mult", C.c_int * 8),
        ("upsample_levels", C.c_int * 8),
        ("attn_levels", C.c_int * 8),
    ]


class VAEEncCfg(C.Structure):
    _fields_ = [
        ("ch", C.c_int),
        ("in_channels", C.c_int),
        ("z_channels", C.c_int),
        ("embed_dim", C.c_int),
        ("kernel_size", C.c_int),
        ("num_res_blocks", C.c_int),
        ("n_levels", C.c_int),
        ("double_z", C.c_int),
        ("ch_mult", C.c_int * 8),
        ("downsample_levels", C.c_int * 8),
        ("attn_levels", C.c_int * 8),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("ms", C.c_double * 5),
        ("flops", C.c_double * 5),
        ("bytes", C.c_double * 5),
        ("launches", C.c_int * 5),
    ]


_P = C.c_void_p
_FP = C.c_void_p  # device float* passed as integer address

# name -> (restype, argtypes); every symbol include/audiolcm_b200.h declares
SYMBOLS = {
    "alcm_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "alcm_ctx_destroy": (None, [_P]),
    "alcm_last_error": (C.c_char_p, []),
    "alcm_vocoder_num_tensors": (C.c_int, [C.POINTER(BigVGANCfg)]),
    "alcm_vocoder_create": (C.c_int, [_P, C.POINTER(BigVGANCfg), C.POINTER(_FP), C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_vocoder_destroy": (None, [_P]),
    "alcm_vocode": (C.c_int, [_P, _FP, C.c_int, C.c_int, _FP, _P]),
    "alcm_vocode_pcm16": (C.c_int, [_P, _FP, C.c_int, C.c_int, _FP, _P]),
    "alcm_vocoder_plan": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "alcm_vocoder_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "alcm_vae_num_tensors": (C.c_int, [C.POINTER(VAECfg)]),
    "alcm_vae_create": (C.c_int, [_P, C.POINTER(VAECfg), C.POINTER(_FP), C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_vae_destroy": (None, [_P]),
    "alcm_vae_decode": (C.c_int, [_P, _FP, C.c_int, C.c_int, C.c_float, _FP, _P]),
    "alcm_vae_plan": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "alcm_vae_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "alcm_vae_encoder_num_tensors": (C.c_int, [C.POINTER(VAEEncCfg)]),
    "alcm_vae_encoder_create": (C.c_int, [_P, C.POINTER(VAEEncCfg), C.POINTER(_FP), C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_vae_encoder_destroy": (None, [_P]),
    "alcm_vae_encode": (C.c_int, [_P, _FP, C.c_int, C.c_int, _FP, _P]),
    "alcm_decode_to_wav": (C.c_int, [_P, _P, _FP, C.c_int, C.c_int, C.c_float, _FP, _FP, _P]),
    "alcm_decode_to_pcm16": (C.c_int, [_P, _P, _FP, C.c_int, C.c_int, C.c_float, _FP, _FP, _P]),
    "alcm_conv1d_create": (C.c_int, [_P, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_conv1d_destroy": (None, [_P]),
    "alcm_conv1d_run": (C.c_int, [_P, _FP, _FP, _FP, C.c_int, C.c_int, _P]),
    "alcm_melspec_create": (C.c_int, [_P, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_melspec_destroy": (None, [_P]),
    "alcm_melspec_run": (C.c_int, [_P, _FP, _FP, C.c_int, C.c_int, _P]),
    "alcm_layernorm_cf": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "alcm_ffn1d_create": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "alcm_ffn1d_destroy": (None, [_P]),
    "alcm_ffn1d_run": (C.c_int, [_P, _FP, _FP, _FP, C.c_int, C.c_int, _P]),
    "alcm_lcm_step": (C.c_int, [_P, _FP, _FP, _FP, _FP, _FP, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "alcm_activation1d_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "alcm_conv1d_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "alcm_conv_transpose1d_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "alcm_upsample_conv3_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "alcm_groupnorm_swish_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "alcm_attn1d_fwd": (C.c_int, [_P, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "alcm_profile_decode": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(Profile), _P]),
    "alcm_profile_stages": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(Profile), C.c_int, _P]),
    "alcm_bench_conv": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "alcm_bench_act": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "alcm_vocoder_check_guards": (C.c_int, [_P, C.POINTER(C.c_longlong)]),
    "alcm_vae_check_guards": (C.c_int, [_P, C.POINTER(C.c_longlong)]),
    "alcm_vocoder_launches": (C.c_int, [_P, C.c_int, C.c_int]),
    "alcm_vae_launches": (C.c_int, [_P, C.c_int, C.c_int]),
}

_lib = None
_lock = threading.RLock()  # re-entrant: check() calls load() while ctx() holds the lock
_ctxs: dict[tuple, int] = {}


class AlcmError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every symbol.  Raises if it is missing (no fallback)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AlcmError(
                f"{LIB_PATH} is missing: build it with `python -m audiolcm_b200.build` "
                "(audiolcm_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        msg = load().alcm_last_error().decode("utf-8", "replace")
        raise AlcmError(f"audiolcm_b200 error {rc}: {msg}")


def ctx(device_index: int) -> int:
    """One alcm_ctx per (CUDA device, Python thread), created on first use.  The library keeps no
    process-global mutable state, so threads with their own ctx and their own model handles run
    concurrently (ctypes releases the GIL for the duration of a call)."""
    lib = load()
    key = (int(device_index), threading.get_ident())
    with _lock:
        if key not in _ctxs:
            h = _P()
            check(lib.alcm_ctx_create(C.byref(h), int(device_index)))
            _ctxs[key] = h.value
        return _ctxs[key]


def ptr_array(tensors):
    arr = (_FP * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
