"""ctypes binding of libs2t_b200.so (C ABI in include/s2t_b200.h).

This is the stub a maintainer of the reference would add next to
``model/joiner/joiner.py`` in place of ``import k2``.
There is no CPU fallback: if the shared library cannot be loaded, importing
this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_size_t, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libs2t_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT_RELU, ACT_TANH = 0, 1
MODE_FP32_SIMT, MODE_BF16_TC = 0, 1

_DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

P = c_void_p
I = c_int
F = c_float

_SIGNATURES = {
    "s2t_abi_version": (c_int, []),
    "s2t_last_error": (ctypes.c_char_p, []),
    "s2t_launch_count": (ctypes.c_longlong, []),
    "s2t_profile_enable": (None, [I]),
    "s2t_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
    "s2t_lattice_workspace_bytes": (c_size_t, [I, I, I, I]),
    "s2t_mutual_information": (c_int, [P, P, P, I, I, I, P, P, P, P, P]),
    "s2t_simple_workspace_bytes": (c_size_t, [I, I, I, I, I]),
    "s2t_simple_loss_fwd": (c_int, [I, P, P, P, P, I, I, I, I, I, F, F, P, P, P, P, P, P, P, P, P, P, I, P]),
    "s2t_simple_loss_bwd": (c_int, [I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, F, P, P, P, P]),
    "s2t_prune_ranges": (c_int, [P, P, P, I, I, I, I, I, P, P]),
    "s2t_logits_loss_fwd": (c_int, [P, I, P, P, P, I, I, I, I, I, I, F, P, P, P, P, P, P, P, P]),
    "s2t_logits_loss_bwd": (c_int, [P, I, P, P, P, P, P, P, I, I, I, I, I, I, F, P, P]),
    "s2t_joiner_workspace_bytes": (c_size_t, [I, I, I, I, I, I]),
    "s2t_joiner_loss_fwd": (c_int, [I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, F, P, P, P, P, P, P, P,
                                    P, P]),
    "s2t_joiner_logprobs_fwd": (c_int, [I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, F, P, P, P, P, P]),
    "s2t_band_lattice_fwd": (c_int, [P, P, P, P, I, I, I, I, P, P, P, P, P]),
    "s2t_rnnt_beam_workspace_bytes": (c_size_t, [I, I, I, I]),
    "s2t_rnnt_beam_decode": (c_int, [P] * 12 + [I] * 12 + [P, P, P, P, P]),
    "s2t_simple_loss_prep_lm": (c_int, [I, P, P, P, I, I, I, I, I, P, P]),
    "s2t_weighted_sum": (c_int, [P, P, I, P, P]),
    "s2t_rescale_groups": (c_int, [P, ctypes.c_int64, P, ctypes.c_int64, I, P, P, P]),
    "s2t_joiner_loss_bwd": (c_int, [I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, F, P, P, P, P, P, P, P,
                                    P, P, P, P, P]),
    "s2t_linear_workspace_bytes": (c_size_t, [ctypes.c_int64, I, I]),
    "s2t_linear_fwd": (c_int, [P, I, P, P, ctypes.c_int64, I, I, P, P, P, P]),
    "s2t_linear_bwd": (c_int, [P, P, P, ctypes.c_int64, I, I, P, P, I, P, P, P]),
    "s2t_ctc_workspace_bytes": (c_size_t, [I, I, I, I]),
    "s2t_ctc_loss_fwd": (c_int, [P, P, P, P, I, I, I, I, I, P, P, P, P]),
    "s2t_ctc_loss_bwd": (c_int, [P, P, P, P, I, I, I, I, I, P, P, P, P, I, P, P]),
    "s2t_predictor_embed_conv_fwd": (c_int, [P, P, P, I, I, I, I, I, P, P]),
    "s2t_predictor_embed_conv_bwd": (c_int, [P, P, P, P, I, I, I, I, I, P, P, P]),
    "s2t_rnnt_greedy_decode": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, P, P, P]),
    "s2t_joiner_materialize": (c_int, [I, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P]),
}

_lib: Optional[ctypes.CDLL] = None


class S2TError(RuntimeError):
    pass


def exported_symbols():
    """Names include/s2t_b200.h declares (kept in sync by tests/test_abi.py)."""
    return list(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S2TError(
                f"{LIB_PATH} is missing: build it with `python -m speech2text_b200.build` "
                "(nvcc, sm_100a). speech2text_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.s2t_abi_version() != 3:
            raise S2TError("libs2t_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def launch_count() -> int:
    return int(lib().s2t_launch_count())


_profiling = False


def profile_enable(on: bool) -> None:
    global _profiling
    _profiling = bool(on)
    lib().s2t_profile_enable(1 if on else 0)


def profiling() -> bool:
    """True while the per-kernel event timer is on: the host side then keeps every kernel on one stream (exclusive
    kernel times), as the library does for its own fork/join."""
    return _profiling


def profile_report() -> dict:
    """{kernel name: (launch groups, total device ms)} recorded since profile_enable(True)."""
    buf = ctypes.create_string_buffer(1 << 16)
    lib().s2t_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out


def check(rc: int) -> None:
    if rc != 0:
        raise S2TError(lib().s2t_last_error().decode("utf-8", "replace"))


def ptr(t: Optional[torch.Tensor]) -> c_void_p:
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise S2TError("speech2text_b200 kernels need CUDA tensors (no CPU path)")
    if not t.is_contiguous():
        raise S2TError("internal: non-contiguous tensor handed to the C ABI")
    return c_void_p(t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise S2TError(f"unsupported logits dtype {dt}") from None
