"""ctypes binding of libflux_b200.so (include/flux_b200.h).

This is the only place the Python host code touches native code.  Raw ``data_ptr()``s and the
current CUDA stream handle cross the boundary; no torch types do.  A missing library or a failing
call raises -- there is no fallback path of any kind.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
#: FLUXB200_LIB points measurements at an alternative build of the library (A/B runs); the product uses the in-tree one
LIB_PATH = os.environ.get("FLUXB200_LIB") or os.path.join(_HERE, "libflux_b200.so")

E4M3, E5M2 = 0, 1
EPI_PLAIN, EPI_GATE_RESIDUAL, EPI_GELU_QUANT, EPI_QKV_ROPE, EPI_LINEAR1 = 0, 1, 2, 3, 4

ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3

#: every symbol include/flux_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = (
    "fluxb200_version",
    "fluxb200_last_error",
    "fluxb200_device_check",
    "fluxb200_quantize",
    "fluxb200_amax",
    "fluxb200_f8_gemm",
    "fluxb200_f8_gemm_grouped",
    "fluxb200_f8_gemm_ln",
    "fluxb200_f8_gemv",
    "fluxb200_modulation_batched",
    "fluxb200_modulation_batched_bf16",
    "fluxb200_bf16_gemv",
    "fluxb200_timestep_embedding",
    "fluxb200_euler_update",
    "fluxb200_bf16_gemm_small",
    "fluxb200_silu_quant",
    "fluxb200_ln_mod_quant",
    "fluxb200_ln_mod_quant_grouped",
    "fluxb200_qknorm_rope",
    "fluxb200_attention",
    "fluxb200_lora_fuse",
    "fluxb200_conv2d_nhwc",
    "fluxb200_group_norm_nhwc",
    "fluxb200_upsample2x_nhwc",
    "fluxb200_softmax_rows",
    "fluxb200_vae_latent_prep",
    "fluxb200_rows_norm",
    "fluxb200_gated_act",
    "fluxb200_attention_d64",
    "fluxb200_debug_counters",
    "fluxb200_gemm_probe_mode",
    "fluxb200_gemm_force_tiling",
    "fluxb200_fp8_mma_probe",
)


class LnArgs(C.Structure):
    """struct fluxb200_ln_args"""

    _fields_ = [
        ("x", C.c_void_p),
        ("shift", C.c_void_p),
        ("scale", C.c_void_p),
        ("y_fp8", C.c_void_p),
        ("in_scale", C.c_void_p),
        ("ldx", C.c_int64),
        ("ldy", C.c_int64),
        ("mod_batch_stride", C.c_int64),
        ("B", C.c_int32),
        ("L", C.c_int32),
    ]


class ConvArgs(C.Structure):
    """struct fluxb200_conv_args"""

    _fields_ = [
        ("x", C.c_void_p),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("ldx", C.c_int64),
        ("ldw", C.c_int64),
        ("ld_res", C.c_int64),
        ("ldo", C.c_int64),
        ("B", C.c_int32),
        ("H", C.c_int32),
        ("W", C.c_int32),
        ("Cin", C.c_int32),
        ("N", C.c_int32),
        ("taps", C.c_int32),
        ("out_mode", C.c_int32),
        ("alpha", C.c_float),
        ("stride", C.c_int32),
        ("gn_stats", C.c_void_p),
    ]


class GemmArgs(C.Structure):
    """struct fluxb200_gemm_args"""

    _fields_ = [
        ("a", C.c_void_p),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("a_scale_recip", C.c_void_p),
        ("w_scale_recip", C.c_void_p),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("a_fmt", C.c_int32),
        ("w_fmt", C.c_int32),
        ("epilogue", C.c_int32),
        ("rows_per_batch", C.c_int32),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("resid", C.c_void_p),
        ("ldr", C.c_int64),
        ("gate", C.c_void_p),
        ("gate_batch_stride", C.c_int64),
        ("out_scale", C.c_void_p),
        ("out_fmt", C.c_int32),
        ("out_col_offset", C.c_int32),
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("num_heads", C.c_int32),
        ("seq_total", C.c_int32),
        ("seq_offset", C.c_int32),
        ("_pad0", C.c_int32),
        ("q_norm_w", C.c_void_p),
        ("k_norm_w", C.c_void_p),
        ("rope_cos", C.c_void_p),
        ("rope_sin", C.c_void_p),
        ("rope_batch_stride", C.c_int64),
    ]


class AttentionArgs(C.Structure):
    """struct fluxb200_attention_args"""

    _fields_ = [
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("out_batch_stride", C.c_int64),
        ("B", C.c_int32),
        ("H", C.c_int32),
        ("S", C.c_int32),
        ("softmax_scale", C.c_float),
        ("out_kind", C.c_int32),
        ("out_fmt", C.c_int32),
        ("split_row", C.c_int32),
        ("variant", C.c_int32),
        ("out_scale0", C.c_void_p),
        ("out_scale1", C.c_void_p),
        ("out1", C.c_void_p),
        ("ldo1", C.c_int64),
        ("out1_batch_stride", C.c_int64),
    ]


class GemvLayer(C.Structure):
    """struct fluxb200_gemv_layer"""

    _fields_ = [
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("in_qscale", C.c_void_p),
        ("a_scale_recip", C.c_void_p),
        ("w_scale_recip", C.c_void_p),
        ("N", C.c_int32),
        ("out_offset", C.c_int32),
        ("block_start", C.c_int32),
        ("_pad", C.c_int32),
    ]


class FluxB200Error(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FluxB200Error(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    lib.fluxb200_version.restype = C.c_int
    lib.fluxb200_last_error.restype = C.c_char_p
    lib.fluxb200_device_check.argtypes = [C.POINTER(C.c_int)]
    lib.fluxb200_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.fluxb200_amax.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fluxb200_f8_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.fluxb200_f8_gemm_grouped.argtypes = [C.POINTER(GemmArgs), C.c_int, C.c_void_p]
    lib.fluxb200_f8_gemm_ln.argtypes = [C.POINTER(GemmArgs), C.c_int, C.POINTER(LnArgs), C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_void_p, C.c_void_p]
    lib.fluxb200_f8_gemv.argtypes = [
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_int, C.c_int, C.c_int, C.c_void_p,
    ]
    lib.fluxb200_modulation_batched.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
        C.c_int, C.c_void_p,
    ]
    lib.fluxb200_modulation_batched_bf16.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
    ]
    lib.fluxb200_bf16_gemv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fluxb200_timestep_embedding.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.fluxb200_euler_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.fluxb200_bf16_gemm_small.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fluxb200_silu_quant.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.fluxb200_ln_mod_quant.argtypes = [
        C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
    ]
    lib.fluxb200_ln_mod_quant_grouped.argtypes = [C.POINTER(LnArgs), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.fluxb200_qknorm_rope.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
        C.c_float, C.c_void_p,
    ]
    lib.fluxb200_attention.argtypes = [C.POINTER(AttentionArgs), C.c_void_p]
    lib.fluxb200_lora_fuse.argtypes = [
        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
    ]
    lib.fluxb200_conv2d_nhwc.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
    lib.fluxb200_group_norm_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int64, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.fluxb200_upsample2x_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fluxb200_softmax_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
    lib.fluxb200_vae_latent_prep.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float,
                                             C.c_float, C.c_void_p]
    lib.fluxb200_rows_norm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_float, C.c_void_p]
    lib.fluxb200_gated_act.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fluxb200_attention_d64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.fluxb200_debug_counters.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.fluxb200_gemm_probe_mode.argtypes = [C.c_int]
    lib.fluxb200_gemm_force_tiling.argtypes = [C.c_int, C.c_int]
    lib.fluxb200_fp8_mma_probe.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_void_p]
    for name in EXPORTS:
        if name != "fluxb200_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


#: number of kernels launched through the C ABI by this process (every successful entry-point call except
#: version / last_error / device_check is exactly one kernel launch)
LAUNCHES = 0


def check(rc: int, what: str) -> None:
    """Translate a negative return code into the exception class the reference would raise
    (ValueError for bad shapes, RuntimeError otherwise; SURVEY.md section 8b 'Error conventions')."""
    if rc == 0:
        global LAUNCHES
        LAUNCHES += 1
        return
    msg = load().fluxb200_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise FluxB200Error(f"{what} failed ({rc}): {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def fp8_fmt(dtype: torch.dtype) -> int:
    if dtype == torch.float8_e4m3fn:
        return E4M3
    if dtype == torch.float8_e5m2:
        return E5M2
    raise ValueError(f"unsupported float8 dtype {dtype}")


def fp8_dtype(fmt: int) -> torch.dtype:
    return torch.float8_e4m3fn if fmt == E4M3 else torch.float8_e5m2


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FluxB200Error(
                "flux-fp8-api_b200 runs on CUDA (sm_100a) tensors only; got a CPU tensor. "
                "There is no CPU path -- the CPU oracle lives under oracle/ and is test infrastructure."
            )
