"""ctypes binding of libvex.so (include/vex.h).  There is no CPU or PyTorch fallback: if the
library cannot be built/loaded, importing callers fail loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

VEX_OK = 0
EPI_PLAIN, EPI_ROPE, EPI_SWIGLU, EPI_RESIDUAL, EPI_DROPOUT_ACC, EPI_CE, EPI_CE_BWD = 0, 1, 2, 3, 4, 5, 6
ACT_NONE, ACT_GELU = 0, 1
COUNT_VISION, COUNT_LANGUAGE, COUNT_VALID, COUNT_MAXLEN, NUM_COUNTS = 0, 1, 2, 3, 4

# every symbol include/vex.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "vex_abi_version", "vex_error_string", "vex_last_cuda_error", "vex_device_check", "vex_partition",
    "vex_rmsnorm_gather", "vex_silu_mul", "vex_residual_scatter", "vex_copy_padded_rows", "vex_grouped_gemm",
    "vex_attention", "vex_attention_decode", "vex_gather_rows", "vex_silu_mul_backward", "vex_rmsnorm_backward",
    "vex_lora_wgrad", "vex_attention_lse", "vex_attention_backward", "vex_dropout_rows", "vex_label_rows", "vex_ce_reduce",
    "vex_attention_blockdiag", "vex_layernorm", "vex_patchify", "vex_maxpool_tokens", "vex_scatter_rows",
    "vex_attention_decode_cache", "vex_advance_counter", "vex_kv_clear_padded", "vex_decode_gemm",
]


class GemmArgs(C.Structure):
    """Mirror of ``struct vexGemmArgs`` (include/vex.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64),
        ("w", (C.c_void_p * 2) * 2), ("ldw", C.c_int64),
        ("lora_t", C.c_void_p * 2), ("ldt", C.c_int64),
        ("lora_b", (C.c_void_p * 2) * 2), ("lora_r", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("counts", C.c_void_p), ("row_map", C.c_void_p), ("residual", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("position_ids", C.c_void_p),
        ("sorted_to_flat", C.c_void_p), ("rope_len", C.c_int32), ("rope_cols", C.c_int32),
        ("rows_cap", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("mode", C.c_int32),
        ("single_expert", C.c_int32), ("alpha", C.c_float), ("w_transposed", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64),
        ("ce_labels", C.c_void_p), ("ce_pmax", C.c_void_p), ("ce_psum", C.c_void_p), ("ce_zlabel", C.c_void_p),
        ("ce_lse", C.c_void_p), ("ce_w", C.c_void_p), ("ce_dloss", C.c_void_p),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("kv_k", C.c_void_p), ("kv_v", C.c_void_p), ("kv_pos", C.c_void_p),
        ("kv_seq_len", C.c_int32), ("kv_capacity", C.c_int32),
    ]


class VexError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("VEX_LIB_PATH")  # experiment builds (tools/*_trace.py) load their own library as is
        if path:
            if not os.path.isfile(path):
                raise VexError(f"VEX_LIB_PATH={path} does not exist")
        else:
            path = _build.LIB
        if path == _build.LIB and not _build.is_fresh():
            if os.path.isfile(_build.NVCC):
                path = _build.build()
            elif not os.path.isfile(path):
                raise VexError("libvex.so is missing and nvcc is unavailable: the CUDA extension is required "
                               "(there is no CPU fallback)")
        L = C.CDLL(path)
        p, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
        L.vex_abi_version.restype = C.c_int
        L.vex_error_string.restype = C.c_char_p
        L.vex_error_string.argtypes = [C.c_int]
        L.vex_last_cuda_error.restype = C.c_int
        L.vex_device_check.restype = C.c_int
        L.vex_partition.argtypes = [p, p, i32, i32, p, p, p, p, p, p, p, p, p]
        L.vex_rmsnorm_gather.argtypes = [p, p, i32, f32, p, p, p, p, i32, i32, p]
        L.vex_silu_mul.argtypes = [p, p, p, p, i32, i32, p]
        L.vex_residual_scatter.argtypes = [p, p, p, p, p, i32, i32, p]
        L.vex_copy_padded_rows.argtypes = [p, p, p, i32, i32, p]
        L.vex_grouped_gemm.argtypes = [C.POINTER(GemmArgs), p]
        L.vex_decode_gemm.argtypes = [C.POINTER(GemmArgs), p]
        L.vex_attention.argtypes = [p, p, i32, i32, i32, p, p, f32, p]
        L.vex_attention_decode.argtypes = [p, i64, p, p, p, p, i32, i32, i32, f32, p]
        L.vex_gather_rows.argtypes = [p, p, p, p, i32, i32, p]
        L.vex_silu_mul_backward.argtypes = [p, p, p, p, p, p, i32, i32, p]
        L.vex_rmsnorm_backward.argtypes = [p, p, p, p, i32, f32, p, p, p, p, p, p, i32, i32, p]
        L.vex_attention_lse.argtypes = [p, p, i32, i32, i32, p, p, f32, p, p]
        L.vex_attention_backward.argtypes = [p, p, p, p, p, p, p, p, p, p, p, i32, i32, i32, i32, p, f32, p]
        L.vex_label_rows.argtypes = [p, p, i32, i32, i64, p, p, p, p, p]
        L.vex_ce_reduce.argtypes = [p, p, p, p, p, i32, i32, p, p, p]
        L.vex_dropout_rows.argtypes = [p, p, p, i32, i32, f32, C.c_uint64, p]
        L.vex_lora_wgrad.argtypes = [p, i64, p, i64, i32, p, p, i64, i32, p, i32, i32, p]
        L.vex_attention_blockdiag.argtypes = [p, p, i32, i32, i32, p, p, f32, p]
        L.vex_layernorm.argtypes = [p, p, p, f32, p, i32, p, p, i32, i32, p]
        L.vex_patchify.argtypes = [p, i32, i32, i32, i32, i32, i32, i32, p, i64, p]
        L.vex_maxpool_tokens.argtypes = [p, i64, i32, i32, i32, i32, i32, i32, p, i64, i32, p]
        L.vex_scatter_rows.argtypes = [p, p, p, i32, p, i32, p]
        L.vex_attention_decode_cache.argtypes = [p, i64, p, p, p, i64, i32, p, i32, i32, i32, p, f32, p]
        L.vex_advance_counter.argtypes = [p, i32, p]
        L.vex_kv_clear_padded.argtypes = [p, p, p, i32, i32, i32, i32, p]
        for name in SYMBOLS:
            fn = getattr(L, name)
            if name not in ("vex_error_string",):
                fn.restype = C.c_int
        _lib = L
        return L


def check(rc: int, what: str) -> None:
    if rc != VEX_OK:
        L = lib()
        msg = L.vex_error_string(rc).decode()
        extra = f" (cudaError {L.vex_last_cuda_error()})" if rc == -3 else ""
        raise VexError(f"{what}: {msg}{extra}")
