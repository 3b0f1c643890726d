"""ctypes binding of libvex_baselines.so -- the attention kernels k4_attention_tc3.cu superseded (mma.sync baseline,
one-tile tc1, non-persistent tc2).  Test / tooling infrastructure: the product package never loads this library."""
from __future__ import annotations

import ctypes as C
import os

import torch

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        from mmmm_b200 import build as b
        path = b.LIB_BASELINES
        if not b.baselines_fresh():
            if os.path.isfile(b.NVCC):
                path = b.build_baselines()
            elif not os.path.isfile(path):
                raise RuntimeError("libvex_baselines.so is missing and nvcc is unavailable")
        L = C.CDLL(path)
        p, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.vex_attention_baseline.argtypes = [C.c_char_p, p, p, i32, i32, i32, p, p, f32, p, i32, p]
        L.vex_attention_baseline.restype = C.c_int
        _lib = L
    return _lib


def attention(impl: str, qkv: torch.Tensor, cu_seqlens: torch.Tensor, batch: int, max_len_cap: int, heads: int,
              out_row_map, out: torch.Tensor, scale: float, lse=None, causal: bool = True) -> None:
    """Same contract as ``ops.attention`` / ``ops.attention_train`` (include/vex.h vex_attention_lse)."""
    rc = lib().vex_attention_baseline(impl.encode(), qkv.data_ptr(), cu_seqlens.data_ptr(), batch, max_len_cap, heads,
                                      None if out_row_map is None else out_row_map.data_ptr(), out.data_ptr(),
                                      float(scale), None if lse is None else lse.data_ptr(), int(causal),
                                      torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"vex_attention_baseline({impl}) failed with {rc}")
