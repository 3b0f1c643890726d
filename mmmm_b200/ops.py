"""Tensor-level wrappers of the C ABI (include/vex.h), registered as ``torch.library`` custom ops
(namespace ``vex``).  Every op launches hand-written sm_100a kernels from libvex.so on the current
CUDA stream; outputs are pre-allocated by the caller and mutated in place.  CPU tensors are an
error -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib, instrument
from ._lib import ACT_GELU, ACT_NONE, EPI_CE, EPI_CE_BWD, EPI_DROPOUT_ACC, EPI_PLAIN, EPI_RESIDUAL, EPI_ROPE, EPI_SWIGLU, GemmArgs  # noqa: F401

_BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libvex has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------ K1
@torch.library.custom_op("vex::partition", mutates_args=(
        "sorted_to_flat", "flat_to_sorted", "sorted_to_token", "token_to_sorted", "token_to_flat", "cu_seqlens",
        "counts", "scratch"))
def partition(token_type_ids: torch.Tensor, padding_mask: torch.Tensor, sorted_to_flat: torch.Tensor,
              flat_to_sorted: torch.Tensor, sorted_to_token: torch.Tensor, token_to_sorted: torch.Tensor,
              token_to_flat: torch.Tensor, cu_seqlens: torch.Tensor, counts: torch.Tensor,
              scratch: torch.Tensor) -> None:
    """K1 (vex_partition): replaces get_expert_mask + boolean-mask nonzero (modeling_cogvlm.py:58-70)."""
    _dev(token_type_ids, "token_type_ids", torch.int64)
    _dev(padding_mask, "padding_mask", torch.bool)
    if token_type_ids.dim() != 2 or token_type_ids.shape != padding_mask.shape:
        raise ValueError("token_type_ids and padding_mask must both be [B, L]")
    B, L = token_type_ids.shape
    for name, t, n in (("sorted_to_flat", sorted_to_flat, B * L), ("flat_to_sorted", flat_to_sorted, B * L),
                       ("sorted_to_token", sorted_to_token, B * L), ("token_to_sorted", token_to_sorted, B * L),
                       ("token_to_flat", token_to_flat, B * L), ("cu_seqlens", cu_seqlens, B + 1),
                       ("counts", counts, _lib.NUM_COUNTS), ("scratch", scratch, 4 * B)):
        _dev(t, name, torch.int32)
        if t.numel() < n:
            raise ValueError(f"{name} needs at least {n} int32 elements")
    with instrument.region("partition", 2):
      rc = _lib.lib().vex_partition(token_type_ids.data_ptr(), padding_mask.data_ptr(), B, L,
                                  sorted_to_flat.data_ptr(), flat_to_sorted.data_ptr(), sorted_to_token.data_ptr(),
                                  token_to_sorted.data_ptr(), token_to_flat.data_ptr(), cu_seqlens.data_ptr(),
                                  counts.data_ptr(), scratch.data_ptr(), _stream())
    _lib.check(rc, "vex_partition")


# ------------------------------------------------------------------------------------------ K2
@torch.library.custom_op("vex::rmsnorm_gather", mutates_args=("out",))
def rmsnorm_gather(x: torch.Tensor, weight: torch.Tensor, eps: float, row_src: Optional[torch.Tensor],
                   n_rows: torch.Tensor, out: torch.Tensor, row_dst: Optional[torch.Tensor] = None) -> None:
    """K2 (vex_rmsnorm_gather): RMSNorm.forward (modeling_cogvlm.py:36-41) fused with the row gather."""
    _dev(x, "x", _BF16)
    _dev(out, "out", _BF16)
    _dev(weight, "weight")
    if weight.dtype not in (_BF16, torch.float32):
        raise TypeError("RMSNorm weight must be bf16 or fp32")
    H = x.shape[-1]
    if weight.numel() != H or out.shape[-1] != H:
        raise ValueError("hidden size mismatch")
    if row_src is not None:
        _dev(row_src, "row_src", torch.int32)
    if row_dst is not None:
        _dev(row_dst, "row_dst", torch.int32)
    _dev(n_rows, "n_rows", torch.int32)
    rows_cap = out.numel() // H
    with instrument.region("rmsnorm"):
      rc = _lib.lib().vex_rmsnorm_gather(x.data_ptr(), weight.data_ptr(), int(weight.dtype == torch.float32),
                                       float(eps), _ptr(row_src), _ptr(row_dst), n_rows.data_ptr(), out.data_ptr(),
                                       rows_cap, H, _stream())
    _lib.check(rc, "vex_rmsnorm_gather")


# ------------------------------------------------------------------------------------------ K5 / K6
@torch.library.custom_op("vex::silu_mul", mutates_args=("out",))
def silu_mul(gate: torch.Tensor, up: torch.Tensor, n_rows: torch.Tensor, out: torch.Tensor) -> None:
    """K5 (vex_silu_mul): act_fn(gate) * up (modeling_cogvlm.py:55)."""
    for n, t in (("gate", gate), ("up", up), ("out", out)):
        _dev(t, n, _BF16)
    if gate.shape != up.shape or gate.shape != out.shape:
        raise ValueError("gate/up/out shapes differ")
    I = gate.shape[-1]
    with instrument.region("silu_mul"):
      rc = _lib.lib().vex_silu_mul(gate.data_ptr(), up.data_ptr(), out.data_ptr(), _dev(n_rows, "n_rows",
                                 torch.int32).data_ptr(), gate.numel() // I, I, _stream())
    _lib.check(rc, "vex_silu_mul")


@torch.library.custom_op("vex::residual_scatter", mutates_args=("out",))
def residual_scatter(y: torch.Tensor, residual: torch.Tensor, row_dst: Optional[torch.Tensor],
                     n_rows: torch.Tensor, out: torch.Tensor) -> None:
    """K6 (vex_residual_scatter): out[dst[r]] = residual[dst[r]] + y[r] (modeling_cogvlm.py:278-279 + :321)."""
    for n, t in (("y", y), ("residual", residual), ("out", out)):
        _dev(t, n, _BF16)
    H = y.shape[-1]
    if residual.shape[-1] != H or out.shape != residual.shape:
        raise ValueError("shape mismatch")
    with instrument.region("residual_scatter"):
      rc = _lib.lib().vex_residual_scatter(y.data_ptr(), residual.data_ptr(), _ptr(row_dst),
                                         _dev(n_rows, "n_rows", torch.int32).data_ptr(), out.data_ptr(),
                                         y.numel() // H, H, _stream())
    _lib.check(rc, "vex_residual_scatter")


@torch.library.custom_op("vex::copy_padded_rows", mutates_args=("out",))
def copy_padded_rows(x: torch.Tensor, flat_to_sorted: torch.Tensor, out: torch.Tensor) -> None:
    """Rows with padding_mask == False: out = x (the reference leaves them uninitialised, :277)."""
    _dev(x, "x", _BF16)
    _dev(out, "out", _BF16)
    _dev(flat_to_sorted, "flat_to_sorted", torch.int32)
    H = x.shape[-1]
    with instrument.region("copy_padded_rows"):
      rc = _lib.lib().vex_copy_padded_rows(x.data_ptr(), flat_to_sorted.data_ptr(), out.data_ptr(), x.numel() // H, H,
                                         _stream())
    _lib.check(rc, "vex_copy_padded_rows")


# ------------------------------------------------------------------------------------------ K3
def grouped_gemm_raw(a: torch.Tensor, w: Sequence[Optional[torch.Tensor]], out: torch.Tensor, counts: torch.Tensor,
                     mode: int, *, row_map: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                     lora_t: Sequence[Optional[torch.Tensor]] = (None, None),
                     lora_b: Sequence[Optional[torch.Tensor]] = (None, None, None, None), lora_r: int = 0,
                     rope: Optional[tuple] = None, rope_cols: int = 0, single_expert: bool = False,
                     alpha: float = 1.0, n_out: Optional[int] = None, w_transposed: bool = False,
                     dropout_p: float = 0.0, dropout_seed: int = 0, ce: Optional[dict] = None,
                     bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                     kv: Optional[tuple] = None, skinny: bool = False) -> None:
    """K3 (vex_grouped_gemm).  ``w`` = [vision_w0, vision_w1, language_w0, language_w1] ([N, K] each; the *_w1
    entries are up_proj for SWIGLU, else None).  ``lora_b`` likewise; ``lora_t`` = [T_half0, T_half1].
    ``rope`` = (cos [S,128], sin [S,128], position_ids int64 [B*L], sorted_to_flat int32).
    ``w_transposed``: the weights are [K, N] (out = a . w, the dgrad form dX = dY . W over the nn.Linear weight as
    stored) and ``lora_b`` holds lora_A [r, N].
    ``bias`` (bf16 [N]) / ``act`` (ACT_GELU): nn.Linear with bias and the erf-GELU of the vision encoder's Linears
    (visual.py:84-85, :110-117), PLAIN / RESIDUAL epilogues only.
    ``kv`` = (k_cache, v_cache [B, heads, capacity, 128], seq_len, kv_pos or None): EPI_ROPE also writes the post-rotary
    K heads and the V heads into the KV cache (vexGemmArgs.kv_k).
    ``skinny``: the rows are a decode batch (<= 32, all live, single expert) -- run K12 (vex_decode_gemm), the
    HBM-bound weight-streaming form, instead of the tcgen05 tile kernel."""
    _dev(a, "a", _BF16)
    if out is not None:
        _dev(out, "out", _BF16)
    _dev(counts, "counts", torch.int32)
    K = a.shape[-1]
    rows_cap = a.numel() // K
    w = list(w) + [None] * (4 - len(w))
    args = GemmArgs()
    args.a, args.lda = a.data_ptr(), K
    N = None
    for i, wt in enumerate(w):
        if wt is None:
            continue
        _dev(wt, f"w[{i}]", _BF16)
        kd, nd = (0, 1) if w_transposed else (1, 0)
        if wt.dim() != 2 or wt.shape[kd] != K:
            raise ValueError(f"w[{i}] must be {'[K, N]' if w_transposed else '[N, K]'} with K = {K}, got {tuple(wt.shape)}")
        N = wt.shape[nd] if N is None else N
        if wt.shape[nd] != N:
            raise ValueError("all weights of one grouped GEMM must share N")
        args.w[i // 2][i % 2] = wt.data_ptr()
    if N is None:
        raise ValueError("no weights given")
    args.ldw = N if w_transposed else K
    args.w_transposed = int(w_transposed)
    if lora_r:
        for h, t in enumerate(lora_t):
            if t is not None:
                _dev(t, f"lora_t[{h}]", _BF16)
                args.lora_t[h] = t.data_ptr()
                args.ldt = t.shape[-1]
        lora_b = list(lora_b) + [None] * (4 - len(lora_b))
        for i, bt in enumerate(lora_b):
            if bt is None:
                continue
            _dev(bt, f"lora_b[{i}]", _BF16)
            if tuple(bt.shape) != ((lora_r, N) if w_transposed else (N, lora_r)):
                raise ValueError(f"lora_b[{i}] must be [{N}, {lora_r}] ([{lora_r}, {N}] = lora_A when transposed)")
            args.lora_b[i // 2][i % 2] = bt.data_ptr()
        args.lora_r = lora_r
    if out is not None:
        args.out, args.ldo = out.data_ptr(), out.shape[-1]
    args.counts = counts.data_ptr()
    if ce is not None:  # fused lm_head + cross-entropy epilogues (EPI_CE / EPI_CE_BWD)
        for key, dt in (("labels", torch.int32), ("pmax", torch.float32), ("psum", torch.float32),
                        ("zlabel", torch.float32), ("lse", torch.float32), ("w", torch.float32),
                        ("dloss", torch.float32)):
            if ce.get(key) is not None:
                setattr(args, "ce_" + key, _dev(ce[key], "ce." + key, dt).data_ptr())
    args.row_map = _ptr(None if row_map is None else _dev(row_map, "row_map", torch.int32))
    if residual is not None:
        _dev(residual, "residual", _BF16)
        if residual.shape[-1] != out.shape[-1]:
            raise ValueError("residual/out layout mismatch")
        args.residual = residual.data_ptr()
    elif mode in (EPI_RESIDUAL, EPI_DROPOUT_ACC):
        args.residual = out.data_ptr()  # in place: out[row] += acc (each 16-byte piece is read then written once)
    if rope is not None:
        cos, sin, pos, s2f = rope
        _dev(cos, "cos", _BF16), _dev(sin, "sin", _BF16)
        _dev(pos, "position_ids", torch.int64), _dev(s2f, "sorted_to_flat", torch.int32)
        if cos.shape[-1] != 128 or cos.shape != sin.shape:
            raise ValueError("rotary tables must be [S, 128]")
        args.rope_cos, args.rope_sin = cos.data_ptr(), sin.data_ptr()
        args.position_ids, args.sorted_to_flat = pos.data_ptr(), s2f.data_ptr()
        args.rope_len, args.rope_cols = cos.shape[0], rope_cols
    args.rows_cap, args.N, args.K, args.mode = rows_cap, (n_out or N), K, mode
    args.single_expert = int(single_expert)
    args.alpha = alpha
    args.dropout_p, args.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    if bias is not None:
        _dev(bias, "bias", _BF16)
        if bias.numel() != (n_out or N):
            raise ValueError("bias must have one entry per output feature")
        args.bias = bias.data_ptr()
    args.act = int(act)
    if kv is not None:
        k_cache, v_cache, kv_seq, kv_pos = kv
        _dev(k_cache, "k_cache", _BF16), _dev(v_cache, "v_cache", _BF16)
        if k_cache.dim() != 4 or k_cache.shape != v_cache.shape or k_cache.shape[-1] != 128:
            raise ValueError("KV cache tensors must be [B, heads, capacity, 128]")
        if mode != EPI_ROPE or k_cache.shape[1] * 128 * 3 != N:
            raise ValueError("the KV-cache output belongs to the QKV projection's EPI_ROPE epilogue")
        args.kv_k, args.kv_v = k_cache.data_ptr(), v_cache.data_ptr()
        args.kv_pos = _ptr(None if kv_pos is None else _dev(kv_pos, "kv_pos", torch.int32))
        args.kv_seq_len, args.kv_capacity = int(kv_seq), k_cache.shape[2]
    name = ("gemm_plain", "gemm_rope", "gemm_swiglu", "gemm_residual", "gemm_dropout_acc", "gemm_ce", "gemm_ce_bwd")[mode] + ("_n64" if N <= 64 else "") + \
        ("_dgrad" if w_transposed else "") + ("_gelu" if act == ACT_GELU else "")
    if skinny:
        if rope is not None and pos.numel() != rows_cap:
            raise ValueError("skinny (decode) form: position_ids must have one entry per batch row")
        with instrument.region("decode_" + name):
          rc = _lib.lib().vex_decode_gemm(C.byref(args), _stream())
        _lib.check(rc, "vex_decode_gemm")
        return
    with instrument.region(name):
      rc = _lib.lib().vex_grouped_gemm(C.byref(args), _stream())
    _lib.check(rc, "vex_grouped_gemm")


@torch.library.custom_op("vex::grouped_gemm", mutates_args=("out",))
def grouped_gemm(a: torch.Tensor, w_vision: torch.Tensor, w_language: Optional[torch.Tensor], out: torch.Tensor,
                 counts: torch.Tensor, row_map: Optional[torch.Tensor], alpha: float) -> None:
    """Plain grouped GEMM as a registered op (the fused-epilogue variants are reached through
    ``vex::visual_expert_layer``): out[map(r)] = bf16(alpha * a[r] . w_e^T)."""
    grouped_gemm_raw(a, [w_vision, None, w_language, None], out, counts, EPI_PLAIN, row_map=row_map, alpha=alpha,
                     single_expert=w_language is None)


@torch.library.custom_op("vex::grouped_gemm_fused", mutates_args=("out", "kv_k", "kv_v"))
def _grouped_gemm_fused(a: torch.Tensor, w: List[Optional[torch.Tensor]], out: torch.Tensor, counts: torch.Tensor,
                        mode: int, row_map: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                        lora_t: List[Optional[torch.Tensor]], lora_b: List[Optional[torch.Tensor]], lora_r: int,
                        rope: List[torch.Tensor], rope_cols: int, single_expert: bool, alpha: float,
                        kv_k: Optional[torch.Tensor], kv_v: Optional[torch.Tensor], kv_seq_len: int,
                        kv_pos: Optional[torch.Tensor], skinny: bool) -> None:
    """K3 with a fused epilogue (``mode`` = EPI_*), LoRA K-extension and scatter; see ``grouped_gemm_raw``.
    ``kv_k`` / ``kv_v`` [B, heads, capacity, 128]: KV-cache second output of EPI_ROPE (prefill: ``kv_seq_len`` = L;
    decode: 1 and ``kv_pos`` = device counter of cached positions)."""
    grouped_gemm_raw(a, w, out, counts, mode, row_map=row_map, residual=residual, lora_t=lora_t, lora_b=lora_b,
                     lora_r=lora_r, rope=tuple(rope) if len(rope) else None, rope_cols=rope_cols,
                     single_expert=single_expert, alpha=alpha,
                     kv=None if kv_k is None else (kv_k, kv_v, kv_seq_len, kv_pos), skinny=skinny)


def grouped_gemm_fused(a, w, out, counts, mode, row_map, residual, lora_t, lora_b, lora_r, rope, rope_cols,
                       single_expert, alpha, kv_k=None, kv_v=None, kv_seq_len=0, kv_pos=None, skinny=False) -> None:
    """``vex::grouped_gemm_fused`` with the KV-cache arguments defaulted (the registered op takes every argument
    positionally: torch.library tracks mutated arguments by position)."""
    _grouped_gemm_fused(a, w, out, counts, mode, row_map, residual, lora_t, lora_b, lora_r, rope, rope_cols,
                        single_expert, alpha, kv_k, kv_v, kv_seq_len, kv_pos, skinny)


@torch.library.custom_op("vex::grouped_gemm_dgrad", mutates_args=("out",))
def grouped_gemm_dgrad(dy: torch.Tensor, w: List[Optional[torch.Tensor]], out: torch.Tensor, counts: torch.Tensor,
                       accumulate: bool, row_map: Optional[torch.Tensor], lora_dt: Optional[torch.Tensor],
                       lora_a: List[Optional[torch.Tensor]], lora_r: int, single_expert: bool, alpha: float,
                       dropout_p: float = 0.0, dropout_seed: int = 0) -> None:
    """Backward of the routed Linear w.r.t. its input (what autograd computes for modeling_cogvlm.py:244-245,
    :278-279, :96-97): out[map(r)] (+)= dy[r] . W_e (+ dT[r] . A_e), with ``w`` = [vision W, language W] as STORED
    ([out_features, in_features]; read as an MN-major B operand, no transposed copy), ``lora_dt`` = scaling * dy .
    lora_B and ``lora_a`` = [vision lora_A, language lora_A] ([r, in_features]).  ``accumulate`` adds onto ``out``
    in place (second term of d(xn) = dgate . Wg + dup . Wu).  ``dropout_p`` > 0 (with ``accumulate``): the product is
    masked with the LoRA input-dropout mask of ``dropout_seed`` and scaled by 1 / (1 - p) before it is added --
    out += keep * (dT . A) / (1 - p), the adjoint of lora_A(dropout(x))."""
    wv, wl = (list(w) + [None])[:2]
    av, al = (list(lora_a) + [None, None])[:2]
    mode = EPI_RESIDUAL if accumulate else EPI_PLAIN
    if dropout_p > 0:
        if not accumulate:
            raise ValueError("the dropout-masked dgrad accumulates onto `out`")
        mode = EPI_DROPOUT_ACC
    grouped_gemm_raw(dy, [wv, None, wl, None], out, counts, mode, row_map=row_map,
                     lora_t=[lora_dt, None], lora_b=[av, None, al, None], lora_r=lora_r, single_expert=single_expert,
                     alpha=alpha, w_transposed=True, dropout_p=dropout_p, dropout_seed=dropout_seed)


# ------------------------------------------------------------------------------------------ K4
def _attention_impl(qkv, cu_seqlens, batch, max_len_cap, heads, out_row_map, out, scale, lse):
    _dev(qkv, "qkv", _BF16)
    _dev(out, "out", _BF16)
    _dev(cu_seqlens, "cu_seqlens", torch.int32)
    if qkv.shape[-1] != 3 * heads * 128:
        raise ValueError("qkv rows must be [3 * heads * 128] wide (head_dim 128 only)")
    if out_row_map is not None:
        _dev(out_row_map, "out_row_map", torch.int32)
    if lse is not None:
        _dev(lse, "lse", torch.float32)
        if lse.numel() < heads * batch * max_len_cap:
            raise ValueError("lse must hold heads * B * max_len_cap floats")
    with instrument.region("attention", 2):  # tail-row zeroing + the attention kernel
      rc = _lib.lib().vex_attention_lse(qkv.data_ptr(), cu_seqlens.data_ptr(), batch, max_len_cap, heads,
                                      _ptr(out_row_map), out.data_ptr(), float(scale), _ptr(lse), _stream())
    _lib.check(rc, "vex_attention")


@torch.library.custom_op("vex::attention", mutates_args=("out",))
def attention(qkv: torch.Tensor, cu_seqlens: torch.Tensor, batch: int, max_len_cap: int, heads: int,
              out_row_map: Optional[torch.Tensor], out: torch.Tensor, scale: float) -> None:
    """K4 (vex_attention): causal block-diagonal attention over token-order QKV [T, 3, heads, 128]
    (attention_fn prefill branch, modeling_cogvlm.py:106-128)."""
    _attention_impl(qkv, cu_seqlens, batch, max_len_cap, heads, out_row_map, out, scale, None)


@torch.library.custom_op("vex::attention_train", mutates_args=("out", "lse"))
def attention_train(qkv: torch.Tensor, cu_seqlens: torch.Tensor, batch: int, max_len_cap: int, heads: int,
                    out_row_map: Optional[torch.Tensor], out: torch.Tensor, scale: float, lse: torch.Tensor) -> None:
    """K4 for the training step (vex_attention_lse): also writes the log2-domain log-sum-exp, fp32 [heads, B*L]."""
    _attention_impl(qkv, cu_seqlens, batch, max_len_cap, heads, out_row_map, out, scale, lse)


@torch.library.custom_op("vex::attention_backward", mutates_args=("d_out_tok", "delta_ws", "dqkv"))
def attention_backward(qkv: torch.Tensor, out_sorted: torch.Tensor, d_out_tok: torch.Tensor, lse: torch.Tensor,
                       delta_ws: torch.Tensor, cu_seqlens: torch.Tensor, token_to_sorted: Optional[torch.Tensor],
                       token_to_flat: Optional[torch.Tensor], position_ids: torch.Tensor, cos: torch.Tensor,
                       sin: torch.Tensor, batch: int, max_len_cap: int, heads: int, dqkv: torch.Tensor,
                       scale: float) -> None:
    """K9 (vex_attention_backward): adjoint of the causal varlen attention and of the rotary embedding
    (modeling_cogvlm.py:106-128, :188-193 under autograd); ``dqkv`` receives d(pre-rotary q | k | v) in sorted order."""
    for n, t in (("qkv", qkv), ("out_sorted", out_sorted), ("d_out_tok", d_out_tok), ("dqkv", dqkv), ("cos", cos),
                 ("sin", sin)):
        _dev(t, n, _BF16)
    _dev(lse, "lse", torch.float32), _dev(delta_ws, "delta_ws", torch.float32)
    _dev(cu_seqlens, "cu_seqlens", torch.int32), _dev(position_ids, "position_ids", torch.int64)
    H = heads * 128
    cap = batch * max_len_cap
    if qkv.shape[-1] != 3 * H or dqkv.shape[-1] != 3 * H or out_sorted.shape[-1] != H or d_out_tok.shape[-1] != H:
        raise ValueError("row widths must be 3*heads*128 (qkv, dqkv) and heads*128 (out, d_out)")
    for n, t, w in (("qkv", qkv, 3 * H), ("dqkv", dqkv, 3 * H), ("out_sorted", out_sorted, H), ("d_out_tok", d_out_tok, H)):
        if t.numel() < cap * w:
            raise ValueError(f"{n} must have B * max_len_cap rows")
    if lse.numel() < heads * cap or delta_ws.numel() < heads * cap:
        raise ValueError("lse / delta_ws must hold heads * B * max_len_cap floats")
    if cos.shape[-1] != 128 or cos.shape != sin.shape:
        raise ValueError("rotary tables must be [S, 128]")
    maps = [None if m is None else _dev(m, "row map", torch.int32) for m in (token_to_sorted, token_to_flat)]
    with instrument.region("attention_backward", 4):
      rc = _lib.lib().vex_attention_backward(qkv.data_ptr(), out_sorted.data_ptr(), d_out_tok.data_ptr(), lse.data_ptr(),
                                             delta_ws.data_ptr(), cu_seqlens.data_ptr(), _ptr(maps[0]), _ptr(maps[1]),
                                             position_ids.data_ptr(), cos.data_ptr(), sin.data_ptr(), cos.shape[0],
                                             batch, max_len_cap, heads, dqkv.data_ptr(), float(scale), _stream())
    _lib.check(rc, "vex_attention_backward")


@torch.library.custom_op("vex::attention_decode", mutates_args=("out",))
def attention_decode(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor, out: torch.Tensor,
                     scale: float) -> None:
    """K4d (vex_attention_decode): generation branch of attention_fn (modeling_cogvlm.py:129-141).
    q [B, heads*128] (a row-strided view is fine), k / v [B, heads, L, 128], mask bool [B, L], out [B, heads*128]."""
    _dev(k, "k", _BF16), _dev(v, "v", _BF16), _dev(out, "out", _BF16), _dev(mask, "mask", torch.bool)
    if not q.is_cuda or q.dtype != _BF16 or q.stride(-1) != 1:
        raise ValueError("q must be a CUDA bf16 tensor with unit inner stride")
    B, heads, L, d = k.shape
    if d != 128 or v.shape != k.shape or q.shape != (B, heads * 128) or mask.shape != (B, L):
        raise ValueError("shape mismatch (head_dim 128 only)")
    with instrument.region("attention_decode"):
      rc = _lib.lib().vex_attention_decode(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), mask.data_ptr(),
                                           out.data_ptr(), B, heads, L, float(scale), _stream())
    _lib.check(rc, "vex_attention_decode")


@torch.library.custom_op("vex::attention_decode_cache", mutates_args=("out",))
def attention_decode_cache(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, mask: torch.Tensor,
                           kv_len: torch.Tensor, out: torch.Tensor, scale: float) -> None:
    """K4d over the pre-allocated cache (vex_attention_decode_cache): k_cache / v_cache [B, heads, capacity, 128],
    mask bool [B, M] covering the attended positions (row-strided view is fine), kv_len int32 [1] on the device =
    positions cached before this step; attends to positions [0, kv_len] (clamped to min(capacity, M))."""
    _dev(k_cache, "k_cache", _BF16), _dev(v_cache, "v_cache", _BF16), _dev(out, "out", _BF16)
    _dev(kv_len, "kv_len", torch.int32)
    if not q.is_cuda or q.dtype != _BF16 or q.stride(-1) != 1:
        raise ValueError("q must be a CUDA bf16 tensor with unit inner stride")
    B, heads, cap, d = k_cache.shape
    if d != 128 or v_cache.shape != k_cache.shape or q.shape != (B, heads * 128):
        raise ValueError("shape mismatch (head_dim 128 only)")
    if not mask.is_cuda or mask.dtype != torch.bool or mask.dim() != 2 or mask.shape[0] != B or mask.stride(1) != 1:
        raise ValueError("mask must be a CUDA bool [B, M] tensor with unit inner stride")
    with instrument.region("attention_decode"):
      rc = _lib.lib().vex_attention_decode_cache(q.data_ptr(), q.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                                                 mask.data_ptr(), mask.stride(0) if B > 1 else mask.shape[1],
                                                 mask.shape[1], out.data_ptr(), B, heads, cap,
                                                 kv_len.data_ptr(), float(scale), _stream())
    _lib.check(rc, "vex_attention_decode_cache")


@torch.library.custom_op("vex::advance_counter", mutates_args=("counter",))
def advance_counter(counter: torch.Tensor, by: int) -> None:
    """counter[0] += by on the stream (vex_advance_counter): the decode graph's own past-length increment."""
    _dev(counter, "counter", torch.int32)
    with instrument.region("advance_counter"):
      rc = _lib.lib().vex_advance_counter(counter.data_ptr(), int(by), _stream())
    _lib.check(rc, "vex_advance_counter")


@torch.library.custom_op("vex::kv_clear_padded", mutates_args=("k_cache", "v_cache"))
def kv_clear_padded(k_cache: torch.Tensor, v_cache: torch.Tensor, flat_to_sorted: torch.Tensor, batch: int,
                    seq_len: int) -> None:
    """Zeroes the cache rows of padded prefill positions (vex_kv_clear_padded; the reference cache holds zeros there,
    modeling_cogvlm.py:243)."""
    _dev(k_cache, "k_cache", _BF16), _dev(v_cache, "v_cache", _BF16), _dev(flat_to_sorted, "flat_to_sorted", torch.int32)
    B, heads, cap, d = k_cache.shape
    if d != 128 or v_cache.shape != k_cache.shape or B != batch or cap < seq_len or flat_to_sorted.numel() < B * seq_len:
        raise ValueError("KV cache must be [B, heads, capacity >= L, 128]")
    with instrument.region("kv_clear_padded"):
      rc = _lib.lib().vex_kv_clear_padded(k_cache.data_ptr(), v_cache.data_ptr(), flat_to_sorted.data_ptr(), B, seq_len,
                                          heads, cap, _stream())
    _lib.check(rc, "vex_kv_clear_padded")


# ------------------------------------------------------------------------------------------ K7 (backward, row-wise)
@torch.library.custom_op("vex::gather_rows", mutates_args=("out",))
def gather_rows(x: torch.Tensor, row_src: Optional[torch.Tensor], n_rows: torch.Tensor, out: torch.Tensor) -> None:
    """out[r] = x[row_src[r]], r < *n_rows (vex_gather_rows): d_out [B*L, H] -> expert-sorted rows."""
    _dev(x, "x", _BF16), _dev(out, "out", _BF16), _dev(n_rows, "n_rows", torch.int32)
    H = x.shape[-1]
    if out.shape[-1] != H:
        raise ValueError("row width mismatch")
    with instrument.region("gather_rows"):
      rc = _lib.lib().vex_gather_rows(x.data_ptr(), _ptr(None if row_src is None else _dev(row_src, "row_src", torch.int32)),
                                      n_rows.data_ptr(), out.data_ptr(), out.numel() // H, H, _stream())
    _lib.check(rc, "vex_gather_rows")


@torch.library.custom_op("vex::dropout_rows", mutates_args=("out",))
def dropout_rows(x: torch.Tensor, n_rows: torch.Tensor, out: torch.Tensor, p: float, seed: int) -> None:
    """LoRA input dropout (vex_dropout_rows): out = keep ? x / (1 - p) : 0 on the first *n_rows rows, mask from a
    counter-based hash of (row * K + col, seed) -- PEFT lora.Linear: lora_A(dropout(x))."""
    _dev(x, "x", _BF16), _dev(out, "out", _BF16), _dev(n_rows, "n_rows", torch.int32)
    if x.shape != out.shape:
        raise ValueError("x / out shapes differ")
    K = x.shape[-1]
    with instrument.region("dropout_rows"):
      rc = _lib.lib().vex_dropout_rows(x.data_ptr(), out.data_ptr(), n_rows.data_ptr(), x.numel() // K, K, float(p),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, _stream())
    _lib.check(rc, "vex_dropout_rows")


@torch.library.custom_op("vex::silu_mul_backward", mutates_args=("dgate", "dup"))
def silu_mul_backward(dact: torch.Tensor, gate: torch.Tensor, up: torch.Tensor, n_rows: torch.Tensor,
                      dgate: torch.Tensor, dup: torch.Tensor) -> None:
    """Adjoint of act_fn(gate) * up (modeling_cogvlm.py:55) -- vex_silu_mul_backward."""
    for n, t in (("dact", dact), ("gate", gate), ("up", up), ("dgate", dgate), ("dup", dup)):
        _dev(t, n, _BF16)
        if t.shape != dact.shape:
            raise ValueError("dact/gate/up/dgate/dup shapes differ")
    I = dact.shape[-1]
    with instrument.region("silu_mul_backward"):
      rc = _lib.lib().vex_silu_mul_backward(dact.data_ptr(), gate.data_ptr(), up.data_ptr(), dgate.data_ptr(),
                                            dup.data_ptr(), _dev(n_rows, "n_rows", torch.int32).data_ptr(),
                                            dact.numel() // I, I, _stream())
    _lib.check(rc, "vex_silu_mul_backward")


@torch.library.custom_op("vex::rmsnorm_backward", mutates_args=("dx", "dweight"))
def rmsnorm_backward(dy: torch.Tensor, x: torch.Tensor, x_map: Optional[torch.Tensor], weight: torch.Tensor, eps: float,
                     add: Optional[torch.Tensor], add_map: Optional[torch.Tensor], dx: torch.Tensor,
                     dx_map: Optional[torch.Tensor], dweight: Optional[torch.Tensor], n_rows: torch.Tensor) -> None:
    """Adjoint of RMSNorm.forward (modeling_cogvlm.py:36-41) fused with the residual-branch gradient add and the
    gather / scatter through the row maps -- vex_rmsnorm_backward.  ``dweight`` (fp32 [H]) is accumulated into."""
    _dev(dy, "dy", _BF16), _dev(x, "x", _BF16), _dev(dx, "dx", _BF16), _dev(weight, "weight")
    if weight.dtype not in (_BF16, torch.float32):
        raise TypeError("RMSNorm weight must be bf16 or fp32")
    H = dy.shape[-1]
    if x.shape[-1] != H or dx.shape[-1] != H or weight.numel() != H:
        raise ValueError("hidden size mismatch")
    if add is not None:
        _dev(add, "add", _BF16)
    if dweight is not None:
        _dev(dweight, "dweight", torch.float32)
        if dweight.numel() != H:
            raise ValueError("dweight must be fp32 [H]")
    maps = [None if m is None else _dev(m, "row map", torch.int32) for m in (x_map, add_map, dx_map)]
    with instrument.region("rmsnorm_backward"):
      rc = _lib.lib().vex_rmsnorm_backward(dy.data_ptr(), x.data_ptr(), _ptr(maps[0]), weight.data_ptr(),
                                           int(weight.dtype == torch.float32), float(eps), _ptr(add), _ptr(maps[1]),
                                           dx.data_ptr(), _ptr(maps[2]), _ptr(dweight),
                                           _dev(n_rows, "n_rows", torch.int32).data_ptr(), dy.numel() // H, H, _stream())
    _lib.check(rc, "vex_rmsnorm_backward")


# ------------------------------------------------------------------------------------------ K10 (lm_head + CE)
@torch.library.custom_op("vex::label_rows", mutates_args=("row_idx", "label_sel", "w_sel", "count"))
def label_rows(labels: torch.Tensor, weight: Optional[torch.Tensor], ignore_index: int, row_idx: torch.Tensor,
               label_sel: torch.Tensor, w_sel: torch.Tensor, count: torch.Tensor) -> None:
    """vex_label_rows: ordered compaction of the positions with labels != ignore_index (the rows
    _sample_weighted_ce keeps, modeling_cogvlm.py:619-626)."""
    _dev(labels, "labels", torch.int64)
    n = labels.numel()
    for nm, t, dt in (("row_idx", row_idx, torch.int32), ("label_sel", label_sel, torch.int32),
                      ("w_sel", w_sel, torch.float32)):
        _dev(t, nm, dt)
        if t.numel() < n:
            raise ValueError(f"{nm} needs {n} elements")
    _dev(count, "count", torch.int32)
    if weight is not None:
        _dev(weight, "weight")
        if weight.dtype not in (_BF16, torch.float32) or weight.numel() != n:
            raise ValueError("weight must be bf16 / fp32 with one entry per label")
    with instrument.region("label_rows"):
      rc = _lib.lib().vex_label_rows(labels.data_ptr(), _ptr(weight),
                                     int(weight is not None and weight.dtype == torch.float32), n, int(ignore_index),
                                     row_idx.data_ptr(), label_sel.data_ptr(), w_sel.data_ptr(), count.data_ptr(),
                                     _stream())
    _lib.check(rc, "vex_label_rows")


@torch.library.custom_op("vex::lm_head_ce_forward", mutates_args=("pmax", "psum", "zlabel", "lse", "loss"))
def lm_head_ce_forward(h_sel: torch.Tensor, w: torch.Tensor, label_sel: torch.Tensor, w_sel: torch.Tensor,
                       count: torch.Tensor, lora_t: Optional[torch.Tensor], lora_b: Optional[torch.Tensor], lora_r: int,
                       pmax: torch.Tensor, psum: torch.Tensor, zlabel: torch.Tensor, lse: torch.Tensor,
                       loss: torch.Tensor) -> None:
    """Fused lm_head + weighted cross-entropy, forward (CogVLMForCausalLM.forward :701-706, _sample_weighted_ce
    :610-627): the vocabulary GEMM over the selected rows with the softmax statistics in its epilogue (VEX_EPI_CE,
    no logits in HBM) + vex_ce_reduce.  ``loss`` (fp32 [1]) must be zero on entry."""
    V = w.shape[0]
    tiles = 2 * ((V + 255) // 256)  # one partial slot per 128-column half tile (include/vex.h)
    cap = h_sel.shape[0]
    if pmax.numel() < cap * tiles or psum.numel() < cap * tiles or zlabel.numel() < cap or lse.numel() < cap:
        raise ValueError("partials / lse buffers too small")
    grouped_gemm_raw(h_sel, [w, None, None, None], None, count, EPI_CE, lora_t=[lora_t, None],
                     lora_b=[lora_b, None, None, None], lora_r=lora_r, single_expert=True,
                     ce=dict(labels=label_sel, pmax=pmax, psum=psum, zlabel=zlabel))
    with instrument.region("ce_reduce"):
      rc = _lib.lib().vex_ce_reduce(pmax.data_ptr(), psum.data_ptr(), zlabel.data_ptr(),
                                    _dev(w_sel, "w_sel", torch.float32).data_ptr(), count.data_ptr(), cap, tiles,
                                    lse.data_ptr(), _dev(loss, "loss", torch.float32).data_ptr(), _stream())
    _lib.check(rc, "vex_ce_reduce")


@torch.library.custom_op("vex::lm_head_ce_backward", mutates_args=("dz",))
def lm_head_ce_backward(h_sel: torch.Tensor, w: torch.Tensor, label_sel: torch.Tensor, w_sel: torch.Tensor,
                        count: torch.Tensor, lora_t: Optional[torch.Tensor], lora_b: Optional[torch.Tensor],
                        lora_r: int, lse: torch.Tensor, dloss: torch.Tensor, dz: torch.Tensor) -> None:
    """d(loss)/d(logits) of the selected rows as bf16 [rows, V] (VEX_EPI_CE_BWD: logits recomputed by the GEMM,
    softmax from the saved log-sum-exp) -- the A operand of the lm_head dgrad GEMM."""
    if tuple(dz.shape) != (h_sel.shape[0], w.shape[0]):
        raise ValueError("dz must be [rows, V]")
    grouped_gemm_raw(h_sel, [w, None, None, None], dz, count, EPI_CE_BWD, lora_t=[lora_t, None],
                     lora_b=[lora_b, None, None, None], lora_r=lora_r, single_expert=True,
                     ce=dict(labels=label_sel, lse=lse, w=w_sel, dloss=dloss))


# ------------------------------------------------------------------------------------------ vision encoder (8(f)-4)
@torch.library.custom_op("vex::linear_bias_act", mutates_args=("out",))
def linear_bias_act(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor,
                    n_rows: torch.Tensor, row_map: Optional[torch.Tensor], accumulate: bool, gelu: bool) -> None:
    """K3 as a single-expert Linear with bias / GELU / accumulate epilogue, rows r < n_rows[0]:
    out[map(r)] = act(bf16(a[r] . w^T + bias))   or, with ``accumulate``,   out[map(r)] += bf16(a[r] . w^T + bias)
    -- the vision encoder's nn.Linear calls (visual.py:92, :99, :114-116) and the patch convolution as a GEMM added
    onto the position-embedding rows (visual.py:65-72)."""
    grouped_gemm_raw(a, [w, None, None, None], out, n_rows, EPI_RESIDUAL if accumulate else EPI_PLAIN,
                     row_map=row_map, single_expert=True, bias=bias, act=ACT_GELU if gelu else ACT_NONE)


@torch.library.custom_op("vex::attention_blockdiag", mutates_args=("out",))
def attention_blockdiag(qkv: torch.Tensor, cu_seqlens: torch.Tensor, batch: int, max_len_cap: int, heads: int,
                        out: torch.Tensor, scale: float) -> None:
    """K4 non-causal (vex_attention_blockdiag): memory_efficient_attention under a BlockDiagonalMask
    (visual.py:96-98) over packed QKV [B * max_len_cap, 3, heads, 128]."""
    _dev(qkv, "qkv", _BF16), _dev(out, "out", _BF16), _dev(cu_seqlens, "cu_seqlens", torch.int32)
    if qkv.shape[-1] != 3 * heads * 128 or out.shape[-1] != heads * 128:
        raise ValueError("qkv rows must be [3 * heads * 128] and out rows [heads * 128] wide (head slots of 128)")
    cap = batch * max_len_cap
    if qkv.numel() < cap * 3 * heads * 128 or out.numel() < cap * heads * 128 or cu_seqlens.numel() < batch + 1:
        raise ValueError("qkv / out must have B * max_len_cap rows")
    with instrument.region("attention_blockdiag", 2):
      rc = _lib.lib().vex_attention_blockdiag(qkv.data_ptr(), cu_seqlens.data_ptr(), batch, max_len_cap, heads, None,
                                              out.data_ptr(), float(scale), _stream())
    _lib.check(rc, "vex_attention_blockdiag")


@torch.library.custom_op("vex::layernorm", mutates_args=("out",))
def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float, accumulate: bool,
              gelu: bool, n_rows: torch.Tensor, out: torch.Tensor) -> None:
    """K11 (vex_layernorm) over the first n_rows[0] rows: out = [gelu] LayerNorm(x), or with ``accumulate``
    out += LayerNorm(x) -- the post-branch LayerNorm + residual of TransformerLayer.forward (visual.py:128-135) and
    LayerNorm + GELU of the GLU projector (:173-174)."""
    _dev(x, "x", _BF16), _dev(out, "out", _BF16), _dev(weight, "weight", _BF16), _dev(n_rows, "n_rows", torch.int32)
    H = x.shape[-1]
    if weight.numel() != H or out.shape != x.shape:
        raise ValueError("hidden size mismatch")
    if bias is not None and _dev(bias, "bias", _BF16).numel() != H:
        raise ValueError("bias size mismatch")
    with instrument.region("layernorm"):
      rc = _lib.lib().vex_layernorm(x.data_ptr(), weight.data_ptr(), _ptr(bias), float(eps),
                                    out.data_ptr() if accumulate else None, ACT_GELU if gelu else ACT_NONE,
                                    n_rows.data_ptr(), out.data_ptr(), x.numel() // H, H, _stream())
    _lib.check(rc, "vex_layernorm")


@torch.library.custom_op("vex::patchify", mutates_args=("out",))
def patchify(image: torch.Tensor, pd: int, ph: int, pw: int, out: torch.Tensor) -> None:
    """K11 (vex_patchify): im2col of the kernel == stride patch convolution (visual.py:65, resample.py:56-63):
    image [C, D, H, W] -> out [(D/pd)(H/ph)(W/pw), >= C*pd*ph*pw] (a row-strided 2-D view is fine)."""
    _dev(image, "image", _BF16)
    if not out.is_cuda or out.dtype != _BF16 or out.dim() != 2 or out.stride(1) != 1:
        raise ValueError("out must be a 2-D CUDA bf16 tensor with unit inner stride")
    C, D, H, W = image.shape
    n = (D // pd) * (H // ph) * (W // pw)
    if out.shape[0] != n or out.shape[1] < C * pd * ph * pw:
        raise ValueError(f"out must be [{n}, >= {C * pd * ph * pw}]")
    with instrument.region("patchify"):
      rc = _lib.lib().vex_patchify(image.data_ptr(), C, D, H, W, pd, ph, pw, out.data_ptr(), out.stride(0), _stream())
    _lib.check(rc, "vex_patchify")


@torch.library.custom_op("vex::maxpool_tokens", mutates_args=("out",))
def maxpool_tokens(x: torch.Tensor, grid: List[int], pool: List[int], out: torch.Tensor) -> None:
    """K11 (vex_maxpool_tokens): F.max_pool3d over the (d, h, w) patch grid in token-major layout
    (visual.py:199-202); x [d*h*w, C] and out [(d/pz)(h/py)(w/px), C] may be row-strided views."""
    for n, t in (("x", x), ("out", out)):
        if not t.is_cuda or t.dtype != _BF16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError(f"{n} must be a 2-D CUDA bf16 tensor with unit inner stride")
    gd, gh, gw = grid
    pz, py, px = pool
    C = x.shape[1]
    if x.shape[0] != gd * gh * gw or out.shape != ((gd // pz) * (gh // py) * (gw // px), C):
        raise ValueError("grid / pool / tensor shapes disagree")
    with instrument.region("maxpool_tokens"):
      rc = _lib.lib().vex_maxpool_tokens(x.data_ptr(), x.stride(0), gd, gh, gw, pz, py, px, out.data_ptr(),
                                         out.stride(0), C, _stream())
    _lib.check(rc, "vex_maxpool_tokens")


@torch.library.custom_op("vex::scatter_rows", mutates_args=("out",))
def scatter_rows(x: torch.Tensor, row_src: Optional[torch.Tensor], row_dst: torch.Tensor, out: torch.Tensor) -> None:
    """K11 (vex_scatter_rows): out[row_dst[r]] = x[row_src[r]] (boi / eoi rows visual.py:204-206, feature scatter
    modeling_cogvlm.py:450-453)."""
    _dev(x, "x", _BF16), _dev(out, "out", _BF16), _dev(row_dst, "row_dst", torch.int32)
    H = x.shape[-1]
    if out.shape[-1] != H:
        raise ValueError("row width mismatch")
    if row_src is not None and _dev(row_src, "row_src", torch.int32).numel() != row_dst.numel():
        raise ValueError("row_src / row_dst lengths differ")
    with instrument.region("scatter_rows"):
      rc = _lib.lib().vex_scatter_rows(x.data_ptr(), _ptr(row_src), row_dst.data_ptr(), row_dst.numel(),
                                       out.data_ptr(), H, _stream())
    _lib.check(rc, "vex_scatter_rows")


# ------------------------------------------------------------------------------------------ K8
@torch.library.custom_op("vex::lora_wgrad", mutates_args=("out_vision", "out_language"))
def lora_wgrad(x: torch.Tensor, y: torch.Tensor, out_vision: Optional[torch.Tensor], out_language: Optional[torch.Tensor],
               transpose_out: bool, counts: torch.Tensor) -> None:
    """K8 (vex_lora_wgrad): out_e (+)= x_e^T . y_e over the rows of expert e (fp32 accumulate).  dB = dy^T . T with
    out [F, r]; dA = dT^T . a with ``transpose_out`` and out [r, F]."""
    _dev(x, "x", _BF16), _dev(y, "y", _BF16), _dev(counts, "counts", torch.int32)
    F, r = x.shape[-1], y.shape[-1]
    if x.numel() // F != y.numel() // r:
        raise ValueError("x and y must have the same number of rows")
    shape = (r, F) if transpose_out else (F, r)
    for n, o in (("out_vision", out_vision), ("out_language", out_language)):
        if o is not None:
            _dev(o, n, torch.float32)
            if tuple(o.shape) != shape:
                raise ValueError(f"{n} must be fp32 {shape}")
    if out_vision is None and out_language is None:
        raise ValueError("no output given")
    with instrument.region("lora_wgrad"):
      rc = _lib.lib().vex_lora_wgrad(x.data_ptr(), F, y.data_ptr(), r, r, _ptr(out_vision), _ptr(out_language),
                                     shape[1], int(transpose_out), counts.data_ptr(), x.numel() // F, F, _stream())
    _lib.check(rc, "vex_lora_wgrad")


for _op in (attention_decode_cache, advance_counter, kv_clear_padded, linear_bias_act, attention_blockdiag, layernorm, patchify, maxpool_tokens, scatter_rows, label_rows, lm_head_ce_forward, lm_head_ce_backward, dropout_rows, attention_train, attention_backward, lora_wgrad, gather_rows, silu_mul_backward, rmsnorm_backward, grouped_gemm_dgrad, attention_decode, partition, rmsnorm_gather, silu_mul, residual_scatter, copy_padded_rows, grouped_gemm, _grouped_gemm_fused,
            attention):
    _op.register_fake(lambda *a, **k: None)
