"""Drop-in visual-expert decoder layer backed by the libvex sm_100a kernels.

Mirrors the module surface of the reference (``/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py``):
class names, child-module tree and state-dict keys are the reference's, so
``from_pretrained`` checkpoints, PEFT ``target_modules`` / adapter files and the ``llm_forward`` caller
loop (:547-569) work unchanged:

    self_attn.rotary_emb.inv_freq                                   (64,)      persistent buffer
    self_attn.{vision,language}_expert_query_key_value.weight       (3H, H)
    self_attn.{vision,language}_expert_dense.weight                 (H, H)
    mlp.{language,vision}_mlp.{gate_proj,up_proj}.weight            (I, H)
    mlp.{language,vision}_mlp.down_proj.weight                      (H, I)
    input_layernorm.weight / post_attention_layernorm.weight        (H,)

What differs is ``CogVLMDecoderLayer.forward``: instead of calling its children it reads their tensors
(through PEFT wrappers, ``peft_compat``) and runs the fused pipeline

    K1 partition (once per forward, cached) -> K2 RMSNorm+gather -> K3 QKV GEMM (+LoRA, rotary, scatter to
    token order) -> K4 causal varlen attention (scatter to expert order) -> K3 dense GEMM (+LoRA, residual,
    scatter to [B, L]) -> K2 -> K3 gate/up GEMM (+LoRA, SwiGLU) -> K3 down GEMM (+LoRA, residual, scatter)

with zero host synchronisations.  bf16 CUDA tensors only; anything else raises (no CPU fallback).
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from . import ops
from .peft_compat import LinearSpec, resolve_linear, resolve_norm
from .plan import GLOBAL_PLAN_CACHE, RoutingPlan

try:  # keep the optimiser's no-weight-decay grouping working when luolib is installed (luolib/optim/utils.py:58)
    from luolib.models.param import NoWeightDecayParameter  # type: ignore
except Exception:  # luolib is absent in this image
    class NoWeightDecayParameter(nn.Parameter):
        """Stand-in for ``luolib.models.param.NoWeightDecayParameter`` (a bare ``nn.Parameter`` subclass)."""

LANGUAGE_TOKEN_TYPE = 0  # mmmm/data/utils.py:192
VISION_TOKEN_TYPE = 1    # mmmm/data/utils.py:193
HEAD_DIM = 128


@dataclass
class VexConfig:
    """The fields of ``CogVLMConfig`` (configuration_cogvlm.py:8-45) this layer reads; any config object
    exposing these attributes (e.g. the reference's ``CogVLMConfig``) is accepted as well."""
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_attention_heads: int = 32
    hidden_act: str = "silu"
    max_position_embeddings: int = 2048
    rms_norm_eps: float = 1e-6
    initializer_range: float = 0.02
    num_hidden_layers: int = 32
    lora_lang: bool = True


class RMSNorm(nn.Module):
    """Same parameters as the reference RMSNorm (:30-41).  ``forward`` normalises every row of a
    [..., H] bf16 CUDA tensor with K2 (used for the caller's final norm); inside the decoder layer the
    weight is read directly."""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = NoWeightDecayParameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        x = hidden_states.contiguous()
        rows = x.numel() // x.shape[-1]
        n = torch.full((1,), rows, dtype=torch.int32, device=x.device)
        out = torch.empty_like(x)
        ops.rmsnorm_gather(x.view(rows, -1), self.weight.detach(), self.variance_epsilon, None, n, out.view(rows, -1))
        return out


class MLP(nn.Module):
    """Parameter container with the reference's names (:44-56); the math runs in K3's SwiGLU epilogue."""

    def __init__(self, config):
        super().__init__()
        if getattr(config, "hidden_act", "silu") != "silu":
            raise NotImplementedError("only hidden_act='silu' (configuration_cogvlm.py:15) is implemented")
        self.hidden_size = config.hidden_size
        self.intermediate_size = config.intermediate_size
        self.gate_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.up_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.down_proj = nn.Linear(self.intermediate_size, self.hidden_size, bias=False)
        self.act_fn = nn.SiLU()  # kept so the module tree matches the reference's (ACT2FN['silu'] is a module, :52)


def _apply_prefix(prefix: str, path: str) -> str:  # mmmm/utils.py:8-9
    return f"{prefix}{path}" if prefix.endswith(".") or not prefix else f"{prefix}.{path}"


def _linear_children(module: nn.Module, prefix: str):
    return [_apply_prefix(prefix, n) for n, m in module.named_modules() if isinstance(m, nn.Linear) and n]


class VisionExpertMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.language_mlp = MLP(config)
        self.vision_mlp = MLP(config)

    def get_lora_modules(self, prefix: str):
        """Same selection as the reference hook (:79-85): all six Linears, or the vision expert's three."""
        if getattr(self.config, "lora_lang", True):
            return _linear_children(self, prefix), []
        return _linear_children(self.vision_mlp, _apply_prefix(prefix, "vision_mlp")), []


class RotaryEmbedding(nn.Module):
    """Keeps the reference's table semantics (:145-180): ``inv_freq`` is a persistent buffer and the cos/sin
    cache is built IN ``inv_freq.dtype`` (bf16 arange under bf16-true -- SURVEY section 0 quirk 2), grow-only.
    The tables are handed to the QKV epilogue as tensors; they are sized from a host-known bound instead
    of ``position_ids.max() + 1`` (:255), which would be a device->host sync."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        self.dim = dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, device=device) / self.dim))
        self.register_buffer("inv_freq", inv_freq)
        self.max_seq_len_cached = 0
        self.cos_cached = None
        self.sin_cached = None
        self._tables = {}

    def _set_cos_sin_cache(self, seq_len, device):
        self.max_seq_len_cached = seq_len
        t = torch.arange(seq_len, device=device, dtype=self.inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq.to(device))
        emb = torch.cat((freqs, freqs), dim=-1)
        self.cos_cached = emb.cos()
        self.sin_cached = emb.sin()
        self._tables = {}

    def tables(self, seq_len: int, device, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
        """cos, sin as contiguous [S, dim] tensors of ``dtype`` with S >= seq_len."""
        # rebuilt when the module was cast after the cache was made, so the table is what a fresh run in the
        # current precision sees (the reference builds it lazily, after Lightning's bf16 conversion)
        if (seq_len > self.max_seq_len_cached or self.cos_cached is None or self.cos_cached.device != device
                or self.cos_cached.dtype != self.inv_freq.dtype):
            self._set_cos_sin_cache(max(seq_len, self.max_seq_len_cached, self.max_position_embeddings), device)
        key = (dtype, device)
        if key not in self._tables:
            self._tables[key] = (self.cos_cached.to(dtype).contiguous(), self.sin_cached.to(dtype).contiguous())
        return self._tables[key]

    def forward(self, x, seq_len):  # reference signature (:172-180): [S, 1, dim] slices in x.dtype
        cos, sin = self.tables(int(seq_len), x.device, x.dtype)
        return cos[:seq_len, None, :], sin[:seq_len, None, :]


class VisionExpertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        if self.head_dim != HEAD_DIM or self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError(f"the attention / rotary kernels are specialised for head_dim {HEAD_DIM}, "
                             f"got hidden {self.hidden_size} / heads {self.num_heads}")
        self.max_position_embeddings = getattr(config, "max_position_embeddings", 2048)
        self.rotary_emb = RotaryEmbedding(self.head_dim, self.max_position_embeddings)
        self.vision_expert_query_key_value = nn.Linear(self.hidden_size, self.hidden_size * 3, bias=False)
        self.vision_expert_dense = nn.Linear(self.hidden_size, self.hidden_size, bias=False)
        self.language_expert_query_key_value = nn.Linear(self.hidden_size, self.hidden_size * 3, bias=False)
        self.language_expert_dense = nn.Linear(self.hidden_size, self.hidden_size, bias=False)

    def get_lora_modules(self, prefix: str):
        """Same selection as the reference hook (:211-220)."""
        if getattr(self.config, "lora_lang", True):
            return _linear_children(self, prefix), []
        return [_apply_prefix(prefix, "vision_expert_query_key_value"),
                _apply_prefix(prefix, "vision_expert_dense")], []


def masked_rms_norm(norm: nn.Module, hidden_states: torch.Tensor, token_type_ids: torch.Tensor,
                    padding_mask: torch.Tensor) -> torch.Tensor:
    """``_mask_set(h, pm, norm(h[pm]))`` -- the caller's final norm (modeling_cogvlm.py:570-573, :390-393): rows
    with ``padding_mask == True`` are normalised in place of a fresh copy, the others pass through."""
    plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
    mod = resolve_norm(norm)
    B, L, H = hidden_states.shape
    x = hidden_states.contiguous().view(B * L, H)
    out = torch.empty_like(x)
    ops.copy_padded_rows(x, plan.flat_to_sorted, out)
    ops.rmsnorm_gather(x, mod.weight.detach(), mod.variance_epsilon, plan.token_to_flat, plan.n_valid, out,
                       plan.token_to_flat)
    return out.view(B, L, H)


def get_expert_mask(token_type_ids: torch.Tensor, padding_mask: torch.Tensor):
    """Boolean masks with the reference's meaning (:58-70), derived from the K1 plan without a host sync."""
    plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
    s = plan.flat_to_sorted.view(plan.batch, plan.seq_len)
    tv = plan.counts[0]
    return (s >= 0) & (s < tv), s >= tv


# --------------------------------------------------------------------------------------------------
# fused forward
# --------------------------------------------------------------------------------------------------
_cast_cache: dict = {}


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """Adapter / norm tensors may be fp32 (PEFT ``autocast_adapter_dtype``); the kernels want bf16.  Cached on
    (storage, version) so a frozen tensor is converted once and an optimiser step invalidates the copy."""
    t = t.detach()
    if t.dtype == torch.bfloat16 and t.is_contiguous():
        return t
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype)
    hit = _cast_cache.get(id(t))
    if hit is not None and hit[0] == key:
        return hit[1]
    c = t.to(torch.bfloat16).contiguous()
    if len(_cast_cache) > 512:
        _cast_cache.clear()
    _cast_cache[id(t)] = (key, c, t)
    return c


def _fuse_default() -> bool:
    return os.environ.get("VEX_FUSE_EPILOGUE", "1") != "0"


_dropout_calls = 0


def next_dropout_seed() -> int:
    """Base seed of one layer call's LoRA dropout masks: a function of ``torch.initial_seed()`` and a call counter,
    so runs are reproducible under ``torch.manual_seed`` (no device sync, no generator state consumed)."""
    global _dropout_calls
    _dropout_calls += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _dropout_calls * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def dropout_stream_seed(base_seed: int, stream: int) -> int:
    """Seed of dropout stream ``stream`` (0 qkv, 1 dense, 2 gate, 3 up, 4 down: one nn.Dropout per wrapped Linear)."""
    return (base_seed + (stream + 1) * 0xA0761D6478BD642F) & 0x7FFFFFFFFFFFFFFF  # torch.library ints are int64


def _lora_t(x_sorted: torch.Tensor, specs: Tuple[LinearSpec, LinearSpec], counts: torch.Tensor, *,
            n_valid: Optional[torch.Tensor] = None, seed: Optional[int] = None, stream: int = 0,
            keep: Optional[dict] = None, name: str = ""):
    """T = scaling * dropout(x) . lora_A^T for the (vision, language) pair; returns (T or None, r, [B_v, B_l]).
    With an active ``lora_dropout`` (wrapper in training mode) the LoRA branch reads a dropped copy of x
    (PEFT: lora_B(lora_A(dropout(x))) * scaling), kept in ``keep[name + '_xd']`` for the weight gradient."""
    sv, sl = specs
    if sv.lora_A is None and sl.lora_A is None:
        return None, 0, [None, None]
    if sv.lora_A is None:
        raise NotImplementedError("LoRA on the language expert only is not supported (the reference's lora_lang "
                                  "switch adds the language adapters on top of the vision ones)")
    r = sv.r
    if r % 8 or r > 64:
        raise NotImplementedError(f"LoRA rank {r}: the fused K-extension handles multiples of 8 up to 64")
    both = sl.lora_A is not None
    if both and (sl.r != r or sl.scaling != sv.scaling or sl.dropout != sv.dropout):
        raise NotImplementedError("vision and language adapters must share rank, scaling and dropout")
    x_in = x_sorted
    if sv.dropout > 0:
        if seed is None or n_valid is None:
            raise RuntimeError("LoRA dropout is active but no seed was provided")
        x_in = torch.empty_like(x_sorted)
        ops.dropout_rows(x_sorted, n_valid, x_in, sv.dropout, dropout_stream_seed(seed, stream))
        if keep is not None:
            keep[name + "_xd"] = x_in
    t = torch.empty(x_sorted.shape[0], r, dtype=torch.bfloat16, device=x_sorted.device)
    ops.grouped_gemm(x_in, _bf16(sv.lora_A), _bf16(sl.lora_A) if both else None, t, counts, None,
                     float(sv.scaling))
    return t, r, [_bf16(sv.lora_B), _bf16(sl.lora_B) if both else None]


def visual_expert_layer_forward(layer: "CogVLMDecoderLayer", hidden_states: torch.Tensor, plan: RoutingPlan,
                                position_ids: torch.Tensor, *, use_cache: bool = False,
                                fuse_epilogue: Optional[bool] = None, keep: Optional[dict] = None,
                                dropout_seed: Optional[int] = None):
    """The whole layer on the device; returns (out [B, L, H], present_kv or None).  ``keep`` (training recompute):
    a dict that receives the intermediates the backward needs (gate/up are then materialised separately, the
    attention also writes its log-sum-exp, and the call returns (None, None) right before the down projection)."""
    fuse = _fuse_default() if fuse_epilogue is None else fuse_epilogue
    attn, mlp = layer.self_attn, layer.mlp
    B, L, H = hidden_states.shape
    cap, heads = B * L, attn.num_heads
    I = mlp.vision_mlp.intermediate_size
    dev = hidden_states.device
    hf = hidden_states.view(cap, H)
    counts, s2f = plan.counts, plan.sorted_to_flat
    new = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=dev)

    ln1, ln2 = resolve_norm(layer.input_layernorm), resolve_norm(layer.post_attention_layernorm)
    qkv_s = (resolve_linear(attn.vision_expert_query_key_value), resolve_linear(attn.language_expert_query_key_value))
    dense_s = (resolve_linear(attn.vision_expert_dense), resolve_linear(attn.language_expert_dense))
    gate_s = (resolve_linear(mlp.vision_mlp.gate_proj), resolve_linear(mlp.language_mlp.gate_proj))
    up_s = (resolve_linear(mlp.vision_mlp.up_proj), resolve_linear(mlp.language_mlp.up_proj))
    down_s = (resolve_linear(mlp.vision_mlp.down_proj), resolve_linear(mlp.language_mlp.down_proj))
    W = lambda pair: [_bf16(pair[0].weight), None, _bf16(pair[1].weight), None]
    if dropout_seed is None and any(sp[0].dropout > 0 for sp in (qkv_s, dense_s, gate_s, up_s, down_s)):
        dropout_seed = next_dropout_seed()  # wrappers in training mode: nn.Dropout would be active
    lt = lambda x, sp, k, nm: _lora_t(x, sp, counts, n_valid=plan.n_valid, seed=dropout_seed, stream=k, keep=keep,
                                      name=nm)

    # ---- attention block ----
    xn = new(cap, H)
    ops.rmsnorm_gather(hf, ln1.weight.detach(), ln1.variance_epsilon, s2f, plan.n_valid, xn)
    cos, sin = attn.rotary_emb.tables(max(L, attn.max_position_embeddings), dev, torch.bfloat16)
    pos_flat = position_ids.reshape(-1)
    qkv = new(cap, 3 * H)  # token order: row t = [q(heads*128) | k | v], q and k rotated
    t, r, lb = lt(xn, qkv_s, 0, "qkv")
    if keep is not None:
        keep.update(xn1=xn, t_qkv=t, specs=dict(qkv=qkv_s, dense=dense_s, gate=gate_s, up=up_s, down=down_s),
                    ln1=ln1, ln2=ln2, cos=cos, sin=sin, dropout_seed=dropout_seed)
        xn = new(cap, H)  # the second norm gets its own buffer (xn1 is needed by the backward)
    ops.grouped_gemm_fused(keep["xn1"] if keep is not None else xn, W(qkv_s), qkv, counts, ops.EPI_ROPE, plan.sorted_to_token, None, [t, None],
                           [lb[0], None, lb[1], None], r, [cos, sin, pos_flat, s2f], 2 * H, False, 1.0)
    ctx = new(cap, H)      # expert-sorted order again: the A operand of the dense GEMM
    if keep is not None:   # training recompute: the backward kernels need the log-sum-exp
        keep["lse"] = torch.empty(heads, cap, dtype=torch.float32, device=dev)
        ops.attention_train(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, ctx, HEAD_DIM ** -0.5, keep["lse"])
    else:
        ops.attention(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, ctx, HEAD_DIM ** -0.5)
    h1 = new(B, L, H)
    ops.copy_padded_rows(hf, plan.flat_to_sorted, h1.view(cap, H))
    t, r, lb = lt(ctx, dense_s, 1, "dense")
    if keep is not None:
        keep.update(qkv=qkv, ctx=ctx, t_dense=t)
    if fuse:
        ops.grouped_gemm_fused(ctx, W(dense_s), h1.view(cap, H), counts, ops.EPI_RESIDUAL, s2f, hf, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
    else:
        y = new(cap, H)
        ops.grouped_gemm_fused(ctx, W(dense_s), y, counts, ops.EPI_PLAIN, None, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        ops.residual_scatter(y, hf, s2f, plan.n_valid, h1.view(cap, H))

    # ---- MLP block ----
    ops.rmsnorm_gather(h1.view(cap, H), ln2.weight.detach(), ln2.variance_epsilon, s2f, plan.n_valid, xn)
    act = new(cap, I)
    tg, rg, lbg = lt(xn, gate_s, 2, "gate")
    tu, ru, lbu = lt(xn, up_s, 3, "up")
    if (tg is None) != (tu is None) or rg != ru:
        raise NotImplementedError("gate_proj and up_proj adapters must come in pairs of equal rank")
    if keep is not None:
        keep.update(xn2=xn, t_gate=tg, t_up=tu)
    if fuse and keep is None:
        w4 = [_bf16(gate_s[0].weight), _bf16(up_s[0].weight), _bf16(gate_s[1].weight), _bf16(up_s[1].weight)]
        ops.grouped_gemm_fused(xn, w4, act, counts, ops.EPI_SWIGLU, None, None, [tg, tu],
                               [lbg[0], lbu[0], lbg[1], lbu[1]], rg, [], 0, False, 1.0)
    else:
        g, u = new(cap, I), new(cap, I)
        ops.grouped_gemm_fused(xn, W(gate_s), g, counts, ops.EPI_PLAIN, None, None, [tg, None],
                               [lbg[0], None, lbg[1], None], rg, [], 0, False, 1.0)
        ops.grouped_gemm_fused(xn, W(up_s), u, counts, ops.EPI_PLAIN, None, None, [tu, None],
                               [lbu[0], None, lbu[1], None], ru, [], 0, False, 1.0)
        ops.silu_mul(g, u, plan.n_valid, act)
        if keep is not None:
            keep.update(gate=g, up=u)
    t, r, lb = lt(act, down_s, 4, "down")
    if keep is not None:
        keep.update(act=act, t_down=t, h1=h1)
        if not keep.get("continue_forward", False):
            return None, None  # the recompute stops here: the down projection's output is not needed by the backward
        # activation-keeping training forward (no recompute in the backward): h1 is needed by the backward, so the
        # down projection writes a fresh buffer instead of accumulating in place
        out = new(B, L, H)
        ops.copy_padded_rows(hf, plan.flat_to_sorted, out.view(cap, H))
        ops.grouped_gemm_fused(act, W(down_s), out.view(cap, H), counts, ops.EPI_RESIDUAL, s2f, h1.view(cap, H),
                               [t, None], [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        return out, None
    if fuse:  # in place: h1[row] += down(act)[row]; padded rows of h1 already hold the input
        ops.grouped_gemm_fused(act, W(down_s), h1.view(cap, H), counts, ops.EPI_RESIDUAL, s2f, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        out = h1
    else:
        y = new(cap, H)
        ops.grouped_gemm_fused(act, W(down_s), y, counts, ops.EPI_PLAIN, None, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        out = h1.clone()
        ops.residual_scatter(y, h1.view(cap, H), s2f, plan.n_valid, out.view(cap, H))

    present = None
    if use_cache:  # post-rotary k and v as [B, heads, L, 128], zeros at padded positions (:243, :262)
        f2s = plan.flat_to_sorted.long()
        valid = f2s >= 0
        tok = plan.sorted_to_token.long()[f2s.clamp_min(0)]
        kv = qkv.view(cap, 3, heads, HEAD_DIM)[tok, 1:] * valid[:, None, None, None]
        kv = kv.view(B, L, 2, heads, HEAD_DIM).permute(2, 0, 3, 1, 4)
        present = (kv[0], kv[1])
    return out, present


_decode_consts: dict = {}


def _decode_constants(batch: int, device):
    """Device constants of a decode step: counts = [B, 0, B, 1] (all rows go through ONE weight set) and the
    identity row map.  Built once per (batch, device): no per-step host->device copy."""
    key = (batch, device)
    if key not in _decode_consts:
        _decode_consts[key] = (torch.tensor([batch, 0, batch, 1], dtype=torch.int32, device=device),
                               torch.arange(batch, dtype=torch.int32, device=device))
    return _decode_consts[key]


def _lora_t_single(x: torch.Tensor, spec: LinearSpec, counts: torch.Tensor):
    if spec.lora_A is None:
        return None, 0, None
    if spec.r % 8 or spec.r > 64:
        raise NotImplementedError(f"LoRA rank {spec.r}: the fused K-extension handles multiples of 8 up to 64")
    t = torch.empty(x.shape[0], spec.r, dtype=torch.bfloat16, device=x.device)
    ops.grouped_gemm(x, _bf16(spec.lora_A), None, t, counts, None, float(spec.scaling))
    return t, spec.r, _bf16(spec.lora_B)


def visual_expert_layer_decode(layer: "CogVLMDecoderLayer", hidden_states: torch.Tensor, position_ids: torch.Tensor,
                               padding_mask: torch.Tensor, past_key_value, use_cache: bool = True):
    """One generation step (q_len == 1 with a KV cache): the reference's L == 1 rules -- every token goes to the
    LANGUAGE expert regardless of padding (get_expert_mask :67), plain RMSNorm on every row (:308-309, :327-328),
    cache concat on dim 2 (:258-260) and the generation branch of attention_fn (:129-141)."""
    attn, mlp = layer.self_attn, layer.mlp
    B, L, H = hidden_states.shape
    heads = attn.num_heads
    I = mlp.language_mlp.intermediate_size
    dev = hidden_states.device
    past_k, past_v = past_key_value
    if past_k.shape[0] != B or past_k.shape[1] != heads or past_k.shape[3] != HEAD_DIM:
        raise ValueError(f"past_key_value must be [B, {heads}, L_past, {HEAD_DIM}]")
    Lkv = past_k.shape[2] + 1
    if padding_mask.shape != (B, Lkv):
        raise ValueError(f"padding_mask must cover past + current positions: expected {(B, Lkv)}")
    counts, ident = _decode_constants(B, dev)
    n_rows = counts[2:3]
    new = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=dev)
    hf = hidden_states.view(B, H)
    ln1, ln2 = resolve_norm(layer.input_layernorm), resolve_norm(layer.post_attention_layernorm)
    qkv_s = resolve_linear(attn.language_expert_query_key_value)
    dense_s = resolve_linear(attn.language_expert_dense)
    gate_s, up_s = resolve_linear(mlp.language_mlp.gate_proj), resolve_linear(mlp.language_mlp.up_proj)
    down_s = resolve_linear(mlp.language_mlp.down_proj)

    def gemm(a, w, out, mode, spec_pair, residual=None, rope=(), rope_cols=0):
        t, r, lb = [None, None], 0, [None] * 4
        for h, sp in enumerate(spec_pair):
            th, rh, bh = _lora_t_single(a, sp, counts)
            if th is not None:
                t[h], r, lb[h] = th, rh, bh
        if len(spec_pair) == 2 and (t[0] is None) != (t[1] is None):
            raise NotImplementedError("gate_proj and up_proj adapters must come in pairs")
        ops.grouped_gemm_fused(a, w, out, counts, mode, None, residual, t, lb, r, list(rope), rope_cols, True, 1.0)

    xn = new(B, H)
    ops.rmsnorm_gather(hf, ln1.weight.detach(), ln1.variance_epsilon, None, n_rows, xn)
    max_pos = max(Lkv, attn.max_position_embeddings)
    cos, sin = attn.rotary_emb.tables(max_pos, dev, torch.bfloat16)
    qkv = new(B, 3 * H)
    gemm(xn, [_bf16(qkv_s.weight)], qkv, ops.EPI_ROPE, [qkv_s],
         rope=(cos, sin, position_ids.reshape(-1), ident), rope_cols=2 * H)
    k_new = qkv[:, H:2 * H].reshape(B, heads, 1, HEAD_DIM)
    v_new = qkv[:, 2 * H:].reshape(B, heads, 1, HEAD_DIM)
    k = torch.cat([past_k, k_new], dim=2)   # :259-260 (the tuple-cache API reallocates every step)
    v = torch.cat([past_v, v_new], dim=2)
    ctx = new(B, H)
    ops.attention_decode(qkv[:, :H], k, v, padding_mask, ctx, HEAD_DIM ** -0.5)
    h1 = new(B, H)
    gemm(ctx, [_bf16(dense_s.weight)], h1, ops.EPI_RESIDUAL, [dense_s], residual=hf)
    ops.rmsnorm_gather(h1, ln2.weight.detach(), ln2.variance_epsilon, None, n_rows, xn)
    act = new(B, I)
    gemm(xn, [_bf16(gate_s.weight), _bf16(up_s.weight)], act, ops.EPI_SWIGLU, [gate_s, up_s])
    gemm(act, [_bf16(down_s.weight)], h1, ops.EPI_RESIDUAL, [down_s])  # in place: h1 += down(act)
    return h1.view(B, 1, H), ((k, v) if use_cache else None)


class CogVLMDecoderLayer(nn.Module):
    """Drop-in for the reference ``CogVLMDecoderLayer`` (:286-340)."""

    def __init__(self, config):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = VisionExpertAttention(config=config)
        self.mlp = VisionExpertMLP(config)
        self.input_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.fuse_epilogue: Optional[bool] = None  # None -> VEX_FUSE_EPILOGUE env (default on)

    def forward(
        self,
        hidden_states: torch.Tensor,
        token_type_ids: torch.LongTensor = None,
        position_ids: torch.LongTensor = None,
        padding_mask: Optional[torch.BoolTensor] = None,
        past_key_value: Optional[Tuple[torch.Tensor]] = None,
        output_attentions: Optional[bool] = False,
        use_cache: Optional[bool] = False,
        attention_mask: Optional[torch.Tensor] = None,
    ):
        if padding_mask is None:
            padding_mask = attention_mask  # BASELINE wording; the reference converts one level up (:539)
        if token_type_ids is None or position_ids is None or padding_mask is None:
            raise TypeError("token_type_ids, position_ids and padding_mask (or attention_mask) are required")
        decode = past_key_value is not None
        if decode and hidden_states.dim() == 3 and hidden_states.shape[1] != 1:
            raise NotImplementedError("past_key_value with q_len > 1 (chunked prefill) is not implemented; the "
                                      "reference's generation branch asserts q_len == 1 as well (:131)")
        if not hidden_states.is_cuda:
            raise ValueError("hidden_states must be a CUDA tensor: the visual-expert layer has no CPU path")
        if hidden_states.dtype != torch.bfloat16:
            raise TypeError(f"hidden_states must be bfloat16 (bf16-true, mmmm.py:468-492), got {hidden_states.dtype}")
        if hidden_states.dim() != 3 or hidden_states.shape[-1] != self.hidden_size:
            raise ValueError(f"hidden_states must be [B, L, {self.hidden_size}]")
        if hidden_states.shape[:2] != token_type_ids.shape or position_ids.shape != token_type_ids.shape:
            raise ValueError("token_type_ids / position_ids must be [B, L] like hidden_states")
        if not decode and hidden_states.shape[1] == 1:
            raise NotImplementedError("q_len == 1 without a KV cache is not part of the prefill path")
        train = torch.is_grad_enabled() and (hidden_states.requires_grad or any(
            p.requires_grad for p in self.parameters()))
        if train and (decode or use_cache):
            raise NotImplementedError("autograd through the decode / use_cache paths is not implemented")
        if output_attentions:
            warnings.warn("output_attentions is not implemented.")  # same as the reference (:281-282)
        hidden_states = hidden_states.contiguous()
        if position_ids.dtype != torch.int64:
            position_ids = position_ids.long()
        if decode:
            out, present = visual_expert_layer_decode(self, hidden_states, position_ids.contiguous(),
                                                      padding_mask.bool().contiguous(), past_key_value,
                                                      use_cache=bool(use_cache))
        elif train:  # LoRA training step (BASELINE config 5): fused forward, self-checkpointing backward
            from .training import layer_forward_train
            plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
            out, present = layer_forward_train(self, hidden_states, plan, position_ids.contiguous()), None
        else:
            plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
            out, present = visual_expert_layer_forward(self, hidden_states, plan, position_ids.contiguous(),
                                                       use_cache=bool(use_cache), fuse_epilogue=self.fuse_epilogue)
        outputs = (out,)
        if output_attentions:
            outputs += (None,)
        if use_cache:
            outputs += (present,)
        return outputs


def swap_decoder_layers(model: nn.Module) -> nn.Module:
    """Replaces every reference ``CogVLMDecoderLayer`` in ``model.layers`` (a ``CogVLMModel``, :402) by the
    B200 layer, moving the parameters over (no copy) so state-dict keys and tensor identity are preserved."""
    layers = model.layers
    for i, old in enumerate(layers):
        if isinstance(old, CogVLMDecoderLayer):
            continue
        cfg = old.self_attn.config
        new = CogVLMDecoderLayer(cfg)
        sd = dict(old.named_parameters())
        for name, _ in list(new.named_parameters()):
            mod_path, _, pname = name.rpartition(".")
            setattr(new.get_submodule(mod_path), pname, sd[name])
        new.self_attn.rotary_emb.inv_freq = old.self_attn.rotary_emb.inv_freq
        layers[i] = new
    return model


class VisualExpertDecoder(nn.Module):
    """The decoder part of the reference ``CogVLMModel``: ``layers`` + final ``norm`` driven like
    ``CogVLMModel.llm_forward`` (modeling_cogvlm.py:477-586) from ``inputs_embeds`` (embedding lookup and the
    vision encoder stay with the caller -- they are outside the hot path).  State-dict keys ``layers.N.*`` and
    ``norm.weight`` equal those of ``CogVLMModel``, so its checkpoint loads with ``strict=False``.

    The routing plan (K1) is computed once and shared by all layers; with ``graph=True`` the whole prefill is
    captured into one CUDA graph per input shape (``GraphedPrefill``)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([CogVLMDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._graphs = {}

    def llm_forward(self, inputs_embeds: torch.Tensor, token_type_ids: torch.Tensor,
                    attention_mask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                    past_key_values=None, use_cache: bool = False, graph: bool = False):
        """Returns ``(last_hidden_state, next_cache)`` -- the non-dict return of the reference (:579-580)."""
        B, L, _ = inputs_embeds.shape
        dev = inputs_embeds.device
        past_len = 0 if past_key_values is None else past_key_values[0][0].shape[2]
        if position_ids is None:  # :523-528
            position_ids = torch.arange(past_len, L + past_len, dtype=torch.long, device=dev).unsqueeze(0).expand(B, L)
        position_ids = position_ids.reshape(-1, L).long().contiguous()
        if attention_mask is None:  # :535-538
            attention_mask = torch.ones(B, L + past_len, dtype=torch.bool, device=dev)
        padding_mask = attention_mask.bool()  # :539
        if graph and past_key_values is None and not use_cache:
            from .graph import GraphedPrefill
            key = (B, L, inputs_embeds.dtype, dev)
            if key not in self._graphs:
                self._graphs[key] = GraphedPrefill(self.layers, inputs_embeds, token_type_ids, position_ids,
                                                   padding_mask, final_norm=self.norm)
            return self._graphs[key](inputs_embeds, token_type_ids, position_ids, padding_mask), None
        h = inputs_embeds
        cache = () if use_cache else None
        for i, layer in enumerate(self.layers):  # :547-569
            out = layer(h, token_type_ids=token_type_ids, position_ids=position_ids, padding_mask=padding_mask,
                        past_key_value=None if past_key_values is None else past_key_values[i], use_cache=use_cache)
            h = out[0]
            if use_cache:
                cache += (out[1],)
        if L > 1:  # :570-573
            h = masked_rms_norm(self.norm, h, token_type_ids, padding_mask)
        else:
            h = self.norm(h)
        return h, cache

    forward = llm_forward
