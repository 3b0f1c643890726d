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
    """Adapter / norm tensors may be fp32 (PEFT ``autocast_adapter_dtype``); the kernels want bf16.  The bf16 copy is
    cached per SOURCE tensor object (the Parameter the wrapper holds -- its identity is stable across calls, unlike
    the result of ``detach()``) and revalidated on (storage, version, dtype, shape), so a frozen tensor is converted
    once and an optimiser step (in-place update -> version bump) refreshes the copy."""
    if t.dtype == torch.bfloat16 and t.is_contiguous():
        return t.detach()
    sig = (t.data_ptr(), None if t.is_inference() else t._version, t.dtype, tuple(t.shape))
    hit = _cast_cache.get(id(t))
    if hit is not None and hit[0] is t and hit[1] == sig:
        return hit[2]
    c = t.detach().to(torch.bfloat16).contiguous()
    if len(_cast_cache) > 4096:  # 32 layers x 20 adapter tensors = 640 live entries; stale ones are dropped wholesale
        _cast_cache.clear()
    _cast_cache[id(t)] = (t, sig, c)  # holding `t` keeps id(t) from being recycled by another tensor
    return c


def _base_weight(t: torch.Tensor) -> torch.Tensor:
    """A frozen base ``nn.Linear`` weight as the TMA operand: must already be contiguous bf16 (bf16-true,
    mmmm.py:468-492).  An fp32 base weight is rejected instead of being converted on every call (one 4096 x 11008
    matrix is 90 MB per copy)."""
    if t.dtype != torch.bfloat16:
        raise TypeError(f"base Linear weights must be bfloat16 (bf16-true model, mmmm.py:468-492), got {t.dtype}; "
                        f"cast the model once with .to(torch.bfloat16)")
    if not t.is_contiguous():
        raise ValueError("base Linear weights must be contiguous [out_features, in_features]")
    return t.detach()


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
                                dropout_seed: Optional[int] = None, sorted_stream: bool = False,
                                kv_capacity: Optional[int] = None, kv_out: Optional[tuple] = None):
    """The whole layer on the device; returns (out, present_kv or None).

    ``sorted_stream=False`` (the drop-in module call): ``hidden_states`` / ``out`` are [B, L, H] in the reference's
    flat layout; rows are gathered into expert-sorted order by K2 and scattered back by the residual epilogues.
    ``sorted_stream=True`` (used by ``decoder_stack_forward``): ``hidden_states`` is the residual
    stream ALREADY in expert-sorted order, [B*L, H]; the layer updates it IN PLACE (h += attn(norm(h)); h +=
    mlp(norm(h))) with identity row maps -- no gather, no scatter, no padded-row copy, no second activation buffer --
    and returns the same tensor.

    ``keep`` (training recompute, flat layout only): a dict that receives the intermediates the backward needs
    (gate/up are then materialised separately, the attention also writes its log-sum-exp, and the call returns
    (None, None) right before the down projection).
    ``use_cache``: post-rotary K and V are written by the QKV epilogue straight into a [B, heads, kv_capacity, 128]
    cache pair (zeros at padded positions, :243, :262) -- ``kv_out`` = (k, v) if the caller owns the cache, otherwise
    allocated here with ``KV_HEADROOM`` spare positions; ``present`` = views of the first L positions."""
    fuse = _fuse_default() if fuse_epilogue is None else fuse_epilogue
    attn, mlp = layer.self_attn, layer.mlp
    B, L = plan.batch, plan.seq_len
    H = hidden_states.shape[-1]
    cap, heads = B * L, attn.num_heads
    I = mlp.vision_mlp.intermediate_size
    dev = hidden_states.device
    hf = hidden_states.view(cap, H)
    counts, s2f = plan.counts, plan.sorted_to_flat
    if sorted_stream and (keep is not None or not fuse):
        raise NotImplementedError("the sorted residual stream is the fused inference path (no keep / unfused mode)")
    # row maps of the residual stream: flat layout goes through sorted_to_flat, the sorted stream is the identity
    stream_map = None if sorted_stream else s2f
    new = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=dev)

    ln1, ln2 = resolve_norm(layer.input_layernorm), resolve_norm(layer.post_attention_layernorm)
    qkv_s = (resolve_linear(attn.vision_expert_query_key_value), resolve_linear(attn.language_expert_query_key_value))
    dense_s = (resolve_linear(attn.vision_expert_dense), resolve_linear(attn.language_expert_dense))
    gate_s = (resolve_linear(mlp.vision_mlp.gate_proj), resolve_linear(mlp.language_mlp.gate_proj))
    up_s = (resolve_linear(mlp.vision_mlp.up_proj), resolve_linear(mlp.language_mlp.up_proj))
    down_s = (resolve_linear(mlp.vision_mlp.down_proj), resolve_linear(mlp.language_mlp.down_proj))
    W = lambda pair: [_base_weight(pair[0].weight), None, _base_weight(pair[1].weight), None]
    if dropout_seed is None and any(sp[0].dropout > 0 for sp in (qkv_s, dense_s, gate_s, up_s, down_s)):
        dropout_seed = next_dropout_seed()  # wrappers in training mode: nn.Dropout would be active
    lt = lambda x, sp, k, nm: _lora_t(x, sp, counts, n_valid=plan.n_valid, seed=dropout_seed, stream=k, keep=keep,
                                      name=nm)

    # ---- attention block ----
    xn = new(cap, H)
    ops.rmsnorm_gather(hf, ln1.weight.detach(), ln1.variance_epsilon, stream_map, plan.n_valid, xn)
    cos, sin = attn.rotary_emb.tables(max(L, attn.max_position_embeddings), dev, torch.bfloat16)
    pos_flat = position_ids.reshape(-1)
    _debug_check_positions(pos_flat, cos.shape[0])
    qkv = new(cap, 3 * H)  # token order: row t = [q(heads*128) | k | v], q and k rotated
    t, r, lb = lt(xn, qkv_s, 0, "qkv")
    if keep is not None:
        keep.update(xn1=xn, t_qkv=t, specs=dict(qkv=qkv_s, dense=dense_s, gate=gate_s, up=up_s, down=down_s),
                    ln1=ln1, ln2=ln2, cos=cos, sin=sin, dropout_seed=dropout_seed)
        xn = new(cap, H)  # the second norm gets its own buffer (xn1 is needed by the backward)
    present, kv = None, None
    if use_cache:
        # a9: K (post-rotary) and V leave the QKV epilogue in the reference's cache layout [B, heads, L_cap, 128]; padded
        # positions are zeroed by one small kernel (the reference multiplies by the mask, :243)
        if kv_out is not None:
            kv = kv_out
        else:
            kcap = max(int(kv_capacity or 0), L) if kv_capacity else L + KV_HEADROOM
            kv = torch.empty(2, B, heads, kcap, HEAD_DIM, dtype=torch.bfloat16, device=dev)
        ops.kv_clear_padded(kv[0], kv[1], plan.flat_to_sorted, B, L)
        present = (kv[0][:, :, :L], kv[1][:, :, :L])
    ops.grouped_gemm_fused(keep["xn1"] if keep is not None else xn, W(qkv_s), qkv, counts, ops.EPI_ROPE,
                           plan.sorted_to_token, None, [t, None], [lb[0], None, lb[1], None], r,
                           [cos, sin, pos_flat, s2f], 2 * H, False, 1.0,
                           None if kv is None else kv[0], None if kv is None else kv[1], L, None)
    ctx = new(cap, H)      # expert-sorted order again: the A operand of the dense GEMM
    if keep is not None:   # training recompute: the backward kernels need the log-sum-exp
        keep["lse"] = torch.empty(heads, cap, dtype=torch.float32, device=dev)
        ops.attention_train(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, ctx, HEAD_DIM ** -0.5, keep["lse"])
    else:
        ops.attention(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, ctx, HEAD_DIM ** -0.5)
    if sorted_stream:
        h1 = hidden_states          # in place: h[r] += dense(ctx)[r]
    else:
        h1 = new(B, L, H)
        ops.copy_padded_rows(hf, plan.flat_to_sorted, h1.view(cap, H))
    h1f = h1.view(cap, H)
    t, r, lb = lt(ctx, dense_s, 1, "dense")
    if keep is not None:
        keep.update(qkv=qkv, ctx=ctx, t_dense=t)
    if fuse:
        ops.grouped_gemm_fused(ctx, W(dense_s), h1f, counts, ops.EPI_RESIDUAL, stream_map,
                               None if sorted_stream else hf, [t, None], [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
    else:
        y = new(cap, H)
        ops.grouped_gemm_fused(ctx, W(dense_s), y, counts, ops.EPI_PLAIN, None, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        ops.residual_scatter(y, hf, s2f, plan.n_valid, h1f)

    # ---- MLP block ----
    ops.rmsnorm_gather(h1f, ln2.weight.detach(), ln2.variance_epsilon, stream_map, plan.n_valid, xn)
    act = new(cap, I)
    tg, rg, lbg = lt(xn, gate_s, 2, "gate")
    tu, ru, lbu = lt(xn, up_s, 3, "up")
    if (tg is None) != (tu is None) or rg != ru:
        raise NotImplementedError("gate_proj and up_proj adapters must come in pairs of equal rank")
    if keep is not None:
        keep.update(xn2=xn, t_gate=tg, t_up=tu)
    if fuse and keep is None:
        w4 = [_base_weight(gate_s[0].weight), _base_weight(up_s[0].weight), _base_weight(gate_s[1].weight),
              _base_weight(up_s[1].weight)]
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
        ops.grouped_gemm_fused(act, W(down_s), out.view(cap, H), counts, ops.EPI_RESIDUAL, s2f, h1f,
                               [t, None], [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        return out, None
    if fuse:  # in place: h1[row] += down(act)[row]; (flat layout) padded rows of h1 already hold the input
        ops.grouped_gemm_fused(act, W(down_s), h1f, counts, ops.EPI_RESIDUAL, stream_map, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        out = h1
    else:
        y = new(cap, H)
        ops.grouped_gemm_fused(act, W(down_s), y, counts, ops.EPI_PLAIN, None, None, [t, None],
                               [lb[0], None, lb[1], None], r, [], 0, False, 1.0)
        out = h1.clone()
        ops.residual_scatter(y, h1f, s2f, plan.n_valid, out.view(cap, H))
    return out, present


def decoder_stack_forward(layers, final_norm: Optional[nn.Module], inputs_embeds: torch.Tensor, plan: RoutingPlan,
                          position_ids: torch.Tensor, *, use_cache: bool = False, kv_capacity: Optional[int] = None,
                          kv_out=None):
    """``CogVLMModel.llm_forward``'s layer loop + final norm (modeling_cogvlm.py:547-573) with the residual stream kept
    in expert-sorted order across ALL layers (routing is layer-invariant): ONE gather at entry, per
    layer two in-place residual GEMM epilogues on the sorted stream, ONE scatter at exit fused into the final masked
    RMSNorm (``_mask_set(h, pm, norm(h[pm]))``).  Rows with ``padding_mask == False`` never enter the stream; they come
    out as the input embeddings (the reference leaves them at whatever the masked assignments skipped).
    Returns (last_hidden_state [B, L, H], tuple of per-layer (k, v) or None)."""
    B, L, H = inputs_embeds.shape
    cap = B * L
    x = inputs_embeds.view(cap, H)
    stream = torch.empty(cap, H, dtype=inputs_embeds.dtype, device=inputs_embeds.device)
    ops.gather_rows(x, plan.sorted_to_flat, plan.n_valid, stream)
    cache = () if use_cache else None
    for i, layer in enumerate(layers):
        _, present = visual_expert_layer_forward(layer, stream, plan, position_ids, use_cache=use_cache,
                                                 sorted_stream=True, kv_capacity=kv_capacity,
                                                 kv_out=None if kv_out is None else kv_out[i])
        if use_cache:
            cache += (present,)
    out = torch.empty_like(inputs_embeds)
    of = out.view(cap, H)
    ops.copy_padded_rows(x, plan.flat_to_sorted, of)
    if final_norm is not None:
        mod = resolve_norm(final_norm)
        ops.rmsnorm_gather(stream, mod.weight.detach(), mod.variance_epsilon, None, plan.n_valid, of,
                           plan.sorted_to_flat)
    else:
        ops.scatter_rows(stream, None, plan.sorted_to_flat, of)
    return out, cache


def _debug_check_positions(position_ids: torch.Tensor, table_len: int) -> None:
    """The rotary tables are sized from a host-known bound, max(L, max_position_embeddings, largest length seen), where
    the reference grows them to ``position_ids.max() + 1`` with a device->host sync (:255, :174); the kernels clamp
    positions into the table.  Real VividMed position ids never exceed the sequence length (data/utils.py:111-124),
    so the bound holds by construction; VEX_DEBUG_POSITIONS=1 verifies it (one host sync per call)."""
    if os.environ.get("VEX_DEBUG_POSITIONS", "0") == "1" and not torch.cuda.is_current_stream_capturing():
        hi = int(position_ids.max())
        if hi >= table_len or int(position_ids.min()) < 0:
            raise ValueError(f"position_ids reach {hi}, outside the rotary table of {table_len} rows: raise "
                             f"config.max_position_embeddings (the kernels clamp instead of growing the table)")


_decode_consts: dict = {}


def _decode_constants(batch: int, device):
    """Device constants of a decode step: counts = [B, 0, B, 1] (all rows go through ONE weight set) and the
    identity row map.  Built once per (batch, device): no per-step host->device copy."""
    key = (batch, device)
    if key not in _decode_consts:
        _decode_consts[key] = (torch.tensor([batch, 0, batch, 1], dtype=torch.int32, device=device),
                               torch.arange(batch, dtype=torch.int32, device=device))
    return _decode_consts[key]


def _lora_t_single(x: torch.Tensor, spec: LinearSpec, counts: torch.Tensor, skinny: bool = False):
    if spec.lora_A is None:
        return None, 0, None
    if spec.r % 8 or spec.r > 64:
        raise NotImplementedError(f"LoRA rank {spec.r}: the fused K-extension handles multiples of 8 up to 64")
    t = torch.empty(x.shape[0], spec.r, dtype=torch.bfloat16, device=x.device)
    if skinny and spec.r % 16 == 0:
        ops.grouped_gemm_raw(x, [_bf16(spec.lora_A)], t, counts, ops.EPI_PLAIN, single_expert=True,
                             alpha=float(spec.scaling), skinny=True)
    else:
        ops.grouped_gemm(x, _bf16(spec.lora_A), None, t, counts, None, float(spec.scaling))
    return t, spec.r, _bf16(spec.lora_B)


def decode_core(layer: "CogVLMDecoderLayer", hf: torch.Tensor, position_ids: torch.Tensor, k_cache: torch.Tensor,
                v_cache: torch.Tensor, mask: torch.Tensor, kv_pos: torch.Tensor) -> torch.Tensor:
    """One generation step of one layer against a PRE-ALLOCATED cache: ``hf`` [B, H]; ``k_cache`` / ``v_cache``
    [B, heads, capacity, 128]; ``kv_pos`` int32 [1] on the device = positions already cached; ``mask`` bool
    [B, >= kv_pos + 1].  The reference's L == 1 rules: every token goes to the LANGUAGE expert regardless of padding
    (get_expert_mask :67), plain RMSNorm on every row (:308-309, :327-328), the new K / V are appended at position
    kv_pos (the ``torch.cat`` of :258-260, done in place by the QKV epilogue) and the generation branch of
    attention_fn (:129-141) runs over positions [0, kv_pos].  No host synchronisation, nothing read from the host:
    the step replays as a CUDA graph while ``kv_pos`` advances on the device.  Returns the layer output [B, H]."""
    attn, mlp = layer.self_attn, layer.mlp
    B, H = hf.shape
    I = mlp.language_mlp.intermediate_size
    dev = hf.device
    counts, ident = _decode_constants(B, dev)
    n_rows = counts[2:3]
    new = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=dev)
    ln1, ln2 = resolve_norm(layer.input_layernorm), resolve_norm(layer.post_attention_layernorm)
    qkv_s = resolve_linear(attn.language_expert_query_key_value)
    dense_s = resolve_linear(attn.language_expert_dense)
    gate_s, up_s = resolve_linear(mlp.language_mlp.gate_proj), resolve_linear(mlp.language_mlp.up_proj)
    down_s = resolve_linear(mlp.language_mlp.down_proj)

    # batches of up to 32 rows take K12, the HBM-bound weight-streaming GEMM (a 128 x 256 tcgen05 tile over 8 live rows
    # leaves most SMs idle on the small-N projections); larger batches and odd LoRA ranks stay on K3
    skinny = B <= 32 and os.environ.get("VEX_DECODE_GEMM", "skinny") != "k3"

    def gemm(a, w, out, mode, spec_pair, residual=None, rope=(), rope_cols=0, kv=(None, None, 0, None)):
        t, r, lb = [None, None], 0, [None] * 4
        sk = skinny and all(sp.r % 64 == 0 for sp in spec_pair)
        for h, sp in enumerate(spec_pair):
            th, rh, bh = _lora_t_single(a, sp, counts, sk)
            if th is not None:
                t[h], r, lb[h] = th, rh, bh
        if len(spec_pair) == 2 and (t[0] is None) != (t[1] is None):
            raise NotImplementedError("gate_proj and up_proj adapters must come in pairs")
        if residual is None and mode == ops.EPI_RESIDUAL:
            residual = out   # in place
        ops.grouped_gemm_fused(a, w, out, counts, mode, None, residual, t, lb, r, list(rope), rope_cols, True, 1.0, *kv,
                               skinny=sk)

    xn = new(B, H)
    ops.rmsnorm_gather(hf, ln1.weight.detach(), ln1.variance_epsilon, None, n_rows, xn)
    max_pos = max(k_cache.shape[2], attn.max_position_embeddings)
    cos, sin = attn.rotary_emb.tables(max_pos, dev, torch.bfloat16)
    pos_flat = position_ids.reshape(-1)
    _debug_check_positions(pos_flat, cos.shape[0])
    qkv = new(B, 3 * H)
    gemm(xn, [_base_weight(qkv_s.weight)], qkv, ops.EPI_ROPE, [qkv_s], rope=(cos, sin, pos_flat, ident),
         rope_cols=2 * H, kv=(k_cache, v_cache, 1, kv_pos))
    ctx = new(B, H)
    ops.attention_decode_cache(qkv[:, :H], k_cache, v_cache, mask, kv_pos, ctx, HEAD_DIM ** -0.5)
    h1 = new(B, H)
    gemm(ctx, [_base_weight(dense_s.weight)], h1, ops.EPI_RESIDUAL, [dense_s], residual=hf)
    ops.rmsnorm_gather(h1, ln2.weight.detach(), ln2.variance_epsilon, None, n_rows, xn)
    act = new(B, I)
    gemm(xn, [_base_weight(gate_s.weight), _base_weight(up_s.weight)], act, ops.EPI_SWIGLU, [gate_s, up_s])
    gemm(act, [_base_weight(down_s.weight)], h1, ops.EPI_RESIDUAL, [down_s])  # in place: h1 += down(act)
    return h1


KV_HEADROOM = int(os.environ.get("VEX_KV_HEADROOM", "256"))  # positions pre-allocated past the prefill length
_kv_pos_cache: dict = {}


def _kv_pos_tensor(past_len: int, device) -> torch.Tensor:
    """Device scalar holding ``past_len`` for the tuple-cache API (shared by the 32 layers of one step)."""
    key = (device, past_len)
    t = _kv_pos_cache.get(key)
    if t is None:
        if len(_kv_pos_cache) > 8:
            _kv_pos_cache.clear()
        t = _kv_pos_cache[key] = torch.tensor([past_len], dtype=torch.int32, device=device)
    return t


def _cache_capacity(t: torch.Tensor) -> int:
    """Capacity (positions) of the pre-allocated buffer a [B, heads, L, 128] cache view lives in, 0 if it is not such
    a view (e.g. the contiguous result of a ``torch.cat`` / beam-search ``index_select``)."""
    B, heads, L, d = t.shape
    if d != HEAD_DIM or t.stride(3) != 1 or t.stride(2) != HEAD_DIM:
        return 0
    cap = t.stride(1) // HEAD_DIM
    if t.stride(1) != cap * HEAD_DIM or (B > 1 and t.stride(0) != heads * cap * HEAD_DIM) or cap < L:
        return 0
    room = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    if room < B * heads * cap * HEAD_DIM:
        return 0
    return cap


def _full_cache_view(t: torch.Tensor, cap: int) -> torch.Tensor:
    B, heads = t.shape[:2]
    return torch.as_strided(t, (B, heads, cap, HEAD_DIM), (heads * cap * HEAD_DIM, cap * HEAD_DIM, HEAD_DIM, 1),
                            t.storage_offset())


def visual_expert_layer_decode(layer: "CogVLMDecoderLayer", hidden_states: torch.Tensor, position_ids: torch.Tensor,
                               padding_mask: torch.Tensor, past_key_value, use_cache: bool = True):
    """One generation step through the reference's tuple-cache interface (q_len == 1 with ``past_key_value`` =
    (k, v) [B, heads, L_past, 128]; returns the layer output and (k, v) [B, heads, L_past + 1, 128]).  The cache the
    prefill returned is a view of a buffer with ``KV_HEADROOM`` spare positions, so the step appends IN PLACE and hands
    back longer views of the same storage; a cache that has no room left (or came from somewhere else, e.g. a
    beam-search reorder) is moved once into a fresh buffer with headroom -- the reference reallocates on every step
    (``torch.cat``, :258-260)."""
    attn = layer.self_attn
    B, L, H = hidden_states.shape
    heads = attn.num_heads
    past_k, past_v = past_key_value
    if past_k.shape[0] != B or past_k.shape[1] != heads or past_k.shape[3] != HEAD_DIM or past_v.shape != past_k.shape:
        raise ValueError(f"past_key_value must be [B, {heads}, L_past, {HEAD_DIM}]")
    Lp = past_k.shape[2]
    Lkv = Lp + 1
    if padding_mask.shape != (B, Lkv):
        raise ValueError(f"padding_mask must cover past + current positions: expected {(B, Lkv)}")
    cap = min(_cache_capacity(past_k), _cache_capacity(past_v))
    if cap >= Lkv:
        k_cache, v_cache = _full_cache_view(past_k, cap), _full_cache_view(past_v, cap)
    else:
        cap = Lkv + KV_HEADROOM
        kv = torch.empty(2, B, heads, cap, HEAD_DIM, dtype=past_k.dtype, device=past_k.device)
        kv[0][:, :, :Lp].copy_(past_k)
        kv[1][:, :, :Lp].copy_(past_v)
        k_cache, v_cache = kv[0], kv[1]
    out = decode_core(layer, hidden_states.view(B, H), position_ids, k_cache, v_cache, padding_mask,
                      _kv_pos_tensor(Lp, hidden_states.device))
    present = (k_cache[:, :, :Lkv], v_cache[:, :, :Lkv]) if use_cache else None
    return out.view(B, 1, H), present


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
        # Training path: autograd is on AND something differentiable is involved -- the input, a LoRA adapter or a
        # modules_to_save norm copy.  A plain (un-wrapped) layer called without torch.no_grad() is inference, as with
        # the reference: its base weights are constants of the fused forward (they are frozen under PEFT anyway).
        train = torch.is_grad_enabled() and (hidden_states.requires_grad or any(
            p.requires_grad and ("lora_" in n or "modules_to_save" in n) for n, p in self.named_parameters()))
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

    Prefill (no grad) runs ``decoder_stack_forward``: the routing plan (K1) is computed once and the residual stream
    stays in expert-sorted order across all layers; with ``graph=True`` the whole prefill is captured into one CUDA
    graph per input shape (``GraphedPrefill``).  Generation: ``prefill_static`` + ``decode_step`` keep the KV cache
    in pre-allocated buffers and replay each step as a CUDA graph (``kv_cache.StaticKVCache``)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([CogVLMDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._graphs = {}

    def _differentiable(self, inputs_embeds: torch.Tensor) -> bool:
        return torch.is_grad_enabled() and (inputs_embeds.requires_grad or any(
            p.requires_grad and ("lora_" in n or "modules_to_save" in n) for n, p in self.named_parameters()))

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
        if past_key_values is None and L > 1 and not self._differentiable(inputs_embeds):
            if not inputs_embeds.is_cuda or inputs_embeds.dtype != torch.bfloat16:
                raise TypeError("inputs_embeds must be a CUDA bfloat16 tensor (no CPU path)")
            plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
            return decoder_stack_forward(self.layers, self.norm, inputs_embeds.contiguous(), plan, position_ids,
                                         use_cache=use_cache)
        h = inputs_embeds
        cache = () if use_cache else None
        for i, layer in enumerate(self.layers):  # :547-569 (training, decode through the tuple-cache interface)
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

    # ---- generation with a static cache ----
    def prefill_static(self, inputs_embeds: torch.Tensor, token_type_ids: torch.Tensor,
                       attention_mask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                       max_new_tokens: int = 256):
        """Prefill that leaves K / V in a ``StaticKVCache`` with room for ``max_new_tokens`` more positions.
        Returns (last_hidden_state [B, L, H], cache)."""
        from .kv_cache import StaticKVCache
        B, L, _ = inputs_embeds.shape
        dev = inputs_embeds.device
        if position_ids is None:
            position_ids = torch.arange(L, dtype=torch.long, device=dev).unsqueeze(0).expand(B, L)
        position_ids = position_ids.reshape(-1, L).long().contiguous()
        if attention_mask is None:
            attention_mask = torch.ones(B, L, dtype=torch.bool, device=dev)
        padding_mask = attention_mask.bool().contiguous()
        cache = StaticKVCache(len(self.layers), B, self.layers[0].self_attn.num_heads, L + max_new_tokens, dev)
        plan = GLOBAL_PLAN_CACHE.get(token_type_ids, padding_mask)
        with torch.no_grad():
            h, _ = decoder_stack_forward(self.layers, self.norm, inputs_embeds.contiguous(), plan, position_ids,
                                         use_cache=True, kv_out=cache.layers)
        cache.start(padding_mask)
        return h, cache

    def decode_step(self, inputs_embeds: torch.Tensor, position_ids: torch.Tensor, cache, graph: bool = True):
        """One generation step for every sample: ``inputs_embeds`` [B, 1, H], ``position_ids`` [B, 1]; appends to
        ``cache`` in place and returns the final-normed hidden state [B, 1, H].  With ``graph`` the 32-layer step is one
        CUDA-graph replay (captured on first use per cache)."""
        return cache.step(self, inputs_embeds, position_ids, graph=graph)
