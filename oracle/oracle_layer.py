"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference visual-expert decoder layer.

This is the oracle the CUDA path is checked against.  It restates, in plain eager PyTorch over a
flat ``weights`` dict (state-dict keys of the reference layer), the algorithm of
``/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py:30-340`` (function-by-function citations
below).  It is *not* the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Pinning status: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through
``oracle/reference_loader.py`` (unmodified source, stubbed imports): ``oracle/make_golden.py`` writes
those outputs to ``tests/golden/`` and ``tests/test_oracle.py`` checks (a) oracle == fixtures
everywhere and (b) oracle == live reference, bit-exact in fp32, wherever ``/root/reference`` exists.
Two pieces of arithmetic live in third-party code absent from ``/root/reference`` and are restated
from their documented behaviour -- parity for them is "unpinned" by any reference-side test:
  * xformers 0.0.27 ``memory_efficient_attention`` + ``BlockDiagonalCausalMask`` (environment.yaml:41,
    call site modeling_cogvlm.py:113-128);
  * peft ``lora.Linear.forward`` / ``ModulesToSaveWrapper`` (environment.yaml:20, wired at
    scripts/cli.py:82-88, hyper-parameters conf/lora.yaml:1-4).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

LANGUAGE_TOKEN_TYPE = 0  # mmmm/data/utils.py:192
VISION_TOKEN_TYPE = 1    # mmmm/data/utils.py:193

EXPERTS = ("vision", "language")


# --------------------------------------------------------------------------------------------- a1
def expert_masks(token_type_ids: torch.Tensor, padding_mask: torch.Tensor):
    """``get_expert_mask`` (modeling_cogvlm.py:58-70).

    vision[b, l] = tt[b, l] == 1 and tt[b, l + 1] == 1 for l < L - 1, False in the last column;
    language = not vision; when L > 1 both are AND-ed with ``padding_mask`` (padded tokens belong to
    neither expert).  For L == 1 the padding mask is ignored.
    """
    B, L = token_type_ids.shape
    is_vis = token_type_ids == VISION_TOKEN_TYPE
    vision = torch.zeros(B, L, dtype=torch.bool, device=token_type_ids.device)
    if L > 1:
        vision[:, : L - 1] = is_vis[:, : L - 1] & is_vis[:, 1:]
    language = ~vision
    if L > 1:
        vision = vision & padding_mask
        language = language & padding_mask
    return vision, language


@dataclass
class RoutingPlan:
    """What the partition/compaction kernel (K1) must reproduce bit-exactly.

    ``x[mask]`` in the reference enumerates True positions in ascending flat (b * L + l) order
    (ATen ``nonzero``), so every list below is ascending.
    """
    vision_idx: torch.Tensor    # int64 [Tv]  flat positions routed to the vision expert
    language_idx: torch.Tensor  # int64 [Tl]  flat positions routed to the language expert
    valid_idx: torch.Tensor     # int64 [T]   flat positions with padding_mask == True
    cu_seqlens: torch.Tensor    # int64 [B+1] prefix sum of valid tokens per sample


def routing_plan(token_type_ids: torch.Tensor, padding_mask: torch.Tensor) -> RoutingPlan:
    vision, language = expert_masks(token_type_ids, padding_mask)
    flat = lambda m: m.reshape(-1).nonzero(as_tuple=True)[0]
    lens = padding_mask.sum(dim=1).to(torch.int64)
    cu = torch.zeros(padding_mask.shape[0] + 1, dtype=torch.int64)
    cu[1:] = torch.cumsum(lens, 0)
    return RoutingPlan(flat(vision), flat(language), flat(padding_mask), cu)


# --------------------------------------------------------------------------------------------- a2
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """``RMSNorm.forward`` (modeling_cogvlm.py:36-41): fp32 statistics, the multiply by ``weight``
    happens in fp32 BEFORE the cast back to the input dtype."""
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return (weight * xf).to(x.dtype)


def masked_rms_norm(h: torch.Tensor, padding_mask: torch.Tensor, weight: torch.Tensor, eps: float):
    """``_mask_set(h, pm, norm(h[pm]))`` (modeling_cogvlm.py:306-309, 390-393): rows with
    ``padding_mask == False`` pass through un-normalised; for L == 1 every row is normalised."""
    if h.shape[1] > 1:
        out = h.clone()
        out[padding_mask] = rms_norm(h[padding_mask], weight, eps)
        return out
    return rms_norm(h, weight, eps)


# --------------------------------------------------------------------------------------------- a10
@dataclass
class LoRA:
    """One PEFT ``lora.Linear`` adapter: y = base(x) + B(A(x.to(A.dtype))) * scaling, cast back to the
    base output dtype.  rsLoRA scaling = alpha / sqrt(r) = 8 / 8 = 1.0 (conf/lora.yaml:1-4)."""
    A: torch.Tensor  # [r, in]
    B: torch.Tensor  # [out, r]
    scaling: float = 1.0
    # training mode: PEFT computes lora_B(lora_A(lora_dropout(x))) (conf/lora.yaml: lora_dropout 0.05).  The mask is an
    # explicit input here -- entries 0 or 1 / (1 - p), same shape as the rows this Linear sees -- so a test can hand the
    # oracle the very mask the CUDA path generated (PyTorch's own Philox stream is not reproducible across devices).
    drop_mask: Optional[torch.Tensor] = None


def linear(x: torch.Tensor, w: torch.Tensor, lora: Optional[LoRA] = None) -> torch.Tensor:
    y = F.linear(x, w)
    if lora is not None:
        xl = x.to(lora.A.dtype)
        if lora.drop_mask is not None:
            xl = (xl * lora.drop_mask.to(xl.dtype)).to(xl.dtype)
        delta = F.linear(F.linear(xl, lora.A), lora.B) * lora.scaling
        y = (y + delta).to(y.dtype)
    return y


# --------------------------------------------------------------------------------------------- a4
def default_inv_freq(head_dim: int, base: float = 10000.0) -> torch.Tensor:
    """``RotaryEmbedding._compute_inv_freq`` (modeling_cogvlm.py:156-160)."""
    return 1.0 / (base ** (torch.arange(0, head_dim, 2) / head_dim))


def rotary_tables(inv_freq: torch.Tensor, seq_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``RotaryEmbedding._set_cos_sin_cache`` (modeling_cogvlm.py:162-170).

    The table is built IN ``inv_freq.dtype``: under bf16-true the positions themselves are a bf16
    ``arange`` (positions >= 257 collapse) -- SURVEY.md section 0 quirk 2.  Returns cos, sin of
    shape [seq_len, head_dim] with emb = cat(freqs, freqs).
    """
    t = torch.arange(seq_len, device=inv_freq.device, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """modeling_cogvlm.py:183-185 (NeoX half-split pairing (j, j + d/2))."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def apply_rotary(q, k, cos, sin, position_ids):
    """``apply_rotary_pos_emb_index_bhs`` (modeling_cogvlm.py:188-193); q, k are [B, heads, L, d],
    cos/sin [S, d] are gathered by explicit ``position_ids`` [B, L]."""
    c = F.embedding(position_ids, cos).unsqueeze(1)
    s = F.embedding(position_ids, sin).unsqueeze(1)
    return (q * c) + (rotate_half(q) * s), (k * c) + (rotate_half(k) * s)


# --------------------------------------------------------------------------------------------- a5
def attention(q, k, v, padding_mask):
    """Prefill branch of ``attention_fn`` (modeling_cogvlm.py:106-128,142) with xformers'
    ``BlockDiagonalCausalMask`` semantics restated: per sample the valid tokens are compacted and
    token i attends to compacted tokens j <= i; scale d ** -0.5; fp32 softmax; P rounded to the value
    dtype before P @ V; padded rows are zero.  q, k, v, out: [B, heads, L, d]."""
    B, H, L, D = q.shape
    out = torch.zeros_like(q)
    scale = D ** -0.5
    for b in range(B):
        idx = padding_mask[b].nonzero(as_tuple=True)[0]
        n = idx.numel()
        if n == 0:
            continue
        qb, kb, vb = q[b, :, idx].float(), k[b, :, idx].float(), v[b, :, idx]
        s = torch.matmul(qb, kb.transpose(-1, -2)) * scale
        keep = torch.ones(n, n, dtype=torch.bool, device=q.device).tril()
        p = torch.softmax(s.masked_fill(~keep, float("-inf")), dim=-1)
        out[b, :, idx] = torch.matmul(p.to(vb.dtype).float(), vb.float()).to(q.dtype)
    return out


def attention_decode(q, k, v, padding_mask):
    """Generation branch of ``attention_fn`` (modeling_cogvlm.py:129-141): q_len == 1 against the cached keys.
    q [B, heads, 1, d]; k, v [B, heads, Lkv, d]; padding_mask [B, Lkv].  The query is scaled IN its dtype
    (``query_layer *= d ** -0.5``), masked keys/values are zeroed, scores are masked to -inf, softmax runs in fp32
    and is cast back to the score dtype before the weighted sum."""
    assert q.shape[2] == 1
    qs = q[:, :, 0] * (q.shape[-1] ** -0.5)                                   # [B, H, d]       :132
    kk = k.permute(0, 2, 1, 3).clone()                                         # [B, Lkv, H, d]
    vv = v.permute(0, 2, 1, 3).clone()
    kk[~padding_mask] = 0                                                      # :134-135
    vv[~padding_mask] = 0
    scores = torch.einsum("nhd,nlhd->nlh", qs, kk)                             # :136
    scores[~padding_mask] = -torch.inf                                         # :137
    scores = scores.softmax(dim=1, dtype=torch.float32).to(dtype=scores.dtype)  # :138
    out = torch.einsum("nlhd,nlh->nhd", vv, scores)[:, None]                   # [B, 1, H, d]   :141
    return out.permute(0, 2, 1, 3)                                             # [B, H, 1, d]   :142


# --------------------------------------------------------------------------------------------- a3/a6/a7/a8/a9
def decoder_layer(
    weights: Dict[str, torch.Tensor],
    hidden_states: torch.Tensor,
    token_type_ids: torch.Tensor,
    position_ids: torch.Tensor,
    padding_mask: torch.Tensor,
    *,
    num_heads: int,
    rms_norm_eps: float = 1e-6,
    lora: Optional[Dict[str, LoRA]] = None,
    use_cache: bool = False,
    cos_sin: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    past_key_value: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
):
    """``CogVLMDecoderLayer.forward`` (modeling_cogvlm.py:295-340): prefill (q_len > 1, no past KV) and the
    decode step (q_len == 1 with ``past_key_value``; ``padding_mask`` then covers past + current positions).

    ``weights`` uses the reference state-dict keys; ``lora`` maps a Linear's
    module path (e.g. ``"self_attn.vision_expert_query_key_value"``) to its adapter.  Rows with
    ``padding_mask == False`` of the result are unspecified in the reference
    (``torch.empty`` at modeling_cogvlm.py:277) -- here they are ``residual + 0``.
    Returns ``(hidden_states,)`` or ``(hidden_states, (k, v))`` with post-rotary k (:262).
    """
    lora = lora or {}
    B, L, Hd = hidden_states.shape
    assert (L > 1) == (past_key_value is None), "supported: prefill without cache, or q_len == 1 with cache"
    d = Hd // num_heads
    vmask, lmask = expert_masks(token_type_ids, padding_mask)
    masks = (vmask, lmask)

    def lin(x, path):
        return linear(x, weights[path + ".weight"], lora.get(path))

    # --- attention block (:305-321) ---
    residual = hidden_states
    if L > 1:
        h = masked_rms_norm(hidden_states, padding_mask, weights["input_layernorm.weight"], rms_norm_eps)
    else:                                                                     # :308-309
        h = rms_norm(hidden_states, weights["input_layernorm.weight"], rms_norm_eps)
    mixed = torch.zeros(B, L, 3 * Hd, dtype=h.dtype, device=h.device)        # :243
    for expert, m in zip(EXPERTS, masks):                                     # :244-245
        mixed[m] = lin(h[m], f"self_attn.{expert}_expert_query_key_value")
    q, k, v = torch.split(mixed, Hd, dim=-1)                                  # :247
    to_heads = lambda t: t.view(B, L, num_heads, d).permute(0, 2, 1, 3)       # :222-226
    q, k, v = to_heads(q), to_heads(k), to_heads(v)
    if cos_sin is None:
        seq_len = int(position_ids.max()) + 1                                 # :255
        cos, sin = rotary_tables(weights["self_attn.rotary_emb.inv_freq"], seq_len)
    else:
        cos, sin = cos_sin
    cos, sin = cos.to(v.dtype), sin.to(v.dtype)                               # :177-180
    q, k = apply_rotary(q, k, cos, sin, position_ids)                         # :256
    if past_key_value is not None:                                            # :258-260
        k = torch.cat([past_key_value[0], k], dim=2)
        v = torch.cat([past_key_value[1], v], dim=2)
    present = (k, v) if use_cache else None                                   # :262
    if L > 1:
        ctx = attention(q, k, v, padding_mask)                                # :264, prefill branch
    else:
        ctx = attention_decode(q, k, v, padding_mask)                         # :264, generation branch
    ctx = ctx.transpose(1, 2).contiguous().reshape(B, L, Hd)                  # :275
    attn_out = torch.zeros(B, L, Hd, dtype=h.dtype, device=h.device)          # :277 (empty in the reference)
    for expert, m in zip(EXPERTS, masks):                                     # :278-279
        attn_out[m] = lin(ctx[m], f"self_attn.{expert}_expert_dense")
    h = residual + attn_out                                                   # :321

    # --- MLP block (:324-330) ---
    residual = h
    if L > 1:
        hn = masked_rms_norm(h, padding_mask, weights["post_attention_layernorm.weight"], rms_norm_eps)
    else:                                                                     # :327-328
        hn = rms_norm(h, weights["post_attention_layernorm.weight"], rms_norm_eps)
    mlp_out = torch.zeros_like(hn)                                            # :95
    for expert, m in zip(EXPERTS, masks):                                     # :96-97, MLP.forward :54-56
        x = hn[m]
        p = f"mlp.{expert}_mlp."
        act = F.silu(lin(x, p + "gate_proj")) * lin(x, p + "up_proj")
        mlp_out[m] = lin(act, p + "down_proj")
    h = residual + mlp_out                                                    # :330
    return (h, present) if use_cache else (h,)


def decoder_stack(layers_weights, hidden_states, token_type_ids, position_ids, padding_mask, *,
                  num_heads, rms_norm_eps=1e-6, final_norm_weight=None, lora=None):
    """The caller loop ``CogVLMModel.llm_forward`` (modeling_cogvlm.py:547-573), prefill, no cache."""
    h = hidden_states
    for i, w in enumerate(layers_weights):
        (h,) = decoder_layer(w, h, token_type_ids, position_ids, padding_mask, num_heads=num_heads,
                             rms_norm_eps=rms_norm_eps, lora=None if lora is None else lora[i])
    if final_norm_weight is not None:
        h = masked_rms_norm(h, padding_mask, final_norm_weight, rms_norm_eps)
    return h


# --------------------------------------------------------------------------------------------- synthetic weights
LINEAR_SHAPES = lambda H, I: {
    "self_attn.vision_expert_query_key_value": (3 * H, H),
    "self_attn.vision_expert_dense": (H, H),
    "self_attn.language_expert_query_key_value": (3 * H, H),
    "self_attn.language_expert_dense": (H, H),
    "mlp.language_mlp.gate_proj": (I, H),
    "mlp.language_mlp.up_proj": (I, H),
    "mlp.language_mlp.down_proj": (H, I),
    "mlp.vision_mlp.gate_proj": (I, H),
    "mlp.vision_mlp.up_proj": (I, H),
    "mlp.vision_mlp.down_proj": (H, I),
}


def random_weights(hidden_size: int, intermediate_size: int, num_heads: int, *, seed: int = 0,
                   dtype=torch.float32, std: float = 0.02, norm_jitter: float = 0.1):
    """Random-init weights with the reference state-dict keys and shapes;
    Linear ~ N(0, initializer_range = 0.02) (modeling_cogvlm.py:350-355, configuration_cogvlm.py:17)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for path, shape in LINEAR_SHAPES(hidden_size, intermediate_size).items():
        w[path + ".weight"] = (torch.randn(shape, generator=g) * std).to(dtype)
    for n in ("input_layernorm", "post_attention_layernorm"):
        w[n + ".weight"] = (1.0 + norm_jitter * torch.randn(hidden_size, generator=g)).to(dtype)
    w["self_attn.rotary_emb.inv_freq"] = default_inv_freq(hidden_size // num_heads).to(dtype)
    return w


def random_lora(hidden_size: int, intermediate_size: int, *, r: int = 64, seed: int = 1,
                dtype=torch.float32, scaling: float = 1.0, b_std: float = 0.02):
    """LoRA adapters on all 10 Linears: A ~ kaiming-uniform(a = sqrt(5)) like PEFT's init, B ~ N(0, b_std)
    (PEFT inits B to zero; non-zero here so that the delta is visible -- SURVEY.md section 8(d))."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for path, (o, i) in LINEAR_SHAPES(hidden_size, intermediate_size).items():
        bound = 1.0 / math.sqrt(i)  # kaiming_uniform_(a=sqrt(5)) on [r, in] -> U(-1/sqrt(in), 1/sqrt(in))
        A = (torch.rand(r, i, generator=g) * 2 - 1) * bound
        Bm = torch.randn(o, r, generator=g) * b_std
        out[path] = LoRA(A.to(dtype), Bm.to(dtype), scaling)
    return out


# --------------------------------------------------------------------------------------------- section 8(f)-3
CE_IGNORE_INDEX = -100  # mmmm/data/defs.py


def sample_weighted_ce(logits: torch.Tensor, labels: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    """``_sample_weighted_ce`` (modeling_cogvlm.py:610-627): plain mean cross-entropy over labels != -100 when ``weight``
    is None, else dot(ce[mask], weight.float()[mask]) / mask.sum() with a per-position weight."""
    logits = logits.view(-1, logits.shape[-1])
    labels = labels.view(-1)
    if weight is None:
        return F.cross_entropy(logits, labels)
    mask = labels != CE_IGNORE_INDEX
    ce = F.cross_entropy(logits, labels, reduction="none")
    return torch.dot(ce[mask], weight.float().view(-1)[mask]) / mask.sum()


def lm_head_loss(hidden_states: torch.Tensor, lm_head_weight: torch.Tensor, labels: torch.Tensor,
                 weight: Optional[torch.Tensor] = None, lora: Optional[LoRA] = None) -> torch.Tensor:
    """``logits = self.lm_head(last_hidden_state).float(); loss = _sample_weighted_ce(logits, labels, weight)``
    (CogVLMForCausalLM.forward, modeling_cogvlm.py:701-706; labels are already shifted by the data module)."""
    logits = linear(hidden_states, lm_head_weight, lora).float()
    return sample_weighted_ce(logits, labels, weight)
