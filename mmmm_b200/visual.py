"""Drop-in EVA2-CLIP vision encoder backed by the libvex sm_100a kernels.

Mirrors the module surface of the reference (``/root/reference/mmmm/models/cogvlm/visual.py:24-208``): the same class
names, child-module tree and state-dict keys

    patch_embedding.proj.{weight (C, 3, pd, ph, pw), bias}           Downsample (nn.Conv3d, kernel == stride)
    patch_embedding.{cls_embedding, cls_pos_embed, position_embedding}.weight      ParameterWrapper children
    transformer.layers.N.{input_layernorm, post_attention_layernorm}.{weight, bias}
    transformer.layers.N.attention.{query_key_value, dense}.{weight, bias}
    transformer.layers.N.mlp.{fc1, fc2}.{weight, bias}
    linear_proj.{linear_proj, gate_proj, dense_h_to_4h, dense_4h_to_h}.weight, linear_proj.norm1.{weight, bias}
    boi, eoi

and the same call ``EVA2CLIPModel.forward(image: list[[C, D, H, W]], patch_size: list[(pd, ph, pw)],
pool_size_list) -> list[[1, n_i + 2, hidden]]`` (visual.py:191-208).  What differs is the forward: instead of
calling its children it reads their tensors and runs, over all images packed into one token sequence,

    K11 patchify (im2col) -> K3 patch GEMM (+bias, + position embedding, scatter behind the class tokens)
    63 x [ K3 QKV (+bias) -> K4 non-causal block-diagonal attention -> K3 dense (+bias) -> K11 LayerNorm + residual
           -> K3 fc1 (+bias, GELU) -> K3 fc2 (+bias) -> K11 LayerNorm + residual ]
    K11 class-token drop / 3-D max-pool -> K3 linear_proj -> K11 LayerNorm + GELU -> K3 gate/up (SwiGLU epilogue)
    -> K3 dense_4h_to_h with the scatter to the output rows fused -> K11 boi / eoi rows

with no host synchronisation (every count is known on the host from the image shapes).  ``encode_into`` additionally
fuses ``CogVLMModel.forward``'s feature scatter (modeling_cogvlm.py:450-453): the last GEMM writes straight into the
rows of ``inputs_embeds``.

Head dimension: EVA2-CLIP-E has 16 heads of 112.  The attention kernel works on head slots of 128, so the QKV /
dense weights are re-laid once per weight version into zero-padded slots (q.k and p.v are unchanged by zero
columns; scale stays 112^-0.5).  A 128-wide head needs no copy.

bf16 CUDA tensors only, inference only (autograd through the vision encoder, LoRA adapters that are active and
unmerged on its Linears, and the HF-checkpoint position-embedding inflation of ``_load_from_state_dict``
visual.py:38-57 are not implemented and raise / are left to the caller); there is no CPU fallback.
"""
from __future__ import annotations

from argparse import Namespace
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .modeling_cogvlm import NoWeightDecayParameter

SLOT = 128  # head slot width of the attention kernel


class ParameterWrapper(nn.Module):
    """``mmmm.utils.ParameterWrapper`` (mmmm/utils.py:62-80): a parameter exposed as ``<name>.weight``."""

    def __init__(self, weight: nn.Parameter):
        super().__init__()
        self.weight = weight

    def extra_repr(self) -> str:
        return f"shape={tuple(self.weight.shape)}"


class Downsample(nn.Conv3d):
    """``mmmm.models.resample.Downsample`` (resample.py:15-63): Conv3d with kernel == stride whose depth can be
    reduced at call time.  Only the parameters are used here (the convolution runs as patchify + GEMM)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, bias: bool = True):
        super().__init__(in_channels, out_channels, kernel_size, kernel_size, bias=bias)


class PatchEmbedding(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.proj = Downsample(config.in_channels, config.hidden_size, config.patch_size)
        self.pos_embed_shape = tuple(config.pos_embed_shape)
        pt = getattr(config, "pt_pos_embed_shape", None)   # grid of the pretrained 2-D position embedding (visual.py:31)
        self.pt_pos_embed_shape = None if pt is None else tuple(pt)
        self.cls_embedding = ParameterWrapper(NoWeightDecayParameter(torch.zeros(1, config.hidden_size)))
        self.cls_pos_embed = ParameterWrapper(NoWeightDecayParameter(torch.zeros(1, config.hidden_size)))
        self.position_embedding = ParameterWrapper(
            NoWeightDecayParameter(torch.zeros(1, config.hidden_size, *config.pos_embed_shape)))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """Same checkpoint inflation as the reference (visual.py:38-57): a pretrained 2-D EVA2-CLIP position embedding
        ``[1 + h*w, C]`` (class row first) is split into ``cls_pos_embed`` and a ``[1, C, d, h, w]`` grid -- resampled
        when the pretrained grid differs from ``pos_embed_shape[-2:]``, repeated along depth -- and a saved
        modules_to_save copy of another grid size is resampled to this module's grid."""
        key = f"{prefix}position_embedding.weight"
        saved = f"{prefix}position_embedding.modules_to_save.default.weight"
        if (pe := state_dict.get(key)) is not None and pe.ndim == 2:
            cls_pe, pe = pe[0:1], pe[1:]
            h, w = self.pt_pos_embed_shape if self.pt_pos_embed_shape is not None else self.pos_embed_shape[-2:]
            if pe.shape[0] != h * w:
                raise ValueError(f"{key}: {pe.shape[0]} rows do not form the pretrained {h} x {w} grid "
                                 f"(config.pt_pos_embed_shape)")
            pe = pe.reshape(h, w, -1).permute(2, 0, 1)[None]                        # '(h w) c -> 1 c h w'
            if (h, w) != tuple(self.pos_embed_shape[-2:]):
                pe = resample(pe.float(), self.pos_embed_shape[-2:]).to(pe.dtype)
            pe = pe[:, :, None].expand(-1, -1, self.pos_embed_shape[0], -1, -1).contiguous()   # '1 c h w -> 1 c d h w'
            del state_dict[key]
            state_dict[f"{prefix}cls_pos_embed"] = cls_pe
            state_dict[f"{prefix}position_embedding"] = pe
        elif (pe := state_dict.get(saved)) is not None and tuple(pe.shape[2:]) != tuple(self.position_embedding.weight.shape[2:]):
            state_dict[saved] = resample(pe.float(), self.position_embedding.weight.shape[2:]).to(pe.dtype)
        # ParameterWrapper.wrap (mmmm/utils.py:71-77): accept the bare-parameter spelling of the three embeddings
        for name in ("cls_embedding", "cls_pos_embed", "position_embedding"):
            if (w := state_dict.pop(prefix + name, None)) is not None:
                state_dict[f"{prefix}{name}.weight"] = w
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_heads = config.num_heads
        self.scale = (config.hidden_size // config.num_heads) ** -0.5
        self.query_key_value = nn.Linear(config.hidden_size, config.hidden_size * 3)
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)


class MLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError("the vision MLP epilogue implements ACT2FN['gelu'] only")
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class TransformerLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.input_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.attention = Attention(config)
        self.mlp = MLP(config)
        self.post_attention_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class Transformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gradient_checkpointing = False
        self._gradient_checkpointing_func = None
        self.layers = nn.ModuleList([TransformerLayer(config) for _ in range(config.num_hidden_layers)])


class GLU(nn.Module):
    def __init__(self, config, in_features):
        super().__init__()
        self.linear_proj = nn.Linear(in_features, config.hidden_size, bias=False)
        self.norm1 = nn.LayerNorm(config.hidden_size)
        self.dense_h_to_4h = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.dense_4h_to_h = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)


# ------------------------------------------------------------------------------------------ weight access
def _unwrap_linear(mod: nn.Module) -> nn.Linear:
    """``nn.Linear`` or a PEFT-style ``lora.Linear`` around one.  The vision Linears are LoRA targets in the
    reference (mmmm/utils.py:19-43); adapters must be merged or disabled for this forward."""
    if hasattr(mod, "base_layer") and hasattr(mod, "lora_A"):
        names = getattr(mod, "active_adapters", None) or []
        names = [names] if isinstance(names, str) else list(names)
        live = [n for n in names if n in mod.lora_A]
        if live and not (getattr(mod, "disable_adapters", False) or getattr(mod, "merged", False)):
            raise NotImplementedError("active, unmerged LoRA adapters on the vision encoder are not implemented: "
                                      "merge_adapter() or disable them")
        base = mod.base_layer
        while hasattr(base, "base_layer"):
            base = base.base_layer
        return base
    return mod


def _unwrap_saved(mod: nn.Module) -> nn.Module:
    """PEFT ``ModulesToSaveWrapper`` -> the active copy (LayerNorms, ParameterWrappers and Conv3d are
    modules_to_save in the reference, mmmm/utils.py:35-37)."""
    if hasattr(mod, "modules_to_save") and hasattr(mod, "original_module"):
        names = getattr(mod, "active_adapters", None) or getattr(mod, "active_adapter", [])
        names = [names] if isinstance(names, str) else list(names)
        if not getattr(mod, "disable_adapters", False):
            for n in names:
                if n in mod.modules_to_save:
                    return mod.modules_to_save[n]
        return mod.original_module
    return mod


def _check(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{what} must live on a CUDA device (the vision encoder has no CPU path)")
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{what} must be bfloat16, got {t.dtype}")
    return t.detach().contiguous()


class _Derived:
    """Re-laid weight copies keyed on (data_ptr, _version) of their sources."""

    def __init__(self):
        self._cache: Dict[tuple, tuple] = {}

    def get(self, key, sources: Sequence[torch.Tensor], make):
        sig = tuple((s.data_ptr(), s._version, tuple(s.shape)) for s in sources)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = make()
        self._cache[key] = (sig, val)
        return val


def _pad_heads_rows(w: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """[3 * heads * hd, ...] -> [3 * heads * 128, ...] with each head's rows in a zero-padded 128-slot."""
    rest = w.shape[1:]
    out = w.new_zeros(3, heads, SLOT, *rest)
    out[:, :, :hd] = w.reshape(3, heads, hd, *rest)
    return out.reshape(3 * heads * SLOT, *rest).contiguous()


def _pad_heads_cols(w: torch.Tensor, heads: int, hd: int) -> torch.Tensor:
    """[out, heads * hd] -> [out, heads * 128] with zero columns in the padded part of every head slot."""
    out = w.new_zeros(w.shape[0], heads, SLOT)
    out[:, :, :hd] = w.reshape(w.shape[0], heads, hd)
    return out.reshape(w.shape[0], heads * SLOT).contiguous()


def resample(x: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """``luolib.models.spadop.resample`` (luolib/models/spadop/resample.py:11-29) for the position embedding: area
    interpolation down, trilinear up, identity when the grid already matches.  Parameter preprocessing (cached per
    grid shape), not token work."""
    shape = tuple(int(s) for s in shape)
    down = tuple(np.minimum(x.shape[2:], shape).tolist())
    if down != tuple(x.shape[2:]):
        x = F.interpolate(x, down, mode="area")
    if shape != tuple(x.shape[2:]):
        x = F.interpolate(x, shape, mode="trilinear" if x.ndim == 5 else "bicubic")
    return x


# ------------------------------------------------------------------------------------------ host-side plan
@dataclass
class _ImagePlan:
    grid: Tuple[int, int, int]     # patch grid (d, h, w)
    n: int                         # patches
    start: int                     # first packed row (the class token)
    pool: Tuple[int, int, int]
    ogrid: Tuple[int, int, int]    # grid after pooling
    m: int                         # feature rows after pooling
    fstart: int                    # first compact feature row


class VisionPlan:
    """Everything derived from the image / patch / pool shapes: packed-sequence offsets, per-patch-size row groups,
    row maps and counts (small int32 device tensors, built once per distinct batch geometry)."""

    def __init__(self, shapes, patch_sizes, pool_sizes, device):
        self.images: List[_ImagePlan] = []
        t = f = 0
        for (C, D, H, W), ps, pool in zip(shapes, patch_sizes, pool_sizes):
            g = (D // ps[0], H // ps[1], W // ps[2])
            if min(g) <= 0:
                raise ValueError(f"image {(D, H, W)} is smaller than its patch size {ps}")
            pool = tuple(int(p) for p in pool)
            og = tuple(a // b for a, b in zip(g, pool)) if any(p > 1 for p in pool) else g
            if min(og) <= 0:
                raise ValueError(f"pool size {pool} exceeds the patch grid {g}")
            n, m = g[0] * g[1] * g[2], og[0] * og[1] * og[2]
            self.images.append(_ImagePlan(g, n, t, pool, og, m, f))
            t += 1 + n
            f += m
        self.B = len(self.images)
        self.T = t                                    # packed rows (class tokens included)
        self.M = f                                    # feature rows
        self.max_len = max(1 + im.n for im in self.images)
        self.rows_cap = self.B * self.max_len         # what the attention kernel's buffers are sized for
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device)
        self.cu_seqlens = i32([im.start for im in self.images] + [t])
        self.n_rows = i32([t, 0, 0, 0])               # single-expert counts of the packed sequence
        self.n_feat = i32([f, 0, 0, 0])
        # patch-convolution groups: images sharing a patch size share one GEMM (one depth-reduced weight)
        self.groups: Dict[Tuple[int, int, int], dict] = {}
        for i, (im, ps) in enumerate(zip(self.images, patch_sizes)):
            self.groups.setdefault(tuple(int(p) for p in ps), dict(images=[], rows=0))["images"].append(i)
        for ps, g in self.groups.items():
            dst, off, offs = [], 0, []
            for i in g["images"]:
                im = self.images[i]
                offs.append(off)
                dst.extend(range(im.start + 1, im.start + 1 + im.n))
                off += im.n
            g.update(rows=off, offsets=offs, row_map=i32(dst), count=i32([off, 0, 0, 0]))
        self.derived = _Derived()                     # tensors that depend on this geometry (+ parameter versions)

    def feature_row_map(self, dest_starts: Sequence[int], device) -> torch.Tensor:
        """compact feature row -> destination row (image i's features start at dest_starts[i])."""
        dst: List[int] = []
        for im, d0 in zip(self.images, dest_starts):
            dst.extend(range(d0, d0 + im.m))
        return torch.tensor(dst, dtype=torch.int32, device=device)


class EVA2CLIPModel(nn.Module):
    """Drop-in for ``mmmm.models.cogvlm.visual.EVA2CLIPModel`` (visual.py:181-208)."""

    def __init__(self, config):
        super().__init__()
        vc = config.vision_config
        vision_config = Namespace(**vc) if isinstance(vc, dict) else vc
        self.vision_config = vision_config
        hd = vision_config.hidden_size // vision_config.num_heads
        if hd > SLOT or hd % 8 != 0:
            raise ValueError(f"head_dim {hd} unsupported (must be a multiple of 8, at most {SLOT})")
        if vision_config.hidden_size % 256 != 0 or config.hidden_size % 256 != 0:
            raise ValueError("hidden sizes must be multiples of 256 (LayerNorm kernel)")
        self.patch_embedding = PatchEmbedding(vision_config)
        self.transformer = Transformer(vision_config)
        self.linear_proj = GLU(config, in_features=vision_config.hidden_size)
        self.boi = NoWeightDecayParameter(torch.zeros(1, 1, config.hidden_size))
        self.eoi = NoWeightDecayParameter(torch.zeros(1, 1, config.hidden_size))
        self._derived = _Derived()
        self._plans: Dict[tuple, VisionPlan] = {}

    # ---------------------------------------------------------------------------------- derived tensors
    def _plan(self, shapes, patch_sizes, pool_sizes, device) -> VisionPlan:
        key = (tuple(shapes), tuple(map(tuple, patch_sizes)), tuple(map(tuple, pool_sizes)), str(device))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 64:
                self._plans.clear()
            plan = self._plans[key] = VisionPlan(shapes, patch_sizes, pool_sizes, device)
        return plan

    def _patch_weight(self, ps: Tuple[int, int, int]):
        """weight.reshape(C_out, -1) of the (depth-reduced, resample.py:56-62) patch kernel, K zero-padded to a
        multiple of 64."""
        proj = _unwrap_saved(self.patch_embedding.proj)
        w = _check(proj.weight, "patch_embedding.proj.weight")

        def make():
            kd, kh, kw = w.shape[2:]
            if (kh, kw) != ps[1:]:
                raise NotImplementedError(f"in-plane patch size {ps[1:]} differs from the kernel's {(kh, kw)}")
            wk = w
            if kd != ps[0]:
                if kd % ps[0] != 0:
                    raise NotImplementedError("patch depth must divide the kernel depth (resample.py:59-60)")
                wk = w.reshape(w.shape[0], w.shape[1], ps[0], kd // ps[0], kh, kw).sum(dim=3)
            flat = wk.reshape(w.shape[0], -1)
            kpad = (flat.shape[1] + 63) // 64 * 64
            out = flat.new_zeros(flat.shape[0], kpad)
            out[:, :flat.shape[1]] = flat
            return out

        wp = self._derived.get(("patch_w", ps), [w], make)
        b = proj.bias
        return wp, (None if b is None else _check(b, "patch_embedding.proj.bias"))

    def _pos_rows(self, plan: VisionPlan) -> torch.Tensor:
        """[T, C]: what the patch GEMM adds its output to -- the resampled position embedding in token order behind
        each image's class row (cls_embedding + cls_pos_embed), visual.py:66-74."""
        pe = self.patch_embedding
        pos = _check(_unwrap_saved(pe.position_embedding).weight, "position_embedding")
        cls = _check(_unwrap_saved(pe.cls_embedding).weight, "cls_embedding")
        cpos = _check(_unwrap_saved(pe.cls_pos_embed).weight, "cls_pos_embed")

        def make():
            rows = []
            per_grid: Dict[tuple, torch.Tensor] = {}
            for im in plan.images:
                if im.grid not in per_grid:
                    per_grid[im.grid] = resample(pos, im.grid).flatten(2)[0].t()
                rows += [cls + cpos, per_grid[im.grid]]
            return torch.cat(rows, dim=0).contiguous()

        return plan.derived.get("pos_rows", [pos, cls, cpos], make)

    def _attn_weights(self, idx: int, attn: Attention):
        heads = attn.num_heads
        qkv, dense = _unwrap_linear(attn.query_key_value), _unwrap_linear(attn.dense)
        wq, bq = _check(qkv.weight, "query_key_value.weight"), _check(qkv.bias, "query_key_value.bias")
        wd, bd = _check(dense.weight, "dense.weight"), _check(dense.bias, "dense.bias")
        hd = wd.shape[1] // heads
        if hd == SLOT:
            return wq, bq, wd, bd
        wq = self._derived.get(("wq", idx), [wq], lambda: _pad_heads_rows(wq, heads, hd))
        bq = self._derived.get(("bq", idx), [bq], lambda: _pad_heads_rows(bq, heads, hd))
        wd = self._derived.get(("wd", idx), [wd], lambda: _pad_heads_cols(wd, heads, hd))
        return wq, bq, wd, bd

    # ---------------------------------------------------------------------------------- forward
    def _encode(self, image: List[torch.Tensor], patch_size, pool_size_list, dest: torch.Tensor,
                feat_starts: Sequence[int]) -> VisionPlan:
        """Runs the encoder over the packed images and writes image i's rows [boi, features, eoi] into ``dest``
        ([rows, lm_hidden] bf16) starting at row ``feat_starts[i]``."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("autograd through the fused vision encoder is not implemented; call it under "
                                      "torch.no_grad()")
        if not (len(image) == len(patch_size) == len(pool_size_list)) or not image:
            raise ValueError("image, patch_size and pool_size_list must be non-empty lists of equal length")
        dev = image[0].device
        images = []
        for im in image:
            if im.dim() != 4:
                raise ValueError("every image must be [C, D, H, W]")
            images.append(_check(im, "image"))
        vc = self.vision_config
        C, heads = vc.hidden_size, vc.num_heads
        plan = self._plan([tuple(im.shape) for im in images], patch_size, pool_size_list, dev)
        cap, T = plan.rows_cap, plan.T
        bf = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)

        # ---- patch embedding: x[packed row] = conv(patch) + bias + pos_embed; class rows = cls + cls_pos
        x = bf(cap, C)
        x[:T].copy_(self._pos_rows(plan))
        for ps, g in plan.groups.items():
            wp, bp = self._patch_weight(ps)
            a = torch.zeros(g["rows"], wp.shape[1], dtype=torch.bfloat16, device=dev)  # K padding stays zero
            for i, off in zip(g["images"], g["offsets"]):
                ops.patchify(images[i], ps[0], ps[1], ps[2], a[off:off + plan.images[i].n])
            ops.linear_bias_act(a, wp, bp, x, g["count"], g["row_map"], True, False)  # x[map(r)] += conv + bias

        # ---- transformer (visual.py:128-135 per layer)
        qkv = bf(cap, 3 * heads * SLOT)
        ctx = bf(cap, heads * SLOT)
        br = bf(cap, C)
        mid = bf(cap, vc.intermediate_size)
        scale = (C // heads) ** -0.5
        for idx, layer in enumerate(self.transformer.layers):
            wq, bq, wd, bd = self._attn_weights(idx, layer.attention)
            ops.linear_bias_act(x, wq, bq, qkv, plan.n_rows, None, False, False)
            ops.attention_blockdiag(qkv, plan.cu_seqlens, plan.B, plan.max_len, heads, ctx, scale)
            ops.linear_bias_act(ctx, wd, bd, br, plan.n_rows, None, False, False)
            ln = _unwrap_saved(layer.input_layernorm)
            ops.layernorm(br, _check(ln.weight, "layernorm.weight"), _check(ln.bias, "layernorm.bias"), ln.eps, True,
                          False, plan.n_rows, x)
            fc1, fc2 = _unwrap_linear(layer.mlp.fc1), _unwrap_linear(layer.mlp.fc2)
            ops.linear_bias_act(x, _check(fc1.weight, "fc1.weight"), _check(fc1.bias, "fc1.bias"), mid, plan.n_rows,
                                None, False, True)
            ops.linear_bias_act(mid, _check(fc2.weight, "fc2.weight"), _check(fc2.bias, "fc2.bias"), br, plan.n_rows,
                                None, False, False)
            ln = _unwrap_saved(layer.post_attention_layernorm)
            ops.layernorm(br, _check(ln.weight, "layernorm.weight"), _check(ln.bias, "layernorm.bias"), ln.eps, True,
                          False, plan.n_rows, x)
        del qkv, ctx, mid

        # ---- class-token drop + optional max-pool (visual.py:197-202) into compact feature rows
        M = plan.M
        fin = bf(M, C)
        for im in plan.images:
            pool = im.pool if any(p > 1 for p in im.pool) else (1, 1, 1)
            ops.maxpool_tokens(x[im.start + 1:im.start + 1 + im.n], list(im.grid), list(pool),
                               fin[im.fstart:im.fstart + im.m])

        # ---- GLU projector (visual.py:172-177), last GEMM scatters into dest
        glu = self.linear_proj
        Hl = dest.shape[-1]
        lin = lambda m, n: _check(_unwrap_linear(m).weight, n)
        p1 = bf(M, Hl)
        ops.linear_bias_act(fin, lin(glu.linear_proj, "linear_proj.weight"), None, p1, plan.n_feat, None, False, False)
        n1 = _unwrap_saved(glu.norm1)
        p2 = bf(M, Hl)
        ops.layernorm(p1, _check(n1.weight, "norm1.weight"), _check(n1.bias, "norm1.bias"), n1.eps, False, True,
                      plan.n_feat, p2)
        wg, wu = lin(glu.gate_proj, "gate_proj.weight"), lin(glu.dense_h_to_4h, "dense_h_to_4h.weight")
        act = bf(M, wg.shape[0])
        ops.grouped_gemm_raw(p2, [wg, wu, None, None], act, plan.n_feat, ops.EPI_SWIGLU, single_expert=True)
        fmap = plan.derived.get(("fmap", tuple(feat_starts)), [],
                                lambda: plan.feature_row_map([s + 1 for s in feat_starts], dev))
        ops.linear_bias_act(act, lin(glu.dense_4h_to_h, "dense_4h_to_h.weight"), None, dest, plan.n_feat, fmap, False,
                            False)
        # ---- boi / eoi rows (visual.py:204-206)
        boi, eoi = _check(self.boi, "boi"), _check(self.eoi, "eoi")
        table = self._derived.get(("boi_eoi",), [boi, eoi], lambda: torch.cat([boi.reshape(1, -1), eoi.reshape(1, -1)]))
        src, dst = plan.derived.get(
            ("be_map", tuple(feat_starts)), [],
            lambda: (torch.tensor([0, 1] * plan.B, dtype=torch.int32, device=dev),
                     torch.tensor([r for im, s in zip(plan.images, feat_starts) for r in (s, s + 1 + im.m)],
                                  dtype=torch.int32, device=dev)))
        ops.scatter_rows(table, src, dst, dest)
        return plan

    def forward(self, image: List[torch.Tensor], patch_size: List[Tuple[int, int, int]],
                pool_size_list: List[Tuple[int, int, int]]) -> List[torch.Tensor]:
        """``EVA2CLIPModel.forward`` (visual.py:191-208): one [1, n_i + 2, hidden] tensor per image."""
        shapes = [tuple(im.shape) for im in image]
        plan = self._plan(shapes, patch_size, pool_size_list, image[0].device)
        starts, t = [], 0
        for im in plan.images:
            starts.append(t)
            t += im.m + 2
        dest = torch.empty(t, self.boi.shape[-1], dtype=torch.bfloat16, device=image[0].device)
        self._encode(image, patch_size, pool_size_list, dest, starts)
        return [dest[s:s + im.m + 2][None] for s, im in zip(starts, plan.images)]

    def encode_into(self, inputs_embeds: torch.Tensor, image: List[torch.Tensor], patch_size, pool_size_list):
        """``CogVLMModel.forward``'s multi-modality branch (modeling_cogvlm.py:447-453) fused: sample i's
        [boi, features, eoi] rows are written IN PLACE over columns [1, 1 + n_i + 2) of ``inputs_embeds`` [B, L, H]
        by the projector's last GEMM (no intermediate feature list).  Returns ``inputs_embeds``."""
        if inputs_embeds.dim() != 3 or not inputs_embeds.is_cuda or inputs_embeds.dtype != torch.bfloat16 \
                or not inputs_embeds.is_contiguous():
            raise ValueError("inputs_embeds must be a contiguous CUDA bf16 [B, L, H] tensor")
        B, L, H = inputs_embeds.shape
        if len(image) != B:
            raise ValueError(f"batch size mismatch: {B} {len(image)}")  # modeling_cogvlm.py:448
        plan = self._plan([tuple(im.shape) for im in image], patch_size, pool_size_list, inputs_embeds.device)
        for im in plan.images:
            if 1 + im.m + 2 > L:
                raise ValueError("image features do not fit into the sequence")
        self._encode(image, patch_size, pool_size_list, inputs_embeds.view(B * L, H), [i * L + 1 for i in range(B)])
        return inputs_embeds
