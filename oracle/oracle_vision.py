"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference vision encoder.

Restates, in plain eager PyTorch over a flat ``weights`` dict (state-dict keys of the reference
``EVA2CLIPModel``), the algorithm of ``/root/reference/mmmm/models/cogvlm/visual.py:24-208`` together with the two
helpers it calls: ``Downsample.forward`` (``mmmm/models/resample.py:56-63``) and ``spadop.resample``
(``third-party/LuoLib/src/luolib/models/spadop/resample.py:11-29``, vendored under ``/root/reference``), and the
feature scatter of ``CogVLMModel.forward`` (``modeling_cogvlm.py:450-453``).  It is *not* the product: only
``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import it.

Pinning status: the reference has no tests or golden vectors for this path, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF (``oracle/reference_loader.load_reference_visual`` executes the unmodified
``visual.py``, ``resample.py`` and the luolib helpers from where they lie): ``oracle/make_golden.py`` writes
``tests/golden/vision_*.pt`` and ``tests/test_oracle_vision.py`` checks oracle == fixtures everywhere and
oracle == live reference (bit-exact, fp32 and bf16) wherever ``/root/reference`` exists.  One piece of arithmetic
lives in third-party code absent from ``/root/reference`` and is restated from its documented behaviour ("parity
unpinned" by any reference-side test): xformers 0.0.27 ``memory_efficient_attention`` under a
``BlockDiagonalMask`` (environment.yaml:41; call site visual.py:89-99) -- non-causal attention inside each image's
block, scale as passed, fp32 softmax, probabilities rounded to the value dtype before P.V.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class VisionConfig:
    """``config.vision_config`` of the checkpoint (THUDM/cogvlm-chat-hf config.json: EVA2-CLIP-E, 63 layers of
    1792 = 16 heads x 112) with VividMed's ``vision_override`` (conf/model.yaml:4-7) and the language-side sizes the
    GLU projector uses (visual.py:162-169 reads ``config.hidden_size`` / ``config.intermediate_size``)."""
    hidden_size: int = 1792
    num_heads: int = 16
    intermediate_size: int = 15360
    num_hidden_layers: int = 63
    layer_norm_eps: float = 1e-6
    in_channels: int = 3
    patch_size: Tuple[int, int, int] = (16, 16, 16)
    pos_embed_shape: Tuple[int, int, int] = (8, 32, 32)
    hidden_act: str = "gelu"
    lm_hidden_size: int = 4096
    lm_intermediate_size: int = 11008
    extra: dict = field(default_factory=dict)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads


# ----------------------------------------------------------------------------------------- helpers the path calls
def resample(x: torch.Tensor, shape: Sequence[int], upsample_mode=None, scale: bool = False) -> torch.Tensor:
    """``luolib.models.spadop.resample`` (resample.py:11-29): area interpolation down to min(shape), then
    trilinear (5-D) / bicubic (4-D) up; identity when the shape already matches."""
    shape = tuple(int(s) for s in shape)
    scale_ratio = np.prod(x.shape[2:]) / np.prod(shape) if scale else 1.0
    down = tuple(np.minimum(x.shape[2:], shape).tolist())
    if down != tuple(x.shape[2:]):
        x = F.interpolate(x, down, mode="area")
    if shape != tuple(x.shape[2:]):
        if upsample_mode is None:
            upsample_mode = "trilinear" if x.ndim == 5 else "bicubic"
        x = F.interpolate(x, shape, mode=upsample_mode)
    if scale:
        x = x * scale_ratio
    return x


def downsample_weight(weight: torch.Tensor, kernel_size: Sequence[int]) -> torch.Tensor:
    """``Downsample.forward`` weight selection (mmmm/models/resample.py:56-62): when the requested depth patch is
    smaller than the module's, groups of ``module_depth / depth`` kernel slices are summed."""
    kd = weight.shape[2]
    if kd == kernel_size[0]:
        return weight
    if kd % kernel_size[0] != 0:
        raise NotImplementedError
    d = int(kernel_size[0])
    co, ci, _, h, w = weight.shape
    return weight.reshape(co, ci, d, kd // d, h, w).sum(dim=3)


# ----------------------------------------------------------------------------------------- visual.py:59-77
def patch_embedding(w: Dict[str, torch.Tensor], image_list: List[torch.Tensor],
                    patch_size_list: List[Sequence[int]], prefix: str = "patch_embedding."):
    """``PatchEmbedding.forward`` (visual.py:59-77).  Per image [C, D, H, W]: strided 3-D convolution with the
    (depth-reduced) patch kernel, + the position embedding resampled to the patch grid, flattened to tokens in
    (d, h, w) order behind one class token (cls_embedding + cls_pos_embed); the images are packed along dim 1
    (``BlockDiagonalMask.from_tensor_list``).  Returns x [1, sum(1 + n_i), C], per-image lengths and grid shapes."""
    xs, shapes = [], []
    pos_full = w[prefix + "position_embedding.weight"]
    cls_row = (w[prefix + "cls_embedding.weight"] + w[prefix + "cls_pos_embed.weight"])[None]  # [1, 1, C]
    for image, ps in zip(image_list, patch_size_list):
        ps = tuple(int(p) for p in ps)
        x = F.conv3d(image[None], downsample_weight(w[prefix + "proj.weight"], ps), w.get(prefix + "proj.bias"), ps)
        grid = tuple(x.shape[2:])
        shapes.append(grid)
        pos = resample(pos_full, grid)
        x = (x + pos).flatten(2).transpose(1, 2)  # '1 c ... -> 1 (...) c'
        xs.append(torch.cat([cls_row, x], dim=1))
    return torch.cat(xs, dim=1), [t.shape[1] for t in xs], shapes


# ----------------------------------------------------------------------------------------- visual.py:79-102
def blockdiag_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, seqlens: Sequence[int], scale: float):
    """xformers ``memory_efficient_attention(q, k, v, BlockDiagonalMask, scale=)`` (call site visual.py:96-98):
    q, k, v [1, T, heads, d]; every token attends to ALL tokens of its own image (non-causal)."""
    out = torch.empty_like(q)
    s0 = 0
    for n in seqlens:
        sl = slice(s0, s0 + n)
        qb, kb, vb = (t[0, sl].permute(1, 0, 2) for t in (q, k, v))
        s = torch.matmul(qb.float(), kb.float().transpose(-1, -2)) * scale
        p = torch.softmax(s, dim=-1)
        out[0, sl] = torch.matmul(p.to(vb.dtype).float(), vb.float()).permute(1, 0, 2).to(q.dtype)
        s0 += n
    return out


def attention(w, prefix: str, x: torch.Tensor, seqlens: Sequence[int], num_heads: int) -> torch.Tensor:
    """``Attention.forward`` (visual.py:89-102): fused QKV Linear (with bias), columns laid out as
    (qkv, head, d); attention; output Linear (with bias); dropout_prob = 0."""
    B, L, C = x.shape
    qkv = F.linear(x, w[prefix + "query_key_value.weight"], w[prefix + "query_key_value.bias"])
    qkv = qkv.reshape(B, L, 3, num_heads, -1).permute(2, 0, 1, 3, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    out = blockdiag_attention(q, k, v, seqlens, (C // num_heads) ** -0.5)
    return F.linear(out.reshape(B, L, -1), w[prefix + "dense.weight"], w[prefix + "dense.bias"])


def mlp(w, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``MLP.forward`` (visual.py:113-117): fc1 -> ACT2FN['gelu'] (exact erf GELU) -> fc2, both with bias."""
    x = F.linear(x, w[prefix + "fc1.weight"], w[prefix + "fc1.bias"])
    x = F.gelu(x)
    return F.linear(x, w[prefix + "fc2.weight"], w[prefix + "fc2.bias"])


def transformer_layer(w, prefix: str, x: torch.Tensor, seqlens, num_heads: int, eps: float) -> torch.Tensor:
    """``TransformerLayer.forward`` (visual.py:128-135): the LayerNorm is applied to the BRANCH OUTPUT
    (h = x + LN(attn(x)); out = h + LN(mlp(h))), not to the branch input."""
    C = x.shape[-1]
    a = attention(w, prefix + "attention.", x, seqlens, num_heads)
    h = x + F.layer_norm(a, (C,), w[prefix + "input_layernorm.weight"], w[prefix + "input_layernorm.bias"], eps)
    m = mlp(w, prefix + "mlp.", h)
    return h + F.layer_norm(m, (C,), w[prefix + "post_attention_layernorm.weight"],
                            w[prefix + "post_attention_layernorm.bias"], eps)


def glu(w, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``GLU.forward`` (visual.py:172-177): linear_proj -> LayerNorm (default eps 1e-5) -> GELU ->
    silu(gate_proj(x)) * dense_h_to_4h(x) -> dense_4h_to_h; no biases on the Linears."""
    x = F.linear(x, w[prefix + "linear_proj.weight"])
    x = F.gelu(F.layer_norm(x, (x.shape[-1],), w[prefix + "norm1.weight"], w[prefix + "norm1.bias"], 1e-5))
    x = F.silu(F.linear(x, w[prefix + "gate_proj.weight"])) * F.linear(x, w[prefix + "dense_h_to_4h.weight"])
    return F.linear(x, w[prefix + "dense_4h_to_h.weight"])


# ----------------------------------------------------------------------------------------- visual.py:181-208
def eva2clip(w, image_list, patch_size_list, pool_size_list, cfg: VisionConfig) -> List[torch.Tensor]:
    """``EVA2CLIPModel.forward`` (visual.py:191-208): patch embedding -> transformer -> per image: drop the class
    token, optional 3-D max-pool over the patch grid, GLU projector, boi / eoi rows around the features.
    Returns one [1, n_i + 2, lm_hidden] tensor per image."""
    x, seqlens, shapes = patch_embedding(w, image_list, patch_size_list)
    for i in range(cfg.num_hidden_layers):
        x = transformer_layer(w, f"transformer.layers.{i}.", x, seqlens, cfg.num_heads, cfg.layer_norm_eps)
    outs, s0 = [], 0
    for n, shape, pool in zip(seqlens, shapes, pool_size_list):
        xi = x[:, s0 + 1:s0 + n]
        s0 += n
        if any(int(p) > 1 for p in pool):
            C = xi.shape[-1]
            # spatialize (luolib/utils/einops.py:34-51): 'n (s0 s1 s2) c -> n c s0 s1 s2' as a permuted VIEW, like
            # einops builds it (the memory format decides which pooling / GEMM kernels ATen picks afterwards)
            xi = xi.reshape(1, *shape, C).permute(0, 4, 1, 2, 3)
            xi = F.max_pool3d(xi, tuple(int(p) for p in pool))
            xi = xi.reshape(1, C, -1).permute(0, 2, 1)             # flatten (luolib/utils/einops.py:31-32)
        xi = glu(w, "linear_proj.", xi)
        outs.append(torch.cat((w["boi"].expand(1, -1, -1), xi, w["eoi"].expand(1, -1, -1)), dim=1))
    return outs


def scatter_image_features(inputs_embeds: torch.Tensor, features: List[torch.Tensor]) -> torch.Tensor:
    """``CogVLMModel.forward`` (modeling_cogvlm.py:450-453): sample i's image rows overwrite columns
    [1, 1 + n_i + 2) of its text embeddings."""
    out = inputs_embeds.clone()
    for i, f in enumerate(features):
        out[i, 1:1 + f.shape[1]] = f[0]
    return out


# ----------------------------------------------------------------------------------------- synthetic weights
def random_vision_weights(cfg: VisionConfig, *, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights under the reference's state-dict keys (bf16-representable values): Linear /
    conv weights ~ N(0, 0.02), biases ~ N(0, 0.02), norm weights 1 + N(0, 0.1), embeddings ~ N(0, 0.02) (the
    reference initialises them to zero, visual.py:33-36; non-zero values exercise the adds)."""
    g = torch.Generator().manual_seed(seed)
    C, I, Hl, Il = cfg.hidden_size, cfg.intermediate_size, cfg.lm_hidden_size, cfg.lm_intermediate_size

    def rn(*shape, std=0.02, mean=0.0):
        return (mean + std * torch.randn(*shape, generator=g)).to(torch.bfloat16).to(dtype)

    w = {
        "patch_embedding.proj.weight": rn(C, cfg.in_channels, *cfg.patch_size),
        "patch_embedding.proj.bias": rn(C),
        "patch_embedding.cls_embedding.weight": rn(1, C),
        "patch_embedding.cls_pos_embed.weight": rn(1, C),
        "patch_embedding.position_embedding.weight": rn(1, C, *cfg.pos_embed_shape),
        "boi": rn(1, 1, Hl), "eoi": rn(1, 1, Hl),
        "linear_proj.linear_proj.weight": rn(Hl, C),
        "linear_proj.norm1.weight": rn(Hl, std=0.1, mean=1.0), "linear_proj.norm1.bias": rn(Hl),
        "linear_proj.dense_h_to_4h.weight": rn(Il, Hl), "linear_proj.gate_proj.weight": rn(Il, Hl),
        "linear_proj.dense_4h_to_h.weight": rn(Hl, Il),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"transformer.layers.{i}."
        w[p + "input_layernorm.weight"] = rn(C, std=0.1, mean=1.0)
        w[p + "input_layernorm.bias"] = rn(C)
        w[p + "attention.query_key_value.weight"] = rn(3 * C, C)
        w[p + "attention.query_key_value.bias"] = rn(3 * C)
        w[p + "attention.dense.weight"] = rn(C, C)
        w[p + "attention.dense.bias"] = rn(C)
        w[p + "mlp.fc1.weight"] = rn(I, C)
        w[p + "mlp.fc1.bias"] = rn(I)
        w[p + "mlp.fc2.weight"] = rn(C, I)
        w[p + "mlp.fc2.bias"] = rn(C)
        w[p + "post_attention_layernorm.weight"] = rn(C, std=0.1, mean=1.0)
        w[p + "post_attention_layernorm.bias"] = rn(C)
    return w


def random_images(shapes: Sequence[Sequence[int]], *, in_channels: int = 3, seed: int = 0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(in_channels, *s, generator=g).to(torch.bfloat16).to(dtype) for s in shapes]
