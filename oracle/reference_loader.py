"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference decoder layer for oracle validation.

The reference file ``/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py`` cannot be imported
directly in this image: it pulls ``luolib`` (-> monai), ``mmmm.utils`` (-> cytoolz),
``mmmm.data.utils`` (-> monai, nibabel) and, inside ``attention_fn`` (modeling_cogvlm.py:113),
``xformers``.  None of these are installed and there is no network.  This loader installs tiny
``sys.modules`` stubs for exactly the names the file imports and executes the
reference source *from where it lies* (nothing is copied into this repo, nothing is patched).  The two
xformers 0.0.27 entry points ``attention_fn`` calls (``memory_efficient_attention`` and
``BlockDiagonalCausalMask.from_tensor_lists_qkv``) are stand-ins written against xformers' documented
semantics; the reference's own ``attention_fn`` -- including its generation branch -- runs unmodified.

Only ``tests/`` and ``oracle/make_golden.py`` may import this module.  It only works where
``/root/reference`` exists (the build container); the GPU box uses the committed fixtures under
``tests/golden/`` and the in-repo restatement ``oracle/oracle_layer.py``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("MMMM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mmmm/models/cogvlm/modeling_cogvlm.py"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    return m


class BlockDiagonalCausalMask:
    """Stand-in for ``xformers.ops.fmha.attn_bias.BlockDiagonalCausalMask`` (xformers 0.0.27, absent here):
    remembers the per-sample lengths; ``from_tensor_lists_qkv`` concatenates the per-sample [1, n_i, H, D]
    tensors along dim 1, as documented."""

    def __init__(self, seqlens):
        self.seqlens = list(seqlens)

    @classmethod
    def from_tensor_lists_qkv(cls, q_list, k_list, v_list):
        bias = cls([q.shape[1] for q in q_list])
        return bias, torch.cat(q_list, dim=1), torch.cat(k_list, dim=1), torch.cat(v_list, dim=1)


def memory_efficient_attention(q, k, v, attn_bias=None, p: float = 0.0):
    """Stand-in for ``xformers.ops.memory_efficient_attention`` under a ``BlockDiagonalCausalMask``:
    q, k, v are [1, T, H, D] (all samples packed); inside block i token a attends to tokens b <= a of the same
    block; scale = D ** -0.5; softmax in fp32; probabilities rounded to the value dtype before P @ V (what the
    flash kernels do); output in the input dtype."""
    assert p == 0.0 and isinstance(attn_bias, BlockDiagonalCausalMask)
    out = torch.empty_like(q)
    scale = q.shape[-1] ** -0.5
    start = 0
    for n in attn_bias.seqlens:
        if n == 0:
            continue
        sl = slice(start, start + n)
        qb = q[0, sl].permute(1, 0, 2).float()   # [H, n, D]
        kb = k[0, sl].permute(1, 0, 2).float()
        vb = v[0, sl].permute(1, 0, 2)
        s = torch.matmul(qb, kb.transpose(-1, -2)) * scale
        causal = torch.ones(n, n, dtype=torch.bool, device=q.device).tril()
        s = s.masked_fill(~causal, float("-inf"))
        pr = torch.softmax(s, dim=-1)
        ob = torch.matmul(pr.to(vb.dtype).float(), vb.float())
        out[0, sl] = ob.permute(1, 0, 2).to(q.dtype)
        start += n
    return out


_LOADED = None


def load_reference():
    """Returns the executed reference module ``mmmm.models.cogvlm.modeling_cogvlm``."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")

    class NoWeightDecayParameter(nn.Parameter):  # luolib/models/param.py:3
        pass

    def forward_gc(model, enable, gc_func, *args, **kwargs):  # luolib/models/utils.py:34-37
        return gc_func(model, *args, **kwargs) if enable else model(*args, **kwargs)

    _stub("luolib")
    _stub("luolib.models")
    _stub("luolib.models.param", NoWeightDecayParameter=NoWeightDecayParameter)
    _stub("luolib.models.utils", forward_gc=forward_gc)
    _stub("luolib.types", tuple3_t=tuple, tuple2_t=tuple)
    ref = os.path.join(REFERENCE_ROOT, "mmmm")
    for pkg in ["mmmm", "mmmm.models", "mmmm.models.cogvlm", "mmmm.data"]:
        _stub(pkg).__path__ = [ref + pkg[4:].replace(".", "/")]

    def apply_prefix(prefix: str, path: str):  # mmmm/utils.py:8-9
        return f"{prefix}{path}" if prefix.endswith(".") or not prefix else f"{prefix}.{path}"

    _stub("mmmm.utils", apply_prefix=apply_prefix, get_lora_modules_default=None)
    _stub("mmmm.data.defs", CE_IGNORE_INDEX=-100)
    _stub("mmmm.data.utils", LANGUAGE_TOKEN_TYPE=0, VISION_TOKEN_TYPE=1)  # mmmm/data/utils.py:192-193
    _stub("mmmm.models.cogvlm.visual", EVA2CLIPModel=nn.Identity)
    # xformers is absent: attention_fn (modeling_cogvlm.py:106-142) runs UNMODIFIED against these two stand-ins
    _stub("xformers")
    _stub("xformers.ops", memory_efficient_attention=memory_efficient_attention)
    _stub("xformers.ops.fmha")
    _stub("xformers.ops.fmha.attn_bias", BlockDiagonalCausalMask=BlockDiagonalCausalMask)

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("mmmm.models.cogvlm.configuration_cogvlm", ref + "/models/cogvlm/configuration_cogvlm.py")
    M = _load("mmmm.models.cogvlm.modeling_cogvlm", ref + "/models/cogvlm/modeling_cogvlm.py")
    _LOADED = M
    return M


def make_reference_layer(hidden_size=4096, intermediate_size=11008, num_heads=32, rms_norm_eps=1e-6,
                         dtype=torch.float32, seed=0):
    """A random-init reference ``CogVLMDecoderLayer`` (Linear ~ N(0, 0.02), modeling_cogvlm.py:350-355)."""
    M = load_reference()
    cfg = M.CogVLMConfig(hidden_size=hidden_size, intermediate_size=intermediate_size,
                         num_attention_heads=num_heads, rms_norm_eps=rms_norm_eps, num_hidden_layers=1)
    cfg.lora_lang = True  # set by MMMMForCausalLM.build (mmmm/models/mmmm.py:134)
    g = torch.Generator().manual_seed(seed)
    layer = M.CogVLMDecoderLayer(cfg)
    with torch.no_grad():
        for mod in layer.modules():
            if isinstance(mod, nn.Linear):
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * cfg.initializer_range)
        for norm in (layer.input_layernorm, layer.post_attention_layernorm):
            norm.weight.copy_(1.0 + 0.1 * torch.randn(norm.weight.shape, generator=g))
    layer = layer.to(dtype)
    # the cos/sin cache is grow-only and keeps its creation dtype (modeling_cogvlm.py:172-180)
    layer.self_attn.rotary_emb.max_seq_len_cached = 0
    return layer, cfg
