"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference decoder layer for oracle validation.

The reference file ``/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py`` cannot be imported
directly in this image: it pulls ``luolib`` (-> monai), ``mmmm.utils`` (-> cytoolz),
``mmmm.data.utils`` (-> monai, nibabel) and, inside ``attention_fn`` (modeling_cogvlm.py:113),
``xformers``.  None of these are installed and there is no network.  This loader installs tiny
``sys.modules`` stubs for exactly the names the file imports, executes the
reference source *from where it lies* (nothing is copied into this repo) and substitutes
``attention_fn`` -- whose arithmetic lives in the absent xformers 0.0.27 -- with an equivalent
block-diagonal-causal softmax attention written against xformers' documented semantics.

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


def block_diag_causal_attention(q, k, v, padding_mask, dropout_p: float = 0.0):
    """Stand-in for ``attention_fn`` (modeling_cogvlm.py:106-142), prefill branch.

    xformers ``memory_efficient_attention(q, k, v, BlockDiagonalCausalMask)`` semantics: per sample,
    the tokens with ``padding_mask == True`` are compacted (order preserved); token i attends to
    compacted tokens j <= i of the same sample; scale = head_dim ** -0.5; softmax in fp32; output in
    the input dtype.  Rows with ``padding_mask == False`` are zero (modeling_cogvlm.py:119,126).
    Inputs/outputs are [B, heads, L, head_dim] like the reference function.
    """
    assert dropout_p == 0.0
    B, H, L, D = q.shape
    if padding_mask.shape[1] != L:
        raise NotImplementedError("decode branch (q_len == 1 with cache) is not part of the oracle")
    out = torch.zeros_like(q)
    scale = D ** -0.5
    for b in range(B):
        idx = padding_mask[b].nonzero(as_tuple=True)[0]
        n = idx.numel()
        if n == 0:
            continue
        qb = q[b, :, idx].float()
        kb = k[b, :, idx].float()
        vb = v[b, :, idx]
        s = torch.matmul(qb, kb.transpose(-1, -2)) * scale
        causal = torch.ones(n, n, dtype=torch.bool, device=q.device).tril()
        s = s.masked_fill(~causal, float("-inf"))
        p = torch.softmax(s, dim=-1)
        # FA-style kernels round P to the value dtype before the PV product
        ob = torch.matmul(p.to(vb.dtype).float(), vb.float())
        out[b, :, idx] = ob.to(q.dtype)
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

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("mmmm.models.cogvlm.configuration_cogvlm", ref + "/models/cogvlm/configuration_cogvlm.py")
    M = _load("mmmm.models.cogvlm.modeling_cogvlm", ref + "/models/cogvlm/modeling_cogvlm.py")
    M.attention_fn = block_diag_causal_attention  # xformers is absent (see module docstring)
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
