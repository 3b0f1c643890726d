"""Reading weights THROUGH PEFT wrappers.

After ``get_peft_model`` (scripts/cli.py:82-88 of the reference) the decoder layer's ten
``nn.Linear`` children are ``peft.tuners.lora.Linear`` modules and its two RMSNorms are
``peft.utils.ModulesToSaveWrapper`` modules (targets chosen by mmmm/utils.py:19-43).  The reference
layer *calls* those children; the fused path must instead read their tensors:

    lora.Linear          : .base_layer.weight, .lora_A[name].weight [r, in], .lora_B[name].weight [out, r],
                           .scaling[name], .active_adapters, .disable_adapters, .merged
    ModulesToSaveWrapper : .original_module, .modules_to_save[name], .active_adapters, .disable_adapters

PEFT is not installed in this image, so the resolver is duck-typed on those attribute names and is
exercised with the stand-ins below (``MockLoraLinear`` / ``MockModulesToSave``), which mimic PEFT's
attribute layout and state-dict key names (``...lora_A.default.weight``); "real-PEFT unverified".
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn


@dataclass
class LinearSpec:
    weight: torch.Tensor                 # [out, in] frozen base weight
    lora_A: Optional[torch.Tensor] = None  # [r, in]
    lora_B: Optional[torch.Tensor] = None  # [out, r]
    scaling: float = 1.0
    dropout: float = 0.0                 # p of lora_dropout when the wrapper is in training mode, else 0

    @property
    def r(self) -> int:
        return 0 if self.lora_A is None else self.lora_A.shape[0]


def _active_names(mod) -> list:
    names = getattr(mod, "active_adapters", None)
    if names is None:
        names = getattr(mod, "active_adapter", [])
    if isinstance(names, str):
        names = [names]
    return list(names)


def resolve_linear(mod: nn.Module) -> LinearSpec:
    """``nn.Linear`` or a PEFT-style LoRA wrapper around one -> tensors for the fused GEMM."""
    if hasattr(mod, "base_layer") and hasattr(mod, "lora_A"):
        base = mod.base_layer
        while hasattr(base, "base_layer"):
            base = base.base_layer
        if getattr(base, "bias", None) is not None:
            raise NotImplementedError("biased Linear is not part of the visual-expert layer (bias=False, :206-209)")
        spec = LinearSpec(base.weight)
        if getattr(mod, "disable_adapters", False) or getattr(mod, "merged", False):
            return spec  # disabled, or the delta already lives in base.weight
        active = [n for n in _active_names(mod) if n in mod.lora_A]
        if not active:
            return spec
        if len(active) > 1:
            raise NotImplementedError("more than one active LoRA adapter on a Linear is not supported by the "
                                      "fused K-extension")
        name = active[0]
        if getattr(mod, "use_dora", {}).get(name, False) if isinstance(getattr(mod, "use_dora", None), dict) else False:
            raise NotImplementedError("DoRA adapters are not supported")
        drop = mod.lora_dropout[name] if hasattr(mod, "lora_dropout") and name in mod.lora_dropout else None
        if mod.training and isinstance(drop, nn.Dropout) and drop.p > 0:
            if drop.p >= 1:
                raise ValueError("lora_dropout must be < 1")
            spec.dropout = float(drop.p)  # active exactly when nn.Dropout would be (module in training mode)
        spec.lora_A = mod.lora_A[name].weight
        spec.lora_B = mod.lora_B[name].weight
        spec.scaling = float(mod.scaling[name])
        return spec
    if isinstance(mod, nn.Linear):
        if mod.bias is not None:
            raise NotImplementedError("biased Linear is not part of the visual-expert layer")
        return LinearSpec(mod.weight)
    raise TypeError(f"cannot resolve weights of {type(mod).__name__}")


def resolve_norm(mod: nn.Module) -> nn.Module:
    """RMSNorm or a ModulesToSaveWrapper-style wrapper -> the module whose ``weight`` is live."""
    if hasattr(mod, "modules_to_save") and hasattr(mod, "original_module"):
        if getattr(mod, "disable_adapters", False):
            return mod.original_module
        for name in _active_names(mod):
            if name in mod.modules_to_save:
                return mod.modules_to_save[name]
        return mod.original_module
    return mod


# --------------------------------------------------------------------------------------------------
# stand-ins with PEFT's attribute layout, for tests and for bench.py's LoRA-active configuration
# --------------------------------------------------------------------------------------------------
class MockLoraLinear(nn.Module):
    """Attribute-compatible stand-in for ``peft.tuners.lora.Linear`` (rsLoRA: scaling = alpha / sqrt(r))."""

    def __init__(self, base_layer: nn.Linear, r: int = 64, lora_alpha: float = 8, lora_dropout: float = 0.0,
                 use_rslora: bool = True, adapter_name: str = "default", b_std: float = 0.0):
        super().__init__()
        self.base_layer = base_layer
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(base_layer.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, base_layer.out_features, bias=False)})
        self.lora_dropout = nn.ModuleDict(
            {adapter_name: nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.scaling = {adapter_name: lora_alpha / math.sqrt(r) if use_rslora else lora_alpha / r}
        self.active_adapters = [adapter_name]
        self.disable_adapters = False
        self.merged = False
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        if b_std > 0:
            nn.init.normal_(self.lora_B[adapter_name].weight, std=b_std)
        else:
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        base_layer.weight.requires_grad_(False)
        self.to(base_layer.weight.device, base_layer.weight.dtype)

    @property
    def weight(self):
        return self.base_layer.weight

    def forward(self, x):  # PEFT lora.Linear.forward semantics
        y = self.base_layer(x)
        if self.disable_adapters or self.merged:
            return y
        for name in self.active_adapters:
            A, Bm = self.lora_A[name], self.lora_B[name]
            y = y + Bm(A(self.lora_dropout[name](x.to(A.weight.dtype)))) * self.scaling[name]
        return y.to(x.dtype)


class MockModulesToSave(nn.Module):
    """Attribute-compatible stand-in for ``peft.utils.ModulesToSaveWrapper``."""

    def __init__(self, module: nn.Module, adapter_name: str = "default"):
        super().__init__()
        import copy
        self.original_module = module
        self.modules_to_save = nn.ModuleDict({adapter_name: copy.deepcopy(module)})
        self.active_adapters = [adapter_name]
        self.disable_adapters = False
        self.original_module.requires_grad_(False)
        self.modules_to_save.requires_grad_(True)

    def forward(self, *a, **k):
        return resolve_norm(self)(*a, **k)


def attach_mock_lora(layer: nn.Module, r: int = 64, lora_alpha: float = 8, b_std: float = 0.02,
                     lora_lang: bool = True, wrap_norms: bool = True, lora_dropout: float = 0.0) -> nn.Module:
    """Wraps a decoder layer's children the way ``get_peft_model`` would with the targets of
    mmmm/utils.py:19-43: all ten Linears (vision-only when ``lora_lang`` is False,
    modeling_cogvlm.py:79-85, :211-220) and both RMSNorms as modules_to_save."""
    for p in layer.parameters():  # get_peft_model freezes every base parameter (mark_only_lora_as_trainable)
        p.requires_grad_(False)

    def wrap(parent, name):
        setattr(parent, name, MockLoraLinear(getattr(parent, name), r=r, lora_alpha=lora_alpha, b_std=b_std,
                                             lora_dropout=lora_dropout))

    wrap(layer.self_attn, "vision_expert_query_key_value")
    wrap(layer.self_attn, "vision_expert_dense")
    for n in ("gate_proj", "up_proj", "down_proj"):
        wrap(layer.mlp.vision_mlp, n)
    if lora_lang:
        wrap(layer.self_attn, "language_expert_query_key_value")
        wrap(layer.self_attn, "language_expert_dense")
        for n in ("gate_proj", "up_proj", "down_proj"):
            wrap(layer.mlp.language_mlp, n)
    if wrap_norms:
        layer.input_layernorm = MockModulesToSave(layer.input_layernorm)
        layer.post_attention_layernorm = MockModulesToSave(layer.post_attention_layernorm)
    return layer
