"""Fused lm_head + sample-weighted cross-entropy.

Reference (``CogVLMForCausalLM.forward``, modeling_cogvlm.py:701-706 and ``_sample_weighted_ce`` :610-627):

    logits = self.lm_head(output.last_hidden_state).float()          # [B, L, V] fp32 for EVERY position
    loss = _sample_weighted_ce(logits, labels, weight)               # keeps only rows with labels != -100

Here the rows with a label are selected first (K10 ``vex_label_rows``, same ascending order as the reference's
boolean-mask indexing), gathered, and the vocabulary GEMM runs over those rows only with the softmax statistics
computed in its epilogue (``VEX_EPI_CE``): the logits never reach HBM.  The backward recomputes the logits tile by
tile, writes d(loss)/d(logits) as bf16 (``VEX_EPI_CE_BWD``) and pushes it through the lm_head dgrad GEMM (K3 with the
weight read as stored); rows without a label receive a zero gradient.  ``lm_head`` may be a plain ``nn.Linear`` or a
PEFT-style LoRA wrapper (it is a LoRA target under mmmm/utils.py:19-43); its base weight stays frozen.

The value equals the reference's up to fp32 summation order: the logits are rounded to bf16 before the softmax
exactly like ``lm_head(...)`` under bf16-true, the log-sum-exp and the weighted mean are fp32.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from .peft_compat import LinearSpec, resolve_linear

CE_IGNORE_INDEX = -100  # mmmm/data/defs.py


def _bf16(t: torch.Tensor) -> torch.Tensor:
    from .modeling_cogvlm import _bf16 as cast
    return cast(t)


class _LMHeadCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden_states, labels, weight, spec: LinearSpec, *trainables):
        B, L, H = hidden_states.shape
        n = B * L
        dev = hidden_states.device
        V = spec.weight.shape[0]
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        row_idx, label_sel, w_sel = i32(n), i32(n), f32(n)
        counts = torch.zeros(4, dtype=torch.int32, device=dev)  # [selected rows, 0, ...]: single-"expert" GEMM counts
        ops.label_rows(labels.reshape(-1).contiguous(), None if weight is None else weight.reshape(-1).contiguous(),
                       CE_IGNORE_INDEX, row_idx, label_sel, w_sel, counts)
        h_sel = torch.empty(n, H, dtype=torch.bfloat16, device=dev)
        ops.gather_rows(hidden_states.reshape(n, H), row_idx, counts[:1], h_sel)
        t, r, lora_b = None, 0, None
        if spec.lora_A is not None:
            if spec.dropout > 0:
                raise NotImplementedError("lora_dropout on lm_head is not implemented by the fused loss")
            r = spec.r
            t = torch.empty(n, r, dtype=torch.bfloat16, device=dev)
            ops.grouped_gemm(h_sel, _bf16(spec.lora_A), None, t, counts, None, float(spec.scaling))
            lora_b = _bf16(spec.lora_B)
        tiles = 2 * ((V + 255) // 256)  # one partial slot per 128-column half tile
        pmax, psum, zlabel, lse = f32(n, tiles), f32(n, tiles), f32(n), f32(n)
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        ops.lm_head_ce_forward(h_sel, _bf16(spec.weight), label_sel, w_sel, counts, t, lora_b, r, pmax, psum, zlabel,
                               lse, loss)
        ctx.spec, ctx.trainables, ctx.shape = spec, trainables, (B, L, H)
        ctx.save_for_backward(h_sel, row_idx, label_sel, w_sel, counts, lse, t if t is not None else lse)
        ctx.has_t = t is not None
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        h_sel, row_idx, label_sel, w_sel, counts, lse, t = ctx.saved_tensors
        spec: LinearSpec = ctx.spec
        B, L, H = ctx.shape
        n, dev = B * L, h_sel.device
        V = spec.weight.shape[0]
        w = _bf16(spec.weight)
        r = spec.r if ctx.has_t else 0
        with torch.no_grad():
            dl = dloss.detach().reshape(1).to(torch.float32).contiguous()
            dz = torch.empty(n, V, dtype=torch.bfloat16, device=dev)
            ops.lm_head_ce_backward(h_sel, w, label_sel, w_sel, counts, t if ctx.has_t else None,
                                    _bf16(spec.lora_B) if ctx.has_t else None, r, lse, dl, dz)
            grads = {}
            dt = None
            if ctx.has_t:
                dt = torch.empty(n, r, dtype=torch.bfloat16, device=dev)
                ops.grouped_gemm_dgrad(dz, [_bf16(spec.lora_B), None], dt, counts, False, None, None, [None, None], 0,
                                       True, float(spec.scaling))
                for p, x, y, tr in ((spec.lora_B, dz, t, False), (spec.lora_A, h_sel, dt, True)):
                    if p.requires_grad:
                        g = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                        ops.lora_wgrad(x, y, g, None, tr, counts)
                        grads[id(p)] = g if g.dtype == p.dtype else g.to(p.dtype)
            d_sel = torch.empty(n, H, dtype=torch.bfloat16, device=dev)
            ops.grouped_gemm_dgrad(dz, [w, None], d_sel, counts, False, None, dt,
                                   [_bf16(spec.lora_A) if ctx.has_t else None, None], r, True, 1.0)
            del dz
            d_hidden = None
            if ctx.needs_input_grad[0]:
                d_hidden = torch.zeros(n, H, dtype=torch.bfloat16, device=dev)  # rows without a label: zero gradient
                ops.residual_scatter(d_sel, d_hidden, row_idx, counts[:1], d_hidden)
                d_hidden = d_hidden.view(B, L, H)
        return (d_hidden, None, None, None, *(grads.get(id(p)) for p in ctx.trainables))


def fused_lm_head_loss(hidden_states: torch.Tensor, lm_head: nn.Module, labels: torch.Tensor,
                       weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``_sample_weighted_ce(lm_head(hidden_states).float(), labels, weight)`` without materialising the logits.

    hidden_states [B, L, H] bf16 CUDA (the decoder's last hidden state after the final norm); labels int64 [B, L]
    (already shifted by the data module, -100 = ignore); weight [B, L] bf16/fp32 per-position weights or None.
    Returns the fp32 scalar loss; differentiable w.r.t. hidden_states and lm_head's LoRA adapters."""
    if not hidden_states.is_cuda:
        raise ValueError("hidden_states must be a CUDA tensor: the fused loss has no CPU path")
    if hidden_states.dtype != torch.bfloat16:
        raise TypeError(f"hidden_states must be bfloat16, got {hidden_states.dtype}")
    if hidden_states.dim() != 3 or labels.shape != hidden_states.shape[:2]:
        raise ValueError("hidden_states must be [B, L, H] and labels [B, L]")
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    if weight is not None and weight.shape != labels.shape:
        raise ValueError("weight must be [B, L] like labels")
    spec = resolve_linear(lm_head)
    if spec.weight.requires_grad:
        raise NotImplementedError("a trainable lm_head base weight (modules_to_save) is not supported by the fused "
                                  "loss; LoRA adapters on lm_head are")
    if spec.weight.shape[1] != hidden_states.shape[-1] or spec.weight.shape[0] <= 64 or spec.weight.shape[0] % 8:
        raise ValueError("lm_head must be [V, H] with V > 64 and V % 8 == 0")
    tr = [p for p in (spec.lora_A, spec.lora_B) if p is not None and p.requires_grad]
    return _LMHeadCE.apply(hidden_states.contiguous(), labels, weight, spec, *tr)
