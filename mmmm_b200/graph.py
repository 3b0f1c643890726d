"""CUDA-graph replay of the prefill forward (one layer or a stack of layers).

The fused path has no host synchronisation and sizes every grid from host-known bounds, so a whole forward --
K1 partition included -- captures into one CUDA graph: replay removes the Python / dispatcher / launch
overhead between the ~9 kernels of a layer (B200 guidance: streams and graphs instead of a tracing compiler).
Shapes [B, L, H] are fixed per graph; the *contents* of all four inputs may change between replays (routing is
recomputed on the device inside the graph)."""
from __future__ import annotations

from typing import Sequence, Union

import torch
from torch import nn

from .plan import GLOBAL_PLAN_CACHE


class GraphedPrefill:
    def __init__(self, layers: Union[nn.Module, Sequence[nn.Module]], hidden_states: torch.Tensor,
                 token_type_ids: torch.Tensor, position_ids: torch.Tensor, padding_mask: torch.Tensor,
                 final_norm: nn.Module = None, warmup: int = 2):
        self.layers = list(layers) if isinstance(layers, (list, tuple, nn.ModuleList)) else [layers]
        self.final_norm = final_norm
        # static input buffers: refill them (copy_) and call replay()
        self.hidden_states = hidden_states.clone()
        self.token_type_ids = token_type_ids.clone()
        self.position_ids = position_ids.clone()
        self.padding_mask = padding_mask.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # first-call work (library load, func attributes, rotary tables) outside capture
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        GLOBAL_PLAN_CACHE.clear()  # K1 must be captured too: routing depends on the (mutable) static id buffers
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.output = self._forward()
        GLOBAL_PLAN_CACHE.clear()

    def _forward(self) -> torch.Tensor:
        h = self.hidden_states
        for layer in self.layers:
            h = layer(h, token_type_ids=self.token_type_ids, position_ids=self.position_ids,
                      padding_mask=self.padding_mask)[0]
        if self.final_norm is not None:  # masked like the reference caller (:570-573)
            from .modeling_cogvlm import masked_rms_norm
            h = masked_rms_norm(self.final_norm, h, self.token_type_ids, self.padding_mask)
        return h

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.output

    def __call__(self, hidden_states, token_type_ids, position_ids, padding_mask) -> torch.Tensor:
        self.hidden_states.copy_(hidden_states, non_blocking=True)
        self.token_type_ids.copy_(token_type_ids, non_blocking=True)
        self.position_ids.copy_(position_ids, non_blocking=True)
        self.padding_mask.copy_(padding_mask, non_blocking=True)
        return self.replay()
