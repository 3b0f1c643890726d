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
    """``sorted_stream`` (default: on for a stack of layers, off for a single layer): run the stack through
    ``decoder_stack_forward`` -- residual stream in expert-sorted order across all layers -- instead
    of the per-layer module calls in the flat [B, L, H] layout."""

    def __init__(self, layers: Union[nn.Module, Sequence[nn.Module]], hidden_states: torch.Tensor,
                 token_type_ids: torch.Tensor, position_ids: torch.Tensor, padding_mask: torch.Tensor,
                 final_norm: nn.Module = None, warmup: int = 2, sorted_stream: bool = None):
        self.layers = list(layers) if isinstance(layers, (list, tuple, nn.ModuleList)) else [layers]
        self.final_norm = final_norm
        self.sorted_stream = (len(self.layers) > 1) if sorted_stream is None else bool(sorted_stream)
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
        if self.sorted_stream:
            from .modeling_cogvlm import decoder_stack_forward
            plan = GLOBAL_PLAN_CACHE.get(self.token_type_ids, self.padding_mask)
            return decoder_stack_forward(self.layers, self.final_norm, self.hidden_states, plan,
                                         self.position_ids.long().contiguous())[0]
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


class PipelinedHostPrefill:
    """Serving-loop form of the prefill for HOST buffers: ``depth`` requests in flight, each slot owning one
    captured graph (static device inputs + output) and one pinned host output.  ``submit`` queues the H2D copies on
    a copy-in stream, the graph replay on the caller's stream and the D2H copy on a copy-out stream, so the PCIe
    transfers of request i+1 / i-1 overlap the kernels of request i and the host does three stream operations per
    request instead of ~15 op dispatches.  Inputs should be pinned (pageable memory makes the copies synchronous).
    ``result(slot)`` blocks until that slot's output has landed in host memory."""

    def __init__(self, layers, hidden_states, token_type_ids, position_ids, padding_mask, final_norm=None,
                 depth: int = 2, device=None, sorted_stream: bool = None):
        dev = torch.device(device) if device is not None else (
            hidden_states.device if hidden_states.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        on_dev = lambda t: t.to(dev, non_blocking=False)
        ex = tuple(map(on_dev, (hidden_states, token_type_ids, position_ids, padding_mask)))
        self.slots = [GraphedPrefill(layers, *ex, final_norm=final_norm, sorted_stream=sorted_stream)
                      for _ in range(depth)]
        self.out_host = [torch.empty(ex[0].shape, dtype=ex[0].dtype).pin_memory() for _ in range(depth)]
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.ev_in, self.ev_free, self.ev_done, self.ev_copied = mk(), mk(), mk(), mk()
        self.count = 0

    def submit(self, hidden_states, token_type_ids, position_ids, padding_mask) -> int:
        k = self.count % len(self.slots)
        g = self.slots[k]
        cur = torch.cuda.current_stream()
        reuse = self.count >= len(self.slots)
        with torch.cuda.stream(self.s_in):
            if reuse:
                self.s_in.wait_event(self.ev_free[k])      # the previous replay of this slot has read its inputs
            g.hidden_states.copy_(hidden_states, non_blocking=True)
            g.token_type_ids.copy_(token_type_ids, non_blocking=True)
            g.position_ids.copy_(position_ids, non_blocking=True)
            g.padding_mask.copy_(padding_mask, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        cur.wait_event(self.ev_in[k])
        if reuse:
            cur.wait_event(self.ev_copied[k])              # ... and its output has been copied out
        g.replay()
        self.ev_free[k].record(cur)
        self.ev_done[k].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_done[k])
            self.out_host[k].copy_(g.output, non_blocking=True)
            self.ev_copied[k].record(self.s_out)
        self.count += 1
        return k

    def result(self, slot: int) -> torch.Tensor:
        self.ev_copied[slot].synchronize()
        return self.out_host[slot]
