"""Pre-allocated KV cache and the graphed decode step.

The reference keeps the cache as per-layer ``(k, v)`` tuples of [B, heads, L, 128] that ``torch.cat`` re-allocates on
every generated token (/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py:258-262) and drives generation through HF
``generate`` (mmmm/models/mmmm.py:354-406).  Here the cache is one buffer per layer with the reference's layout and
spare capacity, [B, heads, capacity, 128]:

* prefill: the QKV epilogue (VEX_EPI_ROPE) writes post-rotary K and V straight into it (a9);
* decode step: the same epilogue appends the new token's K / V at position ``past_len`` -- a DEVICE counter -- and
  K4d attends to positions [0, past_len]; the counter is advanced on the stream at the end of the step.

Nothing in a step depends on host state, so the whole 32-layer step (7 launches per layer + final norm + counter) is
captured once into a CUDA graph and replayed per token.  The attention mask over the cache is a static [B, capacity]
bool buffer: the prefill's padding mask followed by ones (every generated position is valid, which is what HF
``_update_model_kwargs_for_generation`` appends).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

HEAD_DIM = 128


def next_position_ids(position_ids: torch.Tensor, input_ids: torch.Tensor, bop_token_id: int,
                      eop_token_id: int) -> torch.Tensor:
    """Position of the token a decode step is about to process, [B, 1].

    The reference appends ``position_ids[:, -1:] + 1`` after every step (mmmm/models/mmmm.py:354-365) and then takes one
    back when the token BEFORE the new one is ``<bop>`` or the new token itself is ``<eop>`` (``keep_position``,
    :380-384): the phrase brackets share the position of their neighbour.  ``position_ids`` are the positions used so
    far ([B, n], only the last column is read), ``input_ids`` the ids generated so far INCLUDING the token this step
    feeds ([B, >= 2])."""
    keep = (input_ids[:, -2] == bop_token_id) | (input_ids[:, -1] == eop_token_id)
    return (position_ids[:, -1] + 1 - keep.to(position_ids.dtype)).unsqueeze(1)


class StaticKVCache:
    def __init__(self, n_layers: int, batch: int, heads: int, capacity: int, device, dtype=torch.bfloat16):
        self.batch, self.heads, self.capacity = batch, heads, capacity
        kv = torch.empty(n_layers, 2, batch, heads, capacity, HEAD_DIM, dtype=dtype, device=device)
        self._kv = kv
        self.layers = [(kv[i, 0], kv[i, 1]) for i in range(n_layers)]
        self.past_len = torch.zeros(1, dtype=torch.int32, device=device)   # device counter: positions cached
        self.mask = torch.ones(batch, capacity, dtype=torch.bool, device=device)
        self.host_len = 0                                                   # host mirror (bounds checks only)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._static = None

    def start(self, padding_mask: torch.Tensor) -> None:
        """After the prefill wrote positions [0, L): remember L and the prefill's padding mask."""
        L = padding_mask.shape[1]
        self.mask.fill_(True)
        self.mask[:, :L].copy_(padding_mask)
        self.past_len.fill_(L)
        self.host_len = L

    def views(self):
        """The reference's tuple-cache view of the current contents: per layer (k, v) [B, heads, L, 128]."""
        L = self.host_len
        return tuple((k[:, :, :L], v[:, :, :L]) for k, v in self.layers)

    def reorder(self, beam_idx: torch.Tensor) -> None:
        """Beam search: sample b continues beam ``beam_idx[b]`` -- the reference's ``_reorder_cache``
        (modeling_cogvlm.py:782-788: ``index_select(0, beam_idx)`` on every cached k / v), done in place so that the
        buffers a captured decode graph points at stay where they are; the mask rows move with their samples."""
        idx = beam_idx.to(self._kv.device).long()
        L = self.host_len
        self._kv[:, :, :, :, :L].copy_(self._kv[:, :, :, :, :L].index_select(2, idx))
        self.mask.copy_(self.mask.index_select(0, idx))

    # ------------------------------------------------------------------------------------------------------
    def _run(self, model, hidden: torch.Tensor, position_ids: torch.Tensor) -> torch.Tensor:
        from .modeling_cogvlm import decode_core, resolve_norm
        B, _, H = hidden.shape
        h = hidden.view(B, H)
        for layer, (k, v) in zip(model.layers, self.layers):
            h = decode_core(layer, h, position_ids, k, v, self.mask, self.past_len)
        norm = resolve_norm(model.norm)
        out = torch.empty_like(h)
        n = torch.full((1,), B, dtype=torch.int32, device=h.device) if self._n_rows is None else self._n_rows
        ops.rmsnorm_gather(h, norm.weight.detach(), norm.variance_epsilon, None, n, out)   # L == 1: plain norm (:572-573)
        ops.advance_counter(self.past_len, 1)
        return out.view(B, 1, H)

    _n_rows = None

    def step(self, model, inputs_embeds: torch.Tensor, position_ids: torch.Tensor, graph: bool = True) -> torch.Tensor:
        if self.host_len >= self.capacity:
            raise RuntimeError(f"KV cache is full ({self.capacity} positions): allocate it with a larger max_new_tokens")
        B = self.batch
        if inputs_embeds.shape[:2] != (B, 1) or inputs_embeds.dtype != torch.bfloat16 or not inputs_embeds.is_cuda:
            raise ValueError("decode_step expects CUDA bfloat16 inputs_embeds [B, 1, H]")
        if self._n_rows is None:
            self._n_rows = torch.full((1,), B, dtype=torch.int32, device=inputs_embeds.device)
        position_ids = position_ids.reshape(B, 1).long()
        with torch.no_grad():
            if not graph:
                out = self._run(model, inputs_embeds.contiguous(), position_ids.contiguous())
            else:
                if self._graph is None:
                    self._capture(model, inputs_embeds, position_ids)
                self._static[0].copy_(inputs_embeds, non_blocking=True)
                self._static[1].copy_(position_ids, non_blocking=True)
                self._graph.replay()
                out = self._static[2]
        self.host_len += 1
        return out

    def _capture(self, model, inputs_embeds: torch.Tensor, position_ids: torch.Tensor) -> None:
        h, p = inputs_embeds.clone().contiguous(), position_ids.clone().contiguous()
        saved = self.past_len.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # first-call work (function attributes, rotary tables, bf16 adapter copies) outside capture
                self._run(model, h, p)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._run(model, h, p)
        # warm-up and capture runs appended garbage at positions >= past_len and advanced the counter: rewind (the
        # positions are overwritten by the real steps before they are ever attended to)
        self.past_len.copy_(saved)
        self._graph, self._static = g, (h, p, out)
