"""Routing plan: the device-side index lists K1 (vex_partition) derives from ``token_type_ids`` and
``padding_mask``.  Routing is layer-invariant, so the plan is computed once per forward and shared by
all 32 layers (the reference recomputes the masks twice per layer and runs 16 ``nonzero`` host syncs
per layer -- modeling_cogvlm.py:94, :239 and every ``x[mask]``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib, ops


@dataclass
class RoutingPlan:
    batch: int
    seq_len: int
    sorted_to_flat: torch.Tensor   # int32 [B*L]
    flat_to_sorted: torch.Tensor   # int32 [B*L]
    sorted_to_token: torch.Tensor  # int32 [B*L]
    token_to_sorted: torch.Tensor  # int32 [B*L]
    token_to_flat: torch.Tensor    # int32 [B*L]
    cu_seqlens: torch.Tensor       # int32 [B+1]
    counts: torch.Tensor           # int32 [4]: Tv, Tl, T, max len  (device resident; never synced by the layer)

    @property
    def rows_cap(self) -> int:
        return self.batch * self.seq_len

    @property
    def n_valid(self) -> torch.Tensor:  # device scalar view, for kernels that take a row-count pointer
        return self.counts[_lib.COUNT_VALID:_lib.COUNT_VALID + 1]

    # host-side views for tests (these DO sync)
    def vision_idx(self) -> torch.Tensor:
        tv = int(self.counts[_lib.COUNT_VISION])
        return self.sorted_to_flat[:tv].long()

    def language_idx(self) -> torch.Tensor:
        tv, tl = int(self.counts[_lib.COUNT_VISION]), int(self.counts[_lib.COUNT_LANGUAGE])
        return self.sorted_to_flat[tv:tv + tl].long()

    def valid_idx(self) -> torch.Tensor:
        return self.token_to_flat[: int(self.counts[_lib.COUNT_VALID])].long()


def build_plan(token_type_ids: torch.Tensor, padding_mask: torch.Tensor) -> RoutingPlan:
    if token_type_ids.dim() != 2 or token_type_ids.shape != padding_mask.shape:
        raise ValueError("token_type_ids and padding_mask must both be [B, L]")
    if not token_type_ids.is_cuda:
        raise ValueError("token_type_ids must be a CUDA tensor (libvex has no CPU path)")
    if token_type_ids.dtype != torch.int64:
        token_type_ids = token_type_ids.long()      # the reference calls .long() on the ids as well (:530)
    if padding_mask.dtype != torch.bool:
        padding_mask = padding_mask.bool()          # padding_mask = attention_mask.bool() (:539)
    token_type_ids, padding_mask = token_type_ids.contiguous(), padding_mask.contiguous()
    B, L = token_type_ids.shape
    if L < 2:
        raise NotImplementedError("q_len == 1 (decode) is outside the prefill hot path (SURVEY 8(f)-2)")
    dev = token_type_ids.device
    idx = torch.empty(5, B * L, dtype=torch.int32, device=dev)
    cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
    counts = torch.empty(_lib.NUM_COUNTS, dtype=torch.int32, device=dev)
    scratch = torch.empty(4 * B, dtype=torch.int32, device=dev)
    ops.partition(token_type_ids, padding_mask, idx[0], idx[1], idx[2], idx[3], idx[4], cu, counts, scratch)
    return RoutingPlan(B, L, idx[0], idx[1], idx[2], idx[3], idx[4], cu, counts)


def _version_of(t: torch.Tensor):
    """In-place-mutation stamp of a key tensor.  Inference tensors (created under ``torch.inference_mode()``, which
    is how the reference's evaluation drivers call ``generate`` -- scripts/evaluate/models/mmmm.py:132) do not track a
    version counter and raise when ``_version`` is read; they are keyed on identity + storage + shape instead."""
    if t.is_inference():
        return ("inference", t.data_ptr(), tuple(t.shape))
    return t._version


class PlanCache:
    """One-entry cache keyed on tensor identity + version counter (the caller loop passes the same
    ``token_type_ids`` / ``padding_mask`` objects to every layer, modeling_cogvlm.py:547-562).  Holding
    references to the key tensors keeps their storage alive, so a recycled data_ptr cannot alias."""

    def __init__(self):
        self._key = None
        self._plan: Optional[RoutingPlan] = None

    def get(self, token_type_ids: torch.Tensor, padding_mask: torch.Tensor) -> RoutingPlan:
        k = self._key
        if (k is not None and k[0] is token_type_ids and k[1] is padding_mask
                and k[2] == _version_of(token_type_ids) and k[3] == _version_of(padding_mask)):
            return self._plan
        plan = build_plan(token_type_ids, padding_mask)
        self._key = (token_type_ids, padding_mask, _version_of(token_type_ids), _version_of(padding_mask))
        self._plan = plan
        return plan

    def clear(self):
        self._key = self._plan = None


GLOBAL_PLAN_CACHE = PlanCache()
