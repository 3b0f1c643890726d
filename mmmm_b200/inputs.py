"""Synthetic inputs with the shape real VividMed batches have.

Sequence layout of one sample, restated from ``prepare_vlm_inputs``
(/root/reference/mmmm/data/utils.py:104-124) and the collate function
(/root/reference/mmmm/data/datamodule.py:20-39):

    [bos] [boi, Nv image patches, eoi] [grd] [Nt text tokens]        L = 1 + (Nv + 2) + 1 + Nt
    token_type_ids = [0] + [1] * (Nv + 2) + [0] * (1 + Nt)
    position_ids   = [0, 1] + [2] * Nv + [3, 4] + [5, 6, ...]        (all image patches share position 2)
    right padding with token_type 0, position 0, attention_mask 0

No tokenizer is involved (there are no weights for it offline); only the shapes matter on this path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

LANGUAGE_TOKEN_TYPE = 0  # mmmm/data/utils.py:192
VISION_TOKEN_TYPE = 1    # mmmm/data/utils.py:193


@dataclass
class LayerInputs:
    hidden_states: torch.Tensor   # [B, L, H]
    token_type_ids: torch.Tensor  # int64 [B, L]
    position_ids: torch.Tensor    # int64 [B, L]
    padding_mask: torch.Tensor    # bool  [B, L]

    @property
    def num_valid_tokens(self) -> int:
        return int(self.padding_mask.sum())

    def to(self, device, non_blocking: bool = False) -> "LayerInputs":
        return LayerInputs(*(t.to(device, non_blocking=non_blocking) for t in
                             (self.hidden_states, self.token_type_ids, self.position_ids, self.padding_mask)))


def sample_layout(num_vision: int, num_text: int):
    """token_type_ids / position_ids of one unpadded sample."""
    tt = [LANGUAGE_TOKEN_TYPE] + [VISION_TOKEN_TYPE] * (num_vision + 2) + [LANGUAGE_TOKEN_TYPE] * (1 + num_text)
    pos = [0, 1] + [2] * num_vision + [3, 4] + list(range(5, 5 + num_text))
    return torch.tensor(tt, dtype=torch.int64), torch.tensor(pos, dtype=torch.int64)


def make_ids(batch: int, num_vision: int, num_text: int, *, ragged: bool = False, seed: int = 0):
    """Batch of id tensors.  ``ragged``: per-sample text length ~ U{Nt/2 .. Nt} (seeded), right-padded."""
    g = torch.Generator().manual_seed(seed)
    L = 1 + (num_vision + 2) + 1 + num_text
    tt = torch.zeros(batch, L, dtype=torch.int64)
    pos = torch.zeros(batch, L, dtype=torch.int64)
    pm = torch.zeros(batch, L, dtype=torch.bool)
    for b in range(batch):
        nt = num_text
        if ragged:
            lo = max(num_text // 2, 0)
            nt = int(torch.randint(lo, num_text + 1, (1,), generator=g))
        t, p = sample_layout(num_vision, nt)
        n = t.numel()
        tt[b, :n], pos[b, :n], pm[b, :n] = t, p, True
    return tt, pos, pm


def make_inputs(batch: int, num_vision: int, num_text: int, hidden_size: int, *, ragged: bool = False,
                seed: int = 0, dtype=torch.bfloat16, device: Optional[str] = None) -> LayerInputs:
    tt, pos, pm = make_ids(batch, num_vision, num_text, ragged=ragged, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    h = torch.randn(batch, tt.shape[1], hidden_size, generator=g).to(dtype)
    out = LayerInputs(h, tt, pos, pm)
    return out.to(device) if device is not None else out


# BASELINE.json configs; per-GPU sample counts are set by the caller.
CONFIGS = {
    "c1": dict(batch=1, num_vision=1225, num_text=128, layers=2),
    "c2": dict(batch=8, num_vision=1225, num_text=256, layers=1),
    "c3": dict(batch=64, num_vision=1225, num_text=256, layers=32),
    "c4": dict(batch=16, num_vision=2048, num_text=512, layers=32),
}
