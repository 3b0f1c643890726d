"""Sample sharding for the data-parallel path.

Every op in the layer is per-token or per-sample (attention never crosses samples: block-diagonal mask,
/root/reference/mmmm/models/cogvlm/modeling_cogvlm.py:117-128), so a batch shards by sample with no data-path
collective -- what the reference does with DDP + DistributedSamplerWrapper (mmmm/data/datamodule.py:104-111).
Rank g of G takes the contiguous samples [g*B/G, (g+1)*B/G); remainders go to the
first ranks.  The only cross-rank traffic is the max-over-ranks of the timing."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int):
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return tuple(t[lo:hi] for t in tensors)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (identity when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
