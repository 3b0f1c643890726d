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


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs: the PCI device's
    ``numa_node`` and the node's ``cpulist``), so that pinned host buffers allocated afterwards are node-local
    (first-touch / local allocation policy) and H2D / D2H copies do not cross the socket interconnect.  With one
    process per GPU (the reference's DDP launch, conf/phase-vlm/fit.yaml:11-15) eight ranks otherwise share whatever
    node their pages landed on.  Best effort: returns what it did, never raises."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        prop = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:  # sysfs not mounted, attribute missing, ...
        info["error"] = f"{type(e).__name__}: {e}"
    return info
