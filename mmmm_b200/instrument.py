"""Launch accounting and per-kernel CUDA-event timing for bench.py / profiling.

Every libvex call in ``ops`` passes through ``region(name, n_kernels)``: it counts the kernels
launched (bench.py's ``gpu_launches``) and, while ``profile`` is active, brackets the call with
CUDA events on the launching stream."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Callable, Dict, List, Optional

import torch

_launches = 0
_events: Optional[List] = None


def reset() -> None:
    global _launches
    _launches = 0


def launches() -> int:
    return _launches


@contextmanager
def region(name: str, n_kernels: int = 1):
    global _launches
    _launches += n_kernels
    if _events is None:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    _events.append((name, e0, e1))


def profile(step: Callable[[], object], iters: int = 5) -> Dict[str, dict]:
    """Average device time per libvex call name over ``iters`` steps (events serialise nothing: all
    launches are already stream-ordered)."""
    global _events
    step()
    torch.cuda.synchronize()
    _events = []
    try:
        for _ in range(iters):
            step()
        torch.cuda.synchronize()
        acc: Dict[str, List[float]] = {}
        for name, e0, e1 in _events:
            acc.setdefault(name, []).append(e0.elapsed_time(e1))
    finally:
        _events = None
    return {k: {"ms": sum(v) / iters, "calls_per_step": len(v) / iters} for k, v in acc.items()}
