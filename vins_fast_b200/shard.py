"""Multi-GPU partition of the front-end: independent stereo streams, one process per GPU, no data-path collective.

A single frame does not shard (one image pair, <= max_cnt features, and frame k+1 consumes the state of frame k), but
streams do (BASELINE config 4): stream s of S runs on rank s mod G and its state never leaves that
GPU.  torch.distributed is used only around the data path: the start/stop barrier, the max-over-ranks of the device
time, and gathering per-stream result summaries on rank 0 (gloo on CPU in the tests, NCCL on the GPU box).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Stream s -> rank s mod world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_streams, world))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of one scalar (device milliseconds of the timed region)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stream_results(local: Dict[int, np.ndarray], n_streams: int) -> Dict[int, np.ndarray]:
    """Collects {stream: array} from every rank on rank 0 (others get {}).  Host-side bookkeeping, not the data path."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object({int(k): np.asarray(v) for k, v in local.items()}, parts, dst=0)
    if dist.get_rank() != 0:
        return {}
    out: Dict[int, np.ndarray] = {}
    for p in parts:
        out.update(p)
    if sorted(out) != list(range(n_streams)):
        raise RuntimeError(f"streams gathered {sorted(out)} != 0..{n_streams - 1}")
    return out
