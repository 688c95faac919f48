"""Stream sharding across the GPUs of one box (one process per GPU) and the gather of per-stream
results.  Streams are independent (no cross-stream state anywhere in resampler.rs / vad.rs), so the
data path needs no collective; the only exchange step is the result gather."""
from __future__ import annotations

import numpy as np


def partition(costs, world: int):
    """Contiguous blocks of stream indices, balanced by cost (input bytes): returns world (lo, hi) pairs."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world <= 0:
        raise ValueError("world must be positive")
    total = costs.sum()
    bounds = [0]
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary whose prefix cost is closest to the target, keep it monotone
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def gather_vad(states, n_frames, group=None):
    """All-gathers per-stream VAD results of every rank.

    states  : torch uint8 [S_local, stride]  (device tensor for NCCL, CPU tensor for gloo)
    n_frames: torch int32 [S_local]
    Every rank must pass the same stride; S_local may differ (ranks are padded to the largest).
    Returns (states [S_total, stride], n_frames [S_total]) in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = states.device
    s_local = torch.tensor([states.shape[0]], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(s_local) for _ in range(world)]
    dist.all_gather(sizes, s_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    s_max = max(sizes)
    stride = states.shape[1]
    pad_states = torch.zeros((s_max, stride), device=dev, dtype=torch.uint8)
    pad_states[: states.shape[0]] = states
    pad_nf = torch.zeros(s_max, device=dev, dtype=torch.int32)
    pad_nf[: states.shape[0]] = n_frames
    out_states = torch.empty((world * s_max, stride), device=dev, dtype=torch.uint8)   # concatenated along dim 0
    out_nf = torch.empty(world * s_max, device=dev, dtype=torch.int32)
    dist.all_gather_into_tensor(out_states, pad_states, group=group)
    dist.all_gather_into_tensor(out_nf, pad_nf, group=group)
    out_states = out_states.view(world, s_max, stride)
    out_nf = out_nf.view(world, s_max)
    st = torch.cat([out_states[r, : sizes[r]] for r in range(world)])
    nf = torch.cat([out_nf[r, : sizes[r]] for r in range(world)])
    return st, nf
