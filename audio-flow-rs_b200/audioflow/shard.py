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


class VadGather:
    """The same exchange planned once per batch: the shard sizes and the (static) frame counts are exchanged at
    construction, the buffers are allocated once, and every step is ONE copy into the padded send buffer plus ONE
    all_gather_into_tensor -- no host synchronisation on the step path.  (gather_vad above pays an .item() sync, three
    collectives and several allocations per call: 4.6 ms per cfg3 step on 2 GPUs against ~0.1 ms here.)"""

    def __init__(self, n_local: int, stride: int, n_frames, device, group=None):
        import torch
        import torch.distributed as dist
        self.group, self.world = group, dist.get_world_size(group)
        s_local = torch.tensor([n_local], device=device, dtype=torch.int64)
        sizes = [torch.zeros_like(s_local) for _ in range(self.world)]
        dist.all_gather(sizes, s_local, group=group)
        self.sizes = [int(s.item()) for s in sizes]
        self.s_max, self.stride, self.n_local = max(self.sizes), stride, n_local
        self.send = torch.zeros((self.s_max, stride), device=device, dtype=torch.uint8)
        self.recv = torch.empty((self.world * self.s_max, stride), device=device, dtype=torch.uint8)
        pad_nf = torch.zeros(self.s_max, device=device, dtype=torch.int32)
        pad_nf[:n_local] = n_frames
        out_nf = torch.empty(self.world * self.s_max, device=device, dtype=torch.int32)
        dist.all_gather_into_tensor(out_nf, pad_nf, group=group)
        out_nf = out_nf.view(self.world, self.s_max)
        self.n_frames = torch.cat([out_nf[r, : self.sizes[r]] for r in range(self.world)])
        self.dense = all(s == self.s_max for s in self.sizes)

    def run(self, states):
        """states: uint8 [n_local, stride] -> [S_total, stride] in rank order (a view of the receive buffer when every
        rank holds the same number of streams)."""
        import torch
        import torch.distributed as dist
        self.send[: self.n_local].copy_(states)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self.dense:
            return self.recv
        r3 = self.recv.view(self.world, self.s_max, self.stride)
        return torch.cat([r3[r, : self.sizes[r]] for r in range(self.world)])
