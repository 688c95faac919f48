"""world_size-2 gloo tests (CPU) of the N > 1 host logic: byte-balanced stream partition and the
result gather.  The kernels themselves need no collective (streams are independent)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_balances_bytes():
    from audioflow import shard
    costs = [1440000 * 4 if i % 2 == 0 else 1323000 * 4 for i in range(4096)]
    for world in (1, 2, 4, 8):
        parts = shard.partition(costs, world)
        assert parts[0][0] == 0 and parts[-1][1] == 4096
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        loads = [sum(costs[a:b]) for a, b in parts]
        assert max(loads) / (sum(costs) / world) < 1.002
    # ragged: one huge stream
    parts = shard.partition([100, 1, 1, 1, 1, 1], 2)
    assert parts == [(0, 1), (1, 6)]
    assert shard.partition([], 2) == [(0, 0), (0, 0)]
    assert shard.partition([5, 5], 4)[-1][1] == 2


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "audio-flow-rs_b200"))
    from audioflow import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [10, 30, 20, 20, 10, 10, 40]
    lo, hi = shard.partition(costs, world)[rank]
    # fake per-stream results: stream i has i + 1 frames of state (i % 3)
    states = torch.zeros((hi - lo, 16), dtype=torch.uint8)
    nf = torch.zeros(hi - lo, dtype=torch.int32)
    for j, i in enumerate(range(lo, hi)):
        states[j, : i + 1] = i % 3
        nf[j] = i + 1
    st, n = shard.gather_vad(states, nf)
    # the planned exchange (sizes and counts once, one collective per step) must deliver the same thing, repeatedly
    g = shard.VadGather(hi - lo, 16, nf, torch.device("cpu"))
    for rep in range(2):
        st2 = g.run(states)
        assert torch.equal(st2, st) and torch.equal(g.n_frames, n)
    q.put((rank, (lo, hi), st.numpy().copy(), n.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, part, st, n in res:
        assert st.shape == (7, 16) and n.tolist() == [1, 2, 3, 4, 5, 6, 7]     # every rank sees all streams, in order
        for i in range(7):
            assert (st[i, : i + 1] == i % 3).all() and (st[i, i + 1:] == 0).all()
    parts = sorted(r[1] for r in res)
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == 7


def test_c_abi_partition_matches_helper_and_balances():
    """af_shard_partition (pure host code behind the C ABI: what af_sharded_batch_create uses) against the Python helper:
    contiguous, covering, byte-balanced (44.1 kHz and 48 kHz streams differ by 8 %), i16 streams cost half."""
    import audioflow as af
    from audioflow import shard
    descs = [(0, 1440000 if i % 2 == 0 else 1323000, 48000 if i % 2 == 0 else 44100, 1, af.AF_FMT_F32) for i in range(4096)]
    costs = [d[1] * 4 for d in descs]
    for world in (1, 2, 3, 4, 8, 16):
        parts = af.shard_partition(descs, world)
        assert parts == shard.partition(costs, world)
        assert parts[0][0] == 0 and parts[-1][1] == 4096 and all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        loads = [sum(costs[a:b]) for a, b in parts]
        assert max(loads) / (sum(costs) / world) < 1.002
    mixed = [(0, 1000, 48000, 1, af.AF_FMT_I16)] * 4 + [(0, 1000, 48000, 1, af.AF_FMT_F32)] * 2      # 4 x 2000 B, 2 x 4000 B
    assert af.shard_partition(mixed, 2) == [(0, 4), (4, 6)]
    assert af.shard_partition([(0, 100, 1, 1, 0)] + [(0, 1, 1, 1, 0)] * 5, 2) == [(0, 1), (1, 6)]
    assert af.shard_partition([], 2) == [(0, 0), (0, 0)]
