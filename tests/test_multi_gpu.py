"""Multi-GPU behind the C ABI: stream sharding + the NCCL gather of VAD states inside the library.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped on a 1-GPU box.

* ONE process driving 2 GPUs (af_init_multi -> ncclCommInitAll): device-resident shards + gather, and the host-buffer form;
* one process per GPU (af_comm_unique_id / af_comm_init_rank -> ncclCommInitRank), world size 2, through ctypes only --
  no torch.distributed anywhere on the path."""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from test_parity_gpu import assert_bit_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import audioflow
    n = C.c_int(0)
    audioflow.load_library().af_device_count(C.byref(n))
    return n.value


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")

CASES = [(900 + i, 2.0 + 0.37 * (i % 5), 44100 if i % 2 else 48000) for i in range(14)]     # mixed rates, ragged lengths


def _streams():
    from audioflow import synth
    return [(synth.stream(sid, sec, rate, 1), rate) for (sid, sec, rate) in CASES]


def _oracle(orc, x, rate):
    return orc.pipeline_stream(x, 1, rate, None, orc.default_vad_config(), 400, 160, "f32")


@needs2
def test_single_process_two_gpus_sharded_batch_with_gather(af, orc):
    import torch
    xs = _streams()
    af.init_multi(2)
    try:
        geo = [(0, len(x), rate, 1, af.AF_FMT_F32) for (x, rate) in xs]
        parts = af.shard_partition(geo, 2)
        assert parts[0][1] == parts[1][0] and 0 < parts[0][1] < len(xs)
        # every stream's input goes to the GPU of the rank that owns it
        owner = [0 if i < parts[0][1] else 1 for i in range(len(xs))]
        dx = [torch.tensor(x, device=f"cuda:{owner[i]}") for i, (x, _) in enumerate(xs)]
        descs = [(dx[i].data_ptr(), len(x), rate, 1, af.AF_FMT_F32) for i, (x, rate) in enumerate(xs)]
        pipe = af.Pipeline(af.pipeline_config(n_mels=80))
        sb = af.ShardedBatch(pipe, descs, af.AF_MEM_DEVICE)
        outs = af.ShardedOutputsC()
        bufs = []
        for r in range(2):
            lo, cnt, dev = sb.shard(r)
            assert (lo, lo + cnt) == parts[r] and dev == r
            b = sb.local(r)
            d = torch.device("cuda", r)
            pcm = torch.zeros((cnt, b.pcm_stride), device=d)
            lm = torch.zeros((cnt, b.logmel_stride), device=d)
            fin = torch.zeros((cnt, 6), device=d, dtype=torch.int32)
            outs.shard[r] = af.OutputsC(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, None, 0, None, 0, fin.data_ptr())
            bufs.append((b, pcm, lm, fin))
        for d in range(2):
            torch.cuda.synchronize(d)
        for rep in range(3):                       # double-buffered gather: three runs reuse the first buffer
            sb.run(outs, gather=True, wait=(rep != 1))
        sb.wait()
        assert sb.gather_ms(0) >= 0.0
        refs = [_oracle(orc, x, rate) for (x, rate) in xs]
        for r in range(2):
            ptr, stride, rows, nv = sb.gathered(r)
            assert rows == max(p[1] - p[0] for p in parts)
            assert nv.tolist() == [len(ref["vad"]) for ref in refs]
            g = sb.gathered_host(r)
            for i, ref in enumerate(refs):          # EVERY GPU holds the states of EVERY stream
                assert_bit_equal(g[i, :nv[i]], ref["vad"], f"gathered on GPU {r}: stream {i}")
            b, pcm, lm, fin = bufs[r]
            for li in range(parts[r][1] - parts[r][0]):
                i = parts[r][0] + li
                assert_bit_equal(pcm[li, :int(b.n_out[li])].cpu().numpy(), refs[i]["pcm"], f"pcm {i}")
                assert int(fin[li, 1]) == refs[i]["vad_final"]["state"] and int(fin[li, 4]) == refs[i]["vad_final"]["speech_frames"]
        del sb

        # host-buffer form: one call, rows of ALL streams in host memory, each GPU fed by its own host thread
        hdescs = [(x.ctypes.data, len(x), rate, 1, af.AF_FMT_F32) for (x, rate) in xs]
        hb = af.ShardedBatch(pipe, hdescs, af.AF_MEM_HOST)
        loc = [hb.local(r) for r in range(2)]
        pcm_stride = max(b.pcm_stride for b in loc); vad_stride = max(b.vad_stride for b in loc)
        h_pcm = np.zeros((len(xs), pcm_stride), np.float32)
        h_vad = np.zeros((len(xs), vad_stride), np.uint8)
        hb.run_host(af.OutputsC(h_pcm.ctypes.data, pcm_stride, None, 0, h_vad.ctypes.data, vad_stride, None, 0, None))
        for i, ref in enumerate(refs):
            assert_bit_equal(h_pcm[i, :len(ref["pcm"])], ref["pcm"], f"host pcm {i}")
            assert_bit_equal(h_vad[i, :len(ref["vad"])], ref["vad"], f"host vad {i}")
        del hb
    finally:
        af.comm_shutdown()
        af.init(0)


def _rank_worker(rank, world, uid_q, res_q):
    sys.path[:0] = [os.path.join(ROOT, "audio-flow-rs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import torch          # (used for device memory only) BEFORE the library binds NCCL: torch must find its own bundled libnccl.so.2
    import audioflow as af
    try:
        af.init(rank)                                  # one process per GPU
        if rank == 0:
            uid = af.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        af.comm_init_rank(world, rank, uid)
        xs = _streams()
        geo = [(0, len(x), rate, 1, af.AF_FMT_F32) for (x, rate) in xs]
        lo, hi = af.shard_partition(geo, world)[rank]
        L = af.load_library()
        d = torch.device("cuda", rank)                 # device copies of this rank's streams only
        dx = {i: torch.tensor(xs[i][0], device=d) for i in range(lo, hi)}
        descs = [(dx[i].data_ptr() if lo <= i < hi else 0, len(x), rate, 1, af.AF_FMT_F32) for i, (x, rate) in enumerate(xs)]
        pipe = af.Pipeline(af.pipeline_config(n_mels=0))
        sb = af.ShardedBatch(pipe, descs, af.AF_MEM_DEVICE)
        b = sb.local(rank)
        assert sb.local(1 - rank) is None
        pcm = torch.zeros((hi - lo, b.pcm_stride), device=d)
        outs = af.ShardedOutputsC()
        outs.shard[rank] = af.OutputsC(pcm.data_ptr(), b.pcm_stride, None, 0, None, 0, None, 0, None)
        torch.cuda.synchronize()
        for _ in range(2):
            sb.run(outs, gather=True, wait=True)
        ptr, stride, rows, nv = sb.gathered(rank)
        res_q.put((rank, (lo, hi), sb.gathered_host(rank), nv.copy(), sb.gather_ms(rank)))
        del sb
        af.comm_shutdown()
    except Exception as e:          # pragma: no cover
        res_q.put((rank, "error", repr(e), None, None))
        raise


@needs2
def test_one_process_per_gpu_world2_through_ctypes(orc):
    """World size 2, one process per GPU, communicator formed through the C ABI (unique id shipped over a queue)."""
    ctx = mp.get_context("spawn")
    uid_q, res_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, uid_q, res_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [res_q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    xs = _streams()
    refs = [_oracle(orc, x, rate) for (x, rate) in xs]
    parts = {r[0]: r[1] for r in res}
    assert "error" not in parts.values(), res
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == len(xs)
    for rank, part, g, nv, ms in res:
        assert nv.tolist() == [len(ref["vad"]) for ref in refs]
        for i, ref in enumerate(refs):
            assert_bit_equal(g[i, :nv[i]], ref["vad"], f"rank {rank} sees stream {i}")


def test_c_host_drives_every_gpu_through_the_c_abi_only():
    """host/multi_gpu_host.cpp links nothing but libaudioflow_gpu.so: af_init_multi over every GPU of the box, one
    af_sharded_batch_run_host call from pinned host buffers, every output byte equal to the single-GPU run."""
    import subprocess
    exe = os.path.join(ROOT, "audio-flow-rs_b200", "lib", "multi_gpu_host")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe, "0", "24", "3.0"], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatching outputs: 0" in r.stdout and r.stdout.strip().endswith("OK")
