"""Per-role wait statistics (stats build, AF_GPU_LIB=.../libaudioflow_gpu_stats.so) of cfg2 with and without the PCM write-out:
are the resampler warps' global stores what their barrier arrives wait for?  Measured: no -- the shares are the same and the
kernel is only 2 % faster without the stores."""
import os, sys, ctypes as C
sys.path.insert(0, 'audio-flow-rs_b200'); sys.path.insert(0, '.')
import torch, numpy as np
import audioflow as af
from audioflow import synth
import bench
af.init(0)
L = af.load_library()
dev = torch.device('cuda', 0)
S, n = 256, 30 * 48000
x = synth.torch_batch(S, 30.0, 48000, 1, dev, seed=0)
for wp in (True, False):
    pipe = af.Pipeline(af.pipeline_config(n_mels=80, vad_enable=False, write_pcm=wp))
    descs = [(x[i].data_ptr(), n, 48000, 1, af.AF_FMT_F32) for i in range(S)]
    b = pipe.batch(descs, af.AF_MEM_DEVICE)
    pcm = torch.empty((S, b.pcm_stride), device=dev)
    lm = torch.empty((S, b.logmel_stride), device=dev)
    o = b.outputs_struct(pcm.data_ptr() if wp else 0, b.pcm_stride, lm.data_ptr(), b.logmel_stride)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): b.run_device(o, 0)
    torch.cuda.synchronize()
    bench.pipe_stats_clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): b.run_device(o, 0)
    e1.record(); torch.cuda.synchronize()
    print('write_pcm', wp, 'ms/step', e0.elapsed_time(e1) / 20, file=sys.stderr)
    bench.pipe_stats_print()
