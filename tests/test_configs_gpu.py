"""GPU parity tests at the FULL size of the BASELINE.json configs that round 1 only covered reduced:

* cfg4 -- one 3600 s 48 kHz stereo f32 recording -> 16 kHz PCM + 128-bin log-mel + VAD + segmentation
  (quarter-staged instance of the fused kernel + the many-CTA scan + the segmentation kernel together);
* cfg5 -- streaming sessions with S = 17, 65 and 1024 concurrent streams, >= 50 ticks of 20 ms: packed ticks that cross
  the 32-frame step and the 128-frame tile of the fused kernel;
* the measured log-mel error on cfg2 and cfg4 (how often and how far the f32 transform leaves the 1e-4 band).

Bit-exact against the oracle for PCM, energies, VAD states, final detector state and segments; log-mel within the
tolerance stated in test_parity_gpu.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from test_parity_gpu import LOGMEL_ABS, assert_bit_equal, assert_logmel_close, logmel_tolerance

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REPORT = os.path.join(os.path.dirname(HERE), "gpurun_out", "logmel_error_report.json")


def logmel_error_stats(got, ref):
    """max |err|, share of bins beyond 1e-4 and relative L2 of an f32 log-mel block against the f64 oracle."""
    g, r = got.astype(np.float64), ref.astype(np.float64)
    err = np.abs(g - r)
    dr_db = (r.max(axis=1, keepdims=True) - r) * (10.0 / np.log(10.0))          # dB below the frame's strongest mel bin
    near = dr_db <= 50.0
    return {"logmel_max_abs": float(err.max()), "frac_bins_over_1e-4": float((err > LOGMEL_ABS).mean()),
            "rel_l2": float(np.sqrt((err ** 2).sum() / max((r ** 2).sum(), 1e-30))), "bins": int(err.size),
            "max_abs_within_50dB_of_frame_peak": float(err[near].max()) if near.any() else 0.0,
            "frac_bins_within_50dB": float(near.mean()),
            "max_ratio_to_stated_bound": float((err / logmel_tolerance(ref)).max()),
            "min_dB_below_peak_of_bins_over_1e-4": float(dr_db[err > LOGMEL_ABS].min()) if (err > LOGMEL_ABS).any() else None}


def _report(key, stats):
    print(f"[logmel-error] {key}: {json.dumps(stats)}")
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        cur = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        cur[key] = stats
        json.dump(cur, open(REPORT, "w"), indent=1)
    except OSError:
        pass


# =============================================================================================
# cfg4 at full size
# =============================================================================================
def test_cfg4_full_hour_stereo_128mel_vad_segments(af, orc):
    import torch
    from audioflow import synth
    dev = torch.device("cuda")
    sec, rate, ch, M = 3600.0, 48000, 2, 128
    x = synth.torch_batch(1, sec, rate, ch, dev, seed=404)
    pipe = af.Pipeline(af.pipeline_config(n_mels=M))
    b = pipe.batch([(x[0].data_ptr(), x.shape[1], rate, ch, af.AF_FMT_F32)], af.AF_MEM_DEVICE)
    n_out, T = int(b.n_out[0]), int(b.n_feat[0])
    assert n_out == 57599998 and T == 359998 and int(b.n_vad[0]) == T
    pcm = torch.full((1, b.pcm_stride), float("nan"), device=dev)
    lm = torch.full((1, b.logmel_stride), float("nan"), device=dev)
    vad = torch.full((1, b.vad_stride), 255, device=dev, dtype=torch.uint8)
    en = torch.zeros((1, b.energy_stride), device=dev)
    fin = torch.zeros((1, 6), device=dev, dtype=torch.int32)
    o = b.outputs_struct(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, vad.data_ptr(), b.vad_stride,
                         en.data_ptr(), b.energy_stride, fin.data_ptr())
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    b.run_device(o)
    seg_cap = 1 << 16
    seg = torch.zeros((1, seg_cap, 2), device=dev, dtype=torch.int32)
    nseg = torch.zeros(1, device=dev, dtype=torch.int32)
    nfr = torch.tensor(b.n_vad[:1].astype(np.int32), device=dev)
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    assert af.load_library().af_vad_segments(vad.data_ptr(), b.vad_stride, nfr.data_ptr(), 1, seg.data_ptr(), seg_cap,
                                             nseg.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert not torch.isnan(lm[0, :T * M]).any() and torch.isnan(lm[0, T * M:]).all()
    assert torch.isnan(pcm[0, n_out:]).all() and bool((vad[0, T:] == 255).all())

    # ---- the oracle over the whole hour: downmix, BatchResampler(all) + flush, energies, detector ----
    xh = x[0].cpu().numpy()
    ref_pcm = orc.resample_stream(orc.to_mono(xh, ch), rate)
    del xh
    got_pcm = pcm[0, :n_out].cpu().numpy()
    assert_bit_equal(got_pcm, ref_pcm, "cfg4 pcm (57.6 M samples)")
    det = orc.VoiceActivityDetector()
    ref_vad, ref_en = det.stream(ref_pcm, 400, 160)
    assert len(ref_vad) == T
    assert_bit_equal(en[0, :T].cpu().numpy(), ref_en, "cfg4 energies")
    got_vad = vad[0, :T].cpu().numpy()
    assert_bit_equal(got_vad, ref_vad, "cfg4 vad states")
    assert len(np.unique(ref_vad)) == 3, "the synthetic hour must visit Silence, Speech and Ending"
    assert int(fin[0, 1]) == det.state()
    assert int(fin[0, 4]) == det.speech_frame_count()
    assert int(fin[0, 0]) == int(np.float32(det.smoothed_energy()).view(np.int32))
    exp_seg = orc.vad_segments(ref_vad)
    assert 100 < len(exp_seg) < seg_cap
    assert int(nseg[0]) == len(exp_seg)
    assert np.array_equal(seg[0, :len(exp_seg)].cpu().numpy().astype(np.uint32), exp_seg)

    # ---- gated output (f1): speech-only PCM + log-mel rows, packed, against the oracle's compaction of the GPU rows ----
    n_s = len(exp_seg)
    kept = int((exp_seg[:, 1] - exp_seg[:, 0]).sum())
    g_pcm = torch.full((1, kept * 160 + 16), float("nan"), device=dev)
    g_lm = torch.full((1, kept * M + 16), float("nan"), device=dev)
    g_off = torch.zeros((1, seg_cap + 1), device=dev, dtype=torch.int32)
    g_n = torch.zeros(1, device=dev, dtype=torch.int32)
    n_out_d = torch.tensor([n_out], device=dev, dtype=torch.int32)
    go = af.GateOutputsC(g_pcm.data_ptr(), g_pcm.shape[1], g_lm.data_ptr(), g_lm.shape[1], g_off.data_ptr(), g_n.data_ptr())
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    af._check(af.load_library().af_vad_gate(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, M, n_out_d.data_ptr(), 160,
                                            seg.data_ptr(), seg_cap, nseg.data_ptr(), 1, C.byref(go), None))
    torch.cuda.synchronize()
    assert int(g_n[0]) == kept and 0 < kept < T
    lm_rows = lm[0, :T * M].reshape(T, M).cpu().numpy()
    ref_gp, ref_gl, ref_off = orc.vad_gate(got_pcm, lm_rows, exp_seg, 160)
    assert np.array_equal(g_off[0, :n_s + 1].cpu().numpy().astype(np.uint32), ref_off)
    assert_bit_equal(g_pcm[0, :kept * 160].cpu().numpy(), ref_gp, "cfg4 gated pcm")
    assert np.array_equal(g_lm[0, :kept * M].cpu().numpy().view(np.uint32), ref_gl.reshape(-1).view(np.uint32)), "cfg4 gated logmel"
    assert torch.isnan(g_pcm[0, kept * 160:]).all() and torch.isnan(g_lm[0, kept * M:]).all()
    del lm_rows, ref_gp, ref_gl

    # ---- log-mel against the f64 oracle on 60 s windows: the start, one straddling the 8192-frame scan block /
    #      64-tile boundary, the middle, and the end of the recording (incl. the last, partial tile) ----
    fc = orc.default_feat_config(M)
    G, R = [], []
    for name, f0 in (("start", 0), ("straddle-8192", 8192 - 3000), ("middle", 180000 - 17), ("end", T - 6000)):
        nf = 6000
        y = ref_pcm[f0 * 160:(f0 + nf - 1) * 160 + 400]
        ref_lm = orc.logmel(y, fc)
        got_lm = lm[0, f0 * M:(f0 + nf) * M].reshape(nf, M).cpu().numpy()
        assert_logmel_close(got_lm, ref_lm, f"cfg4 logmel window {name}")
        G.append(got_lm); R.append(ref_lm)
    _report("cfg4_4x60s_windows_128mel", logmel_error_stats(np.concatenate(G), np.concatenate(R)))


def test_cfg2_logmel_error_report(af, orc):
    """cfg2 (48 kHz mono, 80 mel): the measured log-mel error of 8 whole 30 s streams -- quantifies the tolerance clause."""
    from audioflow import synth
    streams = [(synth.stream(300 + i, 30.0, 48000, 1), 48000, 1) for i in range(8)]
    got = af.Pipeline(af.pipeline_config(n_mels=80, vad_enable=False)).run_host(streams)
    fc = orc.default_feat_config(80)
    G, R = [], []
    for (x, rate, ch), g in zip(streams, got):
        ref_pcm = orc.resample_stream(x, rate)
        assert_bit_equal(g["pcm"], ref_pcm, "cfg2 pcm")
        ref_lm = orc.logmel(ref_pcm, fc)
        assert_logmel_close(g["logmel"], ref_lm, "cfg2 logmel")
        G.append(g["logmel"]); R.append(ref_lm)
    st = logmel_error_stats(np.concatenate(G), np.concatenate(R))
    _report("cfg2_8x30s_80mel", st)
    assert st["frac_bins_over_1e-4"] < 1e-3 and st["rel_l2"] <= 1e-5
    # the same frames through numpy's own float32 FFT (pocketfft, no GPU code): the band is a property of f32 transforms
    import scipy.fft
    win = orc.hann_window(400).astype(np.float32)
    fb = orc.mel_filterbank(fc).astype(np.float64)
    N32 = []
    for (x, rate, ch), ref_lm in list(zip(streams, R))[:2]:
        y = orc.resample_stream(x, rate)
        T = ref_lm.shape[0]
        fr = np.lib.stride_tricks.sliding_window_view(y, 400)[::160][:T] * win
        X = scipy.fft.rfft(fr.astype(np.float32), n=512, axis=1)
        assert X.dtype == np.complex64
        pw = X.real.astype(np.float64) ** 2 + X.imag.astype(np.float64) ** 2
        N32.append(np.log(np.maximum(pw @ fb, 1e-10)).astype(np.float32))      # fb: [257, 80]
    _report("cfg2_2x30s_80mel_numpy_float32_fft", logmel_error_stats(np.concatenate(N32), np.concatenate(R[:2])))


def test_vad_gate_ragged_batch(af, orc):
    """Gating on a ragged batch (mixed rates / lengths, an empty stream, a stream without speech, 80 mels and a mel
    count that is not a multiple of 4, segment capacity overflow) against the oracle."""
    import torch
    from audioflow import synth
    dev = torch.device("cuda")
    L = af.load_library()
    for M in (80, 30):
        specs = [(210, 6.0, 48000), (211, 4.1, 44100), (212, 0.0, 48000), (213, 3.3, 16000), (214, 9.7, 48000)]
        streams = [(synth.stream(i, sec, rate, 1), rate, 1) for (i, sec, rate) in specs]
        streams.append((np.full(48000 * 2, 1e-4, np.float32), 48000, 1))            # no speech at all
        S = len(streams)
        arrs = [torch.tensor(x, device=dev) if len(x) else torch.zeros(4, device=dev) for (x, _, _) in streams]
        pipe = af.Pipeline(af.pipeline_config(n_mels=M))
        b = pipe.batch([(a.data_ptr(), len(x), r, 1, af.AF_FMT_F32) for a, (x, r, _) in zip(arrs, streams)], af.AF_MEM_DEVICE)
        pcm = torch.zeros((S, b.pcm_stride), device=dev); lm = torch.zeros((S, b.logmel_stride), device=dev)
        vad = torch.zeros((S, b.vad_stride), device=dev, dtype=torch.uint8)
        torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
        b.run_device(b.outputs_struct(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, vad.data_ptr(), b.vad_stride))
        for seg_cap in (64, 2):
            seg = torch.zeros((S, seg_cap, 2), device=dev, dtype=torch.int32)
            nseg = torch.zeros(S, device=dev, dtype=torch.int32)
            nfr = torch.tensor(b.n_vad[:S].astype(np.int32), device=dev)
            nout = torch.tensor(b.n_out[:S].astype(np.int32), device=dev)
            torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
            assert L.af_vad_segments(vad.data_ptr(), b.vad_stride, nfr.data_ptr(), S, seg.data_ptr(), seg_cap, nseg.data_ptr(), None) == 0
            g_pcm = torch.full((S, b.pcm_stride), float("nan"), device=dev)
            g_lm = torch.full((S, b.logmel_stride), float("nan"), device=dev)
            g_off = torch.zeros((S, seg_cap + 1), device=dev, dtype=torch.int32)
            g_n = torch.zeros(S, device=dev, dtype=torch.int32)
            go = af.GateOutputsC(g_pcm.data_ptr(), b.pcm_stride, g_lm.data_ptr(), b.logmel_stride, g_off.data_ptr(), g_n.data_ptr())
            torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
            af._check(L.af_vad_gate(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, M, nout.data_ptr(), 160,
                                    seg.data_ptr(), seg_cap, nseg.data_ptr(), S, C.byref(go), None))
            torch.cuda.synchronize()
            some = 0
            for i in range(S):
                T, n_out = int(b.n_vad[i]), int(b.n_out[i])
                exp = orc.vad_segments(vad[i, :T].cpu().numpy())[:seg_cap]
                assert min(int(nseg[i]), seg_cap) == len(exp)
                rp, rl, roff = orc.vad_gate(pcm[i, :n_out].cpu().numpy(), lm[i, :T * M].reshape(T, M).cpu().numpy(), exp, 160)
                kept = int(roff[-1])
                some += kept
                assert int(g_n[i]) == kept
                assert np.array_equal(g_off[i, :len(exp) + 1].cpu().numpy().astype(np.uint32), roff)
                assert_bit_equal(g_pcm[i, :kept * 160].cpu().numpy(), rp, f"gated pcm {i}")
                assert np.array_equal(g_lm[i, :kept * M].cpu().numpy().view(np.uint32), rl.reshape(-1).view(np.uint32))
                assert torch.isnan(g_pcm[i, kept * 160:]).all() and torch.isnan(g_lm[i, kept * M:]).all()
            assert some > 0
        assert int(g_n[2]) == 0 and int(g_n[5]) == 0


# =============================================================================================
# cfg5: packed ticks across step / tile boundaries
# =============================================================================================
def _run_session_case(af, orc, S, n_ticks, sample_ids, device_twin=False):
    import torch
    from audioflow import synth
    rate, tick, M = 48000, 960, 80
    dev = torch.device("cuda")
    total = tick * n_ticks
    xd = synth.torch_batch(S, total / rate, rate, 1, dev, seed=500 + S)[:, :total].contiguous()
    xs = xd.cpu().numpy()
    pipe = af.Pipeline(af.pipeline_config(n_mels=M))
    ses = af.Session(pipe, S, rate, 1, af.AF_FMT_F32, max_tick_samples=tick)
    twin = None
    if device_twin:        # the bench's call: device-resident ticks and outputs
        L = af.load_library()
        twin = C.c_void_p()
        af._check(L.af_session_create(pipe._h, S, rate, 1, af.AF_FMT_F32, tick, C.byref(twin)))
        d_pcm = torch.zeros((S, 512), device=dev)
        d_lm = torch.zeros((S, 8 * M), device=dev)
        d_vad = torch.zeros((S, 16), device=dev, dtype=torch.uint8)
        d_o = af.OutputsC(d_pcm.data_ptr(), 512, d_lm.data_ptr(), 8 * M, d_vad.data_ptr(), 16, None, 0, None)
        u32 = C.c_uint32 * S
        c_pcm, c_feat, c_vad = u32(), u32(), u32()
    pcm = [[] for _ in sample_ids]
    lm = [[] for _ in sample_ids]
    vad = [[] for _ in sample_ids]
    counts = []
    for t in range(n_ticks):
        r = ses.push(xs[:, t * tick:(t + 1) * tick])
        counts.append((r["pcm"].shape[1], r["n_frames"]))
        for j, i in enumerate(sample_ids):
            pcm[j].append(r["pcm"][i]); lm[j].append(r["logmel"][i]); vad[j].append(r["vad"][i])
        if twin is not None:
            torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
            af._check(L.af_session_push(twin, xd.data_ptr() + t * tick * 4, xd.shape[1], tick, af.AF_MEM_DEVICE, C.byref(d_o),
                                        c_pcm, c_feat, c_vad))
            torch.cuda.synchronize()
            n_p, n_f = int(c_pcm[0]), int(c_feat[0])
            assert (n_p, n_f) == counts[-1] and int(c_vad[S - 1]) == n_f
            # every stream, every tick: the device-buffer session equals the host-buffer one bit for bit
            assert np.array_equal(d_pcm[:, :n_p].cpu().numpy().view(np.uint32), r["pcm"].view(np.uint32)), f"tick {t} pcm"
            assert np.array_equal(d_lm[:, :n_f * M].cpu().numpy().view(np.uint32), r["logmel"].reshape(S, -1).view(np.uint32)), f"tick {t} logmel"
            assert np.array_equal(d_vad[:, :n_f].cpu().numpy(), r["vad"]), f"tick {t} vad"
    if twin is not None:
        L.af_session_destroy(twin)
    fc = orc.default_feat_config(M)
    for j, i in enumerate(sample_ids):
        b = orc.BatchResampler(rate, 16000)
        ref_chunks = [b.process(xs[i, t * tick:(t + 1) * tick]) for t in range(n_ticks)]
        ref_pcm = np.concatenate(ref_chunks)
        if j == 0:
            assert [c[0] for c in counts] == [len(c) for c in ref_chunks]
        assert_bit_equal(np.concatenate(pcm[j]), ref_pcm, f"S={S} stream {i} pcm")
        T = orc.num_frames(len(ref_pcm))
        assert sum(c[1] for c in counts) == T
        st, _ = orc.VoiceActivityDetector().stream(ref_pcm, 400, 160)
        assert_bit_equal(np.concatenate(vad[j]), st, f"S={S} stream {i} vad")
        got_lm = np.concatenate(lm[j])
        assert got_lm.shape == (T, M)
        assert_logmel_close(got_lm, orc.logmel(ref_pcm, fc), f"S={S} stream {i} logmel")


@pytest.mark.parametrize("S", [17, 65])
def test_session_packed_ticks_cross_step_and_tile(af, orc, S):
    """16 streams fill one 32-frame step and 64 one 128-frame tile of a packed tick: S = 17 and 65 put a virtual stream
    across each boundary and leave a ragged last one.  Every stream is checked."""
    _run_session_case(af, orc, S, 50, list(range(S)))


def test_cfg5_1024_streams_50_ticks(af, orc):
    """BASELINE config 5 at full width: 1024 streams x 50 ticks of 20 ms.  Oracle parity on 20 sampled streams incl. the
    packing boundaries (15/16, 63/64, 1023); host-buffer and device-buffer sessions equal on ALL streams, every tick."""
    rng = np.random.default_rng(5)
    ids = sorted(set([0, 1, 15, 16, 17, 31, 32, 63, 64, 65, 127, 128, 511, 512, 1022, 1023] + rng.integers(0, 1024, 6).tolist()))
    _run_session_case(af, orc, 1024, 50, ids, device_twin=True)


def test_session_push_error_leaves_state_untouched(af, orc):
    """a recoverable caller error in af_session_push (an output stride that is too small) must not move the
    session: the same tick pushed again with good strides gives what an undisturbed session gives."""
    from audioflow import synth
    S, tick, rate, M = 3, 960, 48000, 80
    xs = np.stack([synth.stream(700 + i, 0.5, rate, 1)[:tick * 20] for i in range(S)])
    pipe = af.Pipeline(af.pipeline_config(n_mels=M))
    a, b = (af.Session(pipe, S, rate, 1, af.AF_FMT_F32, max_tick_samples=tick) for _ in range(2))
    L = af.load_library()
    pcm = np.zeros((S, 512), np.float32); lm = np.zeros((S, 8 * M), np.float32); vad = np.zeros((S, 16), np.uint8)
    rejected = 0
    for t in range(20):
        ra = a.push(xs[:, t * tick:(t + 1) * tick])
        x = np.ascontiguousarray(xs[:, t * tick:(t + 1) * tick])
        bad = [af.OutputsC(pcm.ctypes.data, 8, lm.ctypes.data, 8 * M, vad.ctypes.data, 16, None, 0, None)]             # pcm_stride
        if ra["n_frames"] > 0:        # only a tick that completes a frame needs the log-mel / state rows
            bad.append(af.OutputsC(pcm.ctypes.data, 512, lm.ctypes.data, 4, vad.ctypes.data, 16, None, 0, None))       # logmel_stride
            bad.append(af.OutputsC(pcm.ctypes.data, 512, lm.ctypes.data, 8 * M, vad.ctypes.data, 0, None, 0, None))    # vad_stride
        for o in bad:
            rc = L.af_session_push(b._h, x.ctypes.data, tick, tick, af.AF_MEM_HOST, C.byref(o), None, None, None)
            assert rc == af.AF_ERR_CAPACITY, (t, rc)
            rejected += 1
        rb = b.push(x)
        for k in ("pcm", "logmel", "vad"):
            assert_bit_equal(ra[k], rb[k], f"tick {t} {k}")
    assert rejected > 40


def test_misaligned_device_pcm_is_rejected(af):
    import torch
    from audioflow import synth
    dev = torch.device("cuda")
    x = synth.torch_batch(1, 1.0, 48000, 1, dev, seed=1)
    pipe = af.Pipeline(af.pipeline_config(n_mels=0, vad_enable=False))
    b = pipe.batch([(x[0].data_ptr(), x.shape[1], 48000, 1, af.AF_FMT_F32)], af.AF_MEM_DEVICE)
    pcm = torch.zeros(b.pcm_stride + 8, device=dev)
    o = b.outputs_struct(pcm.data_ptr() + 4, b.pcm_stride)
    with pytest.raises(ValueError):
        torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
        b.run_device(o)
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    b.run_device(b.outputs_struct(pcm.data_ptr(), b.pcm_stride))     # aligned: fine
    torch.cuda.synchronize()


# =============================================================================================
# long-scan sync point with silence_timeout_frames == 0
# =============================================================================================
@pytest.mark.parametrize("timeout,min_speech", [(0, 1), (0, 3), (1, 1)])
def test_many_cta_scan_timeout_zero(af, orc, timeout, min_speech):
    """vad.rs accepts silence_timeout_frames = 0: one non-speech frame ends Speech and the NEXT frame -- speech or not --
    is swallowed by Ending -> Silence (vad.rs:147-151).  On the many-CTA path (> 16384 frames) a block's sync point must
    therefore be a speech frame after TWO non-speech frames, not one.  Content: speech / single-zero flicker across every
    8192-frame block boundary."""
    rng = np.random.default_rng(3)
    fs, T_frames = 16000, 5 * 8192 + 100
    n = (T_frames - 1) * 160 + 400
    # frame-level pattern: mostly speech with isolated single non-speech frames (every ~7 frames), some double gaps
    x = (0.3 * rng.standard_normal(n)).astype(np.float32).clip(-1, 1)
    pos = 2000
    while pos < n - 2000:
        gap = int(rng.choice([400, 400, 400, 560, 720]))      # a window with no speech energy: 1, 2 or 3 silent frames
        x[pos:pos + gap] = 0.0
        pos += gap + int(rng.integers(3, 12)) * 160
    vc = af.VadConfig(threshold_db=-50.0, smoothing_factor=0.0, silence_timeout_frames=timeout, min_speech_frames=min_speech)
    oc = orc.default_vad_config()
    oc.threshold_db, oc.smoothing_factor, oc.silence_timeout_frames, oc.min_speech_frames = -50.0, 0.0, timeout, min_speech
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad=vc)).run_host([(x, fs, 1)])[0]
    ref = orc.pipeline_stream(x, 1, fs, None, oc, 400, 160, "f32")
    assert len(ref["vad"]) == T_frames and len(np.unique(ref["vad"])) >= 2
    assert_bit_equal(got["vad"], ref["vad"], f"timeout {timeout}")
    assert got["vad_final"]["state"] == ref["vad_final"]["state"]
    assert got["vad_final"]["speech_frames"] == ref["vad_final"]["speech_frames"]


def test_pcm16_wire_output_host_and_device(af, orc):
    """af_pipeline_config.pcm16: the batch delivers (x.clamp(-1,1) * 32767) as i16 (websocket.rs:246-251) instead of f32 --
    host-buffer and device-buffer runs, mixed rates / formats / ragged lengths, against the oracle's encode of the oracle's PCM."""
    import torch
    from audioflow import synth
    specs = [(610, 2.3, 48000, 1, "f32"), (611, 1.9, 44100, 2, "f32"), (612, 1.1, 48000, 1, "i16"), (613, 0.0, 48000, 1, "f32"), (614, 3.7, 32000, 1, "f32")]
    streams = [(synth.stream(i, sec, rate, ch, fmt), rate, ch) for (i, sec, rate, ch, fmt) in specs]
    streams[4][0][100:140] = np.array([1.5, -1.5, np.nan, 1.0, -1.0] * 8, np.float32)        # clamp / NaN / full scale (PCM checked only)
    pipe = af.Pipeline(af.pipeline_config(n_mels=80, pcm16=True))
    got = pipe.run_host(streams)
    refs = []
    for (x, rate, ch), g in zip(streams, got):
        fmt = "i16" if x.dtype == np.int16 else "f32"
        ref = orc.pipeline_stream(x, ch, rate, orc.default_feat_config(80), orc.default_vad_config(), 400, 160, fmt)
        refs.append(ref)
        assert g["pcm"].dtype == np.int16
        assert np.array_equal(g["pcm"], orc.pcm16_encode(ref["pcm"])), "host pcm16"
        if not np.isnan(ref["pcm"]).any():
            assert_bit_equal(g["vad"], ref["vad"], "vad with pcm16")
            assert_logmel_close(g["logmel"], ref["logmel"], "logmel with pcm16")
    # device buffers
    dev = torch.device("cuda")
    keep = [torch.tensor(x.astype(np.float32) if x.dtype != np.int16 else x, device=dev) if len(x) else torch.zeros(4, device=dev) for (x, _, _) in streams]
    descs = [(t.data_ptr(), len(x), rate, ch, af.AF_FMT_I16 if x.dtype == np.int16 else af.AF_FMT_F32) for t, (x, rate, ch) in zip(keep, streams)]
    b = pipe.batch(descs, af.AF_MEM_DEVICE)
    S = len(streams)
    pcm = torch.full((S, b.pcm_stride), -7, device=dev, dtype=torch.int16)
    torch.cuda.synchronize()
    b.run_device(b.outputs_struct(pcm.data_ptr(), b.pcm_stride))
    torch.cuda.synchronize()
    for i, ref in enumerate(refs):
        n = int(b.n_out[i])
        assert np.array_equal(pcm[i, :n].cpu().numpy(), orc.pcm16_encode(ref["pcm"])), f"device pcm16 {i}"


@pytest.mark.parametrize("seconds", [40.0, 200.0])
def test_scan_chains_hand_down_silence_count(af, orc, seconds):
    """The word walk of the scan kernels is chain-parallel: a word whose predecessor ends with timeout + 1 non-speech frames is
    walked from a fresh Silence.  What a fresh Silence cannot know is the silence_frames count Silence was entered with
    (vad.rs:136-145: Speech that times out below min_speech_frames goes straight to Silence and keeps the count; Ending ->
    Silence resets it).  Bursts shorter than min_speech_frames, bursts long enough for Ending, long pauses between them and a
    silent tail: the states and the WHOLE final machine (silence_frames included) must equal the sequential reference, on the
    one-CTA scan (40 s) and on the many-CTA one (200 s, > 16384 frames)."""
    fs = 16000
    rng = np.random.default_rng(int(seconds))
    n = int(seconds * fs)
    y = (1e-4 * rng.standard_normal(n)).astype(np.float32)
    pos = 8000
    k = 0
    while pos + 40000 < n - 6 * fs:                      # alternate 20 ms bursts (one or two frames) and 400 ms bursts, 0.6-2.2 s apart
        ln, amp = (320, 0.07) if k % 3 else (6400, 0.5)   # (a 320-sample burst at 0.07 lifts one or two frames over -50 dB)
        y[pos:pos + ln] += (amp * rng.standard_normal(ln)).astype(np.float32)
        pos += ln + int(rng.integers(int(0.6 * fs), int(2.2 * fs)))
        k += 1
    y[pos:pos + 320] += (0.07 * rng.standard_normal(320)).astype(np.float32)   # the last burst is a short one: Silence keeps the count
    vc = orc.default_vad_config()
    vc.smoothing_factor = 0.0                            # decisions follow the bursts frame by frame
    cfg = af.pipeline_config(n_mels=0, vad_enable=True)
    cfg.vad.smoothing_factor = 0.0
    got = af.Pipeline(cfg).run_host([(y, fs, 1)])[0]
    ref = orc.pipeline_stream(y, 1, fs, None, vc, 400, 160, "f32")
    assert_bit_equal(got["vad"], ref["vad"], "states")
    v = ref["vad"]
    assert (v == 2).any() and ((v[:-1] == 1) & (v[1:] == 0)).any()      # both ways out of Speech occur
    gf, rf = got["vad_final"], ref["vad_final"]
    assert (gf["state"], gf["speech_frames"], gf["silence_frames"]) == (rf["state"], rf["speech_frames"], rf["silence_frames"])
    assert rf["state"] == 0 and rf["silence_frames"] == 15
