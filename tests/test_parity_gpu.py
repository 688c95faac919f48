"""GPU parity tests: the CUDA path behind the C ABI against the CPU oracle and the committed golden
vectors.  Bit-exact for PCM, frame energies, VAD states and counts; log-mel within
max-abs 1e-4 and relative L2 1e-5 (the tolerance BASELINE.json's north_star states)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
LOGMEL_ABS, LOGMEL_REL_L2 = 1e-4, 1e-5


def _frame(spec):
    return np.full(spec["len"], spec["value"], np.float32)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype == np.float32:
        bad = np.nonzero(_bits(a) != _bits(b))[0]
        assert bad.size == 0, f"{what}: {bad.size} of {a.size} differ, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"
    else:
        assert np.array_equal(a, b), what


F32_ULP, NOISE_ULPS = 2.0 ** -24, 2.0


def logmel_tolerance(ref, log10=False):
    """Stated tolerance of the f32 log-mel against the f64 oracle ("Tolerances"): 1e-4 (the north star's
    figure) for every mel bin within ~52 dB of its frame's strongest mel bin.  Below that an f32 transform is
    conditioning-limited: the rounding noise of the strong bins lands on the weak ones, so the bound is stated on the
    AMPLITUDE noise instead -- at most NOISE_ULPS = 2 f32 ulps of the frame's strongest mel-band amplitude, i.e.
    |d log mel| <= 2 * NOISE_ULPS * 2^-24 * sqrt(peak / mel).  (numpy's own float32 FFT leaves the same band on the same
    frames; tests/test_configs_gpu.py::test_cfg2_logmel_error_report prints both.)"""
    ref = ref.astype(np.float64)
    ln = ref * (np.log(10.0) if log10 else 1.0)
    dr = ln.max(axis=1, keepdims=True) - ln                      # natural-log dynamic range below the frame peak
    return np.maximum(LOGMEL_ABS, 2.0 * NOISE_ULPS * F32_ULP * np.exp(0.5 * dr))


def assert_logmel_close(got, ref, what="", log10=False):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    rel = np.sqrt((err ** 2).sum() / max((ref.astype(np.float64) ** 2).sum(), 1e-30))
    tol = logmel_tolerance(ref, log10)
    worst = np.unravel_index((err / tol).argmax(), err.shape)
    assert (err <= tol).all(), f"{what}: err {err[worst]:.3e} > tol {tol[worst]:.3e} at {worst}; max abs err {err.max():.3e}"
    assert rel <= LOGMEL_REL_L2, f"{what}: rel L2 {rel:.3e}"
    return err.max(), rel


# =============================================================================================
# the reference's own tests, through the API mirror
# =============================================================================================
def test_audio_frame_to_mono_single_channel(af):          # capture.rs:372-382
    frame = af.AudioFrame.new([0.5, -0.5], 16000, 1, 1000)
    mono = frame.to_mono()
    assert mono.channels == 1
    assert mono.samples.tolist() == [0.5, -0.5]


def test_audio_frame_to_mono_stereo(af):                  # capture.rs:385-400
    frame = af.AudioFrame.new([0.5, 0.25, -0.5, -0.25], 16000, 2, 1000)
    mono = frame.to_mono()
    assert mono.channels == 1 and len(mono.samples) == 2
    assert abs(mono.samples[0] - 0.375) < 0.001
    assert abs(mono.samples[1] - (-0.375)) < 0.001


def test_no_resample_needed(af):                          # resampler.rs:185-190
    r = af.AudioResampler.new(16000, 16000)
    x = np.array([0.1, 0.2, 0.3, 0.4], np.float32)
    assert np.array_equal(r.process(x), x)


def test_resample_rates(af):                              # resampler.rs:193-198
    r = af.AudioResampler.new(48000, 16000)
    assert r.input_rate() == 48000 and r.output_rate() == 16000 and r.needs_resampling()


def test_same_rates_no_resampling(af):                    # resampler.rs:200-203
    assert not af.AudioResampler.new(48000, 48000).needs_resampling()


def test_vad_silence_detection(af):                       # vad.rs:212-223
    vad = af.VoiceActivityDetector.new(af.VadConfig(threshold_db=-50.0))
    assert vad.detect(np.full(480, 0.0001, np.float32)) == af.VadState.Silence


def test_vad_speech_detection(af):                        # vad.rs:226-237
    vad = af.VoiceActivityDetector.new(af.VadConfig(threshold_db=-50.0))
    assert vad.detect(np.full(480, 0.5, np.float32)) == af.VadState.Speech


def test_vad_state_transitions(af):                       # vad.rs:240-265
    vad = af.VoiceActivityDetector.new(af.VadConfig(threshold_db=-50.0, silence_timeout_frames=2,
                                                    min_speech_frames=1, smoothing_factor=0.0))
    assert vad.state() == af.VadState.Silence
    speech, silence = np.full(480, 0.5, np.float32), np.full(480, 0.0001, np.float32)
    assert vad.detect(speech) == af.VadState.Speech
    assert vad.detect(silence) == af.VadState.Speech
    assert vad.detect(silence) == af.VadState.Ending
    assert vad.detect(silence) == af.VadState.Silence


def test_vad_reset(af):                                   # vad.rs:268-281
    vad = af.VoiceActivityDetector.new(af.VadConfig())
    vad.detect(np.full(480, 0.5, np.float32))
    assert vad.is_speaking()
    vad.reset()
    assert vad.state() == af.VadState.Silence and not vad.is_speaking()


def test_energy_calculation(af):                          # vad.rs:284-298
    vad = af.VoiceActivityDetector.new(af.VadConfig())
    assert vad.calculate_energy(np.zeros(480, np.float32)) == 0.0
    assert abs(vad.calculate_energy(np.full(480, 0.5, np.float32)) - 0.25) < 0.0001


def test_kat_json_through_gpu(af):
    for case in KAT["to_mono"]:
        out = af.AudioFrame.new(case["samples"], 16000, case["channels"]).to_mono().samples
        exp = np.array(case["expect"], np.float32)
        assert np.all(np.abs(out - exp) <= case.get("tol", 0.0))
    for case in KAT["vad"]:
        v = af.VoiceActivityDetector(af.VadConfig(**case["config"]))
        got = [v.detect(_frame(f)).name for f in case["frames"]]
        assert got == case["expect"]


# =============================================================================================
# compat objects against the oracle, bit exact
# =============================================================================================
@pytest.mark.parametrize("channels", [2, 3, 6])
def test_to_mono_matches_oracle(af, orc, channels):
    rng = np.random.default_rng(channels)
    x = rng.standard_normal(channels * 1000 + (channels - 1)).astype(np.float32)   # trailing partial frame
    x[::7] = -0.0
    got = af.AudioFrame.new(x, 48000, channels).to_mono().samples
    assert_bit_equal(got, orc.to_mono(x, channels), "to_mono")


@pytest.mark.parametrize("rate", [48000, 44100, 32000, 22050, 8000])
def test_audio_resampler_chunks_match_oracle(af, orc, rate):
    rng = np.random.default_rng(rate)
    x = rng.standard_normal(128 * 40).astype(np.float32)
    g, o = af.AudioResampler.new(rate, 16000), orc.AudioResampler(rate, 16000)
    for c in range(40):
        a, b = g.process(x[c * 128:(c + 1) * 128]), o.process(x[c * 128:(c + 1) * 128])
        assert_bit_equal(a, b, f"chunk {c} @ {rate}")


def test_audio_resampler_short_input_is_resampling_failed(af):
    r = af.AudioResampler.new(48000, 16000)
    with pytest.raises(af.ResamplingFailed) as e:
        r.process(np.zeros(100, np.float32))
    assert "Insufficient buffer size 100" in str(e.value)
    a = af.AudioResampler.new(48000, 16000).process(np.arange(300, dtype=np.float32))
    b = af.AudioResampler.new(48000, 16000).process(np.arange(128, dtype=np.float32))
    assert_bit_equal(a, b, "extra input ignored")


@pytest.mark.parametrize("rate", [48000, 44100, 11025])
def test_batch_resampler_matches_oracle(af, orc, rate):
    rng = np.random.default_rng(rate + 1)
    x = rng.standard_normal(30000).astype(np.float32)
    g, o = af.BatchResampler.new(rate, 16000), orc.BatchResampler(rate, 16000)
    pos = 0
    while pos < len(x):
        n = int(rng.integers(1, 3000))
        assert_bit_equal(g.process(x[pos:pos + n]), o.process(x[pos:pos + n]), f"process @ {pos}")
        pos += n
    assert_bit_equal(g.flush(), o.flush(), "flush")
    assert len(g.flush()) == 0


def test_vad_sequences_match_oracle(af, orc):
    rng = np.random.default_rng(7)
    for alpha, thr in [(0.3, -50.0), (0.0, -50.0), (0.9, -30.0), (0.05, -62.5)]:
        cfg = af.VadConfig(threshold_db=thr, smoothing_factor=alpha, silence_timeout_frames=4, min_speech_frames=2)
        oc = orc.default_vad_config()
        oc.threshold_db, oc.smoothing_factor, oc.silence_timeout_frames, oc.min_speech_frames = thr, alpha, 4, 2
        g, o = af.VoiceActivityDetector(cfg), orc.VoiceActivityDetector(oc)
        for i in range(60):
            amp = float(rng.choice([0.0, 1e-4, 3e-3, 0.05, 0.06, 0.5]))
            fr = (amp * rng.standard_normal(int(rng.integers(1, 700)))).astype(np.float32)
            assert int(g.detect(fr)) == o.detect(fr), (alpha, thr, i)
            assert np.float32(g.smoothed_energy()).view(np.uint32) == np.float32(o.smoothed_energy()).view(np.uint32)
            assert g.speech_frame_count() == o.speech_frame_count()
            a, b = g.energy_db(), o.energy_db()
            assert a == b or (np.isinf(a) and np.isinf(b))
        assert int(g.detect(np.zeros(0, np.float32))) == o.detect(np.zeros(0, np.float32))


def test_vad_detect_frames_matches_oracle(af, orc):
    from audioflow import synth
    y = synth.stream(11, 4.0, 16000)
    g, o = af.VoiceActivityDetector(), orc.VoiceActivityDetector()
    st, _ = o.stream(y, 320, 320)
    assert_bit_equal(g.detect_frames(y, 320, 320), st, "20 ms frames")
    assert g.speech_frame_count() == o.speech_frame_count() and int(g.state()) == o.state()


def test_pcm16_matches_oracle(af, orc):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1.5, 1.5, 5000), [0.0, 1.0, -1.0, np.nan, 1e-5, -1e-5, 0.99999]]).astype(np.float32)
    assert np.array_equal(af.pcm16_encode(x), orc.pcm16_encode(x))


def test_pcm16_base64_wire_payload(af, orc):              # websocket.rs:244-254 ("audio_base_64")
    rng = np.random.default_rng(6)
    special = np.array([0.0, 1.0, -1.0, np.nan, 1e-5, -1e-5, 0.99999, 2.0, -3.0, np.inf, -np.inf], np.float32)
    for n in (0, 1, 2, 3, 4, 5, 6, 7, 320, 959, 960, 16000, 100003):
        x = rng.uniform(-1.2, 1.2, n).astype(np.float32)
        x[:min(n, len(special))] = special[:min(n, len(special))]
        got = af.pcm16_base64(x)
        assert len(got) == 4 * ((2 * n + 2) // 3)
        assert got == orc.pcm16_base64(x), f"n = {n}"


# =============================================================================================
# the batched pipeline
# =============================================================================================
def _oracle_case(orc, x, ch, rate, mels, fmt, vad_len=400, vad_hop=160, vcfg=None):
    return orc.pipeline_stream(x, ch, rate, orc.default_feat_config(mels) if mels else None,
                               vcfg if vcfg is not None else orc.default_vad_config(), vad_len, vad_hop, fmt)


def _check_stream(got, ref, what, log10=False):
    assert_bit_equal(got["pcm"], ref["pcm"], what + " pcm")
    if ref.get("vad") is not None and got["vad"] is not None:
        assert_bit_equal(got["energy"], ref["energy"], what + " energy")
        assert_bit_equal(got["vad"], ref["vad"], what + " vad")
        assert got["vad_final"]["state"] == ref["vad_final"]["state"]
        assert got["vad_final"]["speech_frames"] == ref["vad_final"]["speech_frames"]
        assert np.float32(got["vad_final"]["smoothed"]).view(np.uint32) == np.float32(ref["vad_final"]["smoothed"]).view(np.uint32)
    if ref.get("logmel") is not None and got["logmel"] is not None:
        assert_logmel_close(got["logmel"], ref["logmel"], what + " logmel", log10)


@pytest.mark.parametrize("variant", ["sync", "tma"])
def test_golden_vectors(af, variant):
    """The committed oracle vectors (tests/golden/oracle_vectors.npz) against the CUDA path."""
    from audioflow import synth
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    try:
        af.set_kernel_variant(variant)
        for (sid, sec, rate, ch, fmt, mels) in mg.CASES:
            x = synth.stream(sid, sec, rate, ch, fmt)
            got = af.Pipeline(af.pipeline_config(n_mels=mels)).run_host([(x, rate, ch)])[0]
            k = f"s{sid}"
            assert_bit_equal(got["pcm"], g[k + "_pcm"], k + " pcm")
            assert_bit_equal(got["energy"], g[k + "_energy"], k + " energy")
            assert_bit_equal(got["vad"], g[k + "_vad"], k + " vad")
            assert_logmel_close(got["logmel"], g[k + "_logmel"], k + " logmel")
    finally:
        af.set_kernel_variant("auto")


@pytest.mark.parametrize("variant", ["sync", "tma"])
def test_batch_mixed_streams_match_oracle(af, orc, variant):
    """Ragged batch: mixed rates, channels, formats and lengths incl. empty, < 1 chunk, < 1 frame,
    multi-tile (> 20480 output samples) and tile-boundary lengths."""
    from audioflow import synth
    cases = [  # (id, seconds, rate, channels, fmt)
        (20, 3.1, 48000, 1, "f32"), (21, 2.9, 44100, 1, "f32"), (22, 1.3, 48000, 2, "f32"),
        (23, 1.7, 44100, 2, "i16"), (24, 2.0, 16000, 1, "f32"), (25, 0.0, 48000, 1, "f32"),
        (26, 0.002, 48000, 1, "f32"), (27, 0.02, 48000, 1, "f32"), (28, 1.28 + 0.025, 48000, 1, "f32"),
        (29, 1.0, 32000, 1, "i16"), (30, 0.9, 22050, 3, "f32"), (31, 1.28, 48000, 1, "f32"),
        (32, 2.5601, 48000, 1, "i16"), (33, 0.4, 8000, 1, "f32"),
        # f32 stereo is staged in quarter steps: 44.1 kHz (general-ratio quads), a length that ends inside a quarter,
        # and one that ends exactly on a step boundary (5120 outputs = 0.32 s)
        (34, 2.1, 44100, 2, "f32"), (35, 1.28 + 0.333, 48000, 2, "f32"), (36, 0.32 * 3 + 0.0001, 48000, 2, "f32"),
    ]
    streams = []
    for (sid, sec, rate, ch, fmt) in cases:
        x = synth.stream(sid, sec, rate, ch, fmt)
        if sid == 30:
            x = x[:-2]                       # trailing partial frame of a 3-channel stream
        streams.append((x, rate, ch))
    try:
        af.set_kernel_variant(variant)
        got = af.Pipeline(af.pipeline_config(n_mels=80)).run_host(streams)
    finally:
        af.set_kernel_variant("auto")
    for (case, (x, rate, ch), g) in zip(cases, streams, got):
        fmt = "i16" if x.dtype == np.int16 else "f32"
        ref = _oracle_case(orc, x, ch, rate, 80, fmt)
        _check_stream(g, ref, f"stream {case}")


def test_batch_128_mels_log10_and_custom_vad(af, orc):
    from audioflow import synth
    x = synth.stream(40, 2.2, 48000, 2)
    vc = af.VadConfig(threshold_db=-45.0, smoothing_factor=0.5, silence_timeout_frames=5, min_speech_frames=2)
    oc = orc.default_vad_config()
    oc.threshold_db, oc.smoothing_factor, oc.silence_timeout_frames, oc.min_speech_frames = -45.0, 0.5, 5, 2
    # 128 mels, log10, VAD on 20 ms non-overlapping frames (BASELINE config 1 / 4 flavour)
    got = af.Pipeline(af.pipeline_config(n_mels=128, vad=vc, vad_frame_len=320, vad_hop=320, log10=True)).run_host([(x, 48000, 2)])[0]
    fc = orc.default_feat_config(128)
    fc.log10_flag = 1
    ref = orc.pipeline_stream(x, 2, 48000, fc, oc, 320, 320)
    _check_stream(got, ref, "128 mel / log10 / 20 ms VAD", log10=True)
    # resample + VAD only (no features, PCM not returned)
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad=vc, vad_frame_len=320, vad_hop=320, write_pcm=False)).run_host([(x, 48000, 2)])[0]
    assert got["pcm"] is None and got["logmel"] is None
    assert_bit_equal(got["vad"], ref["vad"], "vad only")


def test_config1_reference_clip(af, orc):
    """BASELINE config 1: single 10 s 48 kHz stereo f32 clip -> mono 16 kHz + VAD (20 ms frames)."""
    from audioflow import synth
    x = synth.stream(1, 10.0, 48000, 2)
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad_frame_len=320, vad_hop=320)).run_host([(x, 48000, 2)])[0]
    ref = orc.pipeline_stream(x, 2, 48000, None, orc.default_vad_config(), 320, 320)
    assert len(got["pcm"]) == 159998
    _check_stream(got, ref, "config 1")
    # and the same clip driven through the per-object API, as the reference's processing loop would
    mono = af.AudioFrame.new(x, 48000, 2).to_mono()
    rs = af.BatchResampler.new(48000, 16000)
    pcm = np.concatenate([rs.process(mono.samples), rs.flush()])
    assert_bit_equal(pcm, ref["pcm"], "object API pcm")
    vad = af.VoiceActivityDetector.new(af.VadConfig())
    states = [int(vad.detect(pcm[i:i + 320])) for i in range(0, len(pcm) - 319, 320)]
    assert states == ref["vad"].tolist()


def test_device_buffers_and_full_size_properties(af, orc):
    """BASELINE config 2 at full size on device buffers: 256 x 30 s 48 kHz mono -> PCM + 80-mel + VAD.
    Size-independent properties on the whole batch, oracle parity on a sample of streams."""
    import torch
    from audioflow import synth
    dev = torch.device("cuda")
    S, sec, rate = 256, 30.0, 48000
    x = synth.torch_batch(S, sec, rate, 1, dev, seed=3)
    n = x.shape[1]
    pipe = af.Pipeline(af.pipeline_config(n_mels=80))
    descs = [(x[i].data_ptr(), n, rate, 1, af.AF_FMT_F32) for i in range(S)]
    b = pipe.batch(descs, af.AF_MEM_DEVICE)
    assert int(b.n_out[0]) == 479998 and int(b.n_feat[0]) == 2998
    pcm = torch.full((S, b.pcm_stride), float("nan"), device=dev)
    lm = torch.full((S, b.logmel_stride), float("nan"), device=dev)
    vad = torch.full((S, b.vad_stride), 255, device=dev, dtype=torch.uint8)
    en = torch.zeros((S, b.energy_stride), device=dev)
    fin = torch.zeros((S, 6), device=dev, dtype=torch.int32)
    o = b.outputs_struct(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, vad.data_ptr(), b.vad_stride,
                         en.data_ptr(), b.energy_stride, fin.data_ptr())
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    b.run_device(o)
    torch.cuda.synchronize()
    n_out, T = int(b.n_out[0]), int(b.n_feat[0])
    # (1) 48k -> 16k is pure decimation: y[n] = x[3n - 1], y[0] = 0
    assert torch.equal(pcm[:, 1:n_out], x[:, 2:3 * n_out - 1:3])
    assert bool((pcm[:, 0] == 0).all())
    # (2) every defined output was written, nothing outside its range was touched
    assert not torch.isnan(lm[:, :T * 80]).any() and torch.isnan(lm[:, T * 80:]).all()
    assert torch.isnan(pcm[:, n_out:]).all()
    assert int(vad[:, :T].max()) <= 2 and bool((vad[:, T:] == 255).all())
    # (3) energies are the mean squares of the frames (fp tolerance; exact order checked below)
    fr = pcm[:8, :n_out].unfold(1, 400, 160)
    assert torch.allclose(en[:8, :T], (fr.double() ** 2).mean(-1).float(), rtol=1e-5, atol=1e-12)
    # (4) idempotence: a second run reproduces every byte
    pcm2, lm2, vad2 = pcm.clone(), lm.clone(), vad.clone()
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    b.run_device(o)
    torch.cuda.synchronize()
    assert torch.equal(pcm2[:, :n_out], pcm[:, :n_out]) and torch.equal(lm2[:, :T * 80], lm[:, :T * 80]) and torch.equal(vad2, vad)
    # (5) oracle parity on a sample of streams
    for i in (0, 101, 255):
        ref = _oracle_case(orc, x[i].cpu().numpy(), 1, rate, 80, "f32")
        got = {"pcm": pcm[i, :n_out].cpu().numpy(), "logmel": lm[i, :T * 80].reshape(T, 80).cpu().numpy(),
               "vad": vad[i, :T].cpu().numpy(), "energy": en[i, :T].cpu().numpy(),
               "vad_final": dict(state=int(fin[i, 1]), smoothed=float(fin[i, 0:1].view(torch.float32)[0]),
                                 speech_frames=int(fin[i, 4]))}
        _check_stream(got, ref, f"cfg2 stream {i}")
    # (6) segmentation of the device states agrees with the oracle's
    seg = torch.zeros((S, 64, 2), device=dev, dtype=torch.int32)
    nseg = torch.zeros(S, device=dev, dtype=torch.int32)
    nfr = torch.tensor(b.n_vad[:S].astype(np.int32), device=dev)
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    rc = af.load_library().af_vad_segments(vad.data_ptr(), b.vad_stride, nfr.data_ptr(), S, seg.data_ptr(), 64,
                                           nseg.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    for i in (0, 17):
        exp = orc.vad_segments(vad[i, :T].cpu().numpy())
        assert int(nseg[i]) == len(exp)
        assert seg[i, :len(exp)].cpu().numpy().tolist() == exp.tolist()


def test_mixed_rate_device_batch_properties(af, orc):
    """BASELINE config 3 flavour (mixed 44.1/48 kHz, device buffers), reduced to 64 streams."""
    import torch
    from audioflow import synth
    dev = torch.device("cuda")
    S, sec = 64, 30.0
    xs = [synth.torch_batch(1, sec, 44100 if i % 2 else 48000, 1, dev, seed=100 + i)[0] for i in range(S)]
    descs = [(xs[i].data_ptr(), xs[i].numel(), 44100 if i % 2 else 48000, 1, af.AF_FMT_F32) for i in range(S)]
    pipe = af.Pipeline(af.pipeline_config(n_mels=80))
    b = pipe.batch(descs, af.AF_MEM_DEVICE)
    assert int(b.n_out[0]) == 479998 and int(b.n_out[1]) == 480001
    pcm = torch.zeros((S, b.pcm_stride), device=dev)
    lm = torch.zeros((S, b.logmel_stride), device=dev)
    vad = torch.zeros((S, b.vad_stride), device=dev, dtype=torch.uint8)
    en = torch.zeros((S, b.energy_stride), device=dev)
    o = b.outputs_struct(pcm.data_ptr(), b.pcm_stride, lm.data_ptr(), b.logmel_stride, vad.data_ptr(), b.vad_stride,
                         en.data_ptr(), b.energy_stride, 0)
    torch.cuda.synchronize()            # torch fills run on the legacy stream; the library uses its own non-blocking stream
    b.run_device(o)
    torch.cuda.synchronize()
    for i in (1, 2, 63):
        rate = 44100 if i % 2 else 48000
        ref = _oracle_case(orc, xs[i].cpu().numpy(), 1, rate, 80, "f32")
        n_out, T = int(b.n_out[i]), int(b.n_feat[i])
        assert_bit_equal(pcm[i, :n_out].cpu().numpy(), ref["pcm"], f"cfg3 pcm {i}")
        assert_bit_equal(en[i, :T].cpu().numpy(), ref["energy"], f"cfg3 energy {i}")
        assert_bit_equal(vad[i, :T].cpu().numpy(), ref["vad"], f"cfg3 vad {i}")
        assert_logmel_close(lm[i, :T * 80].reshape(T, 80).cpu().numpy(), ref["logmel"], f"cfg3 logmel {i}")


def test_linearity_of_resampler(af):
    """resample(a x) == a resample(x) for a power of two (exact in floating point), at full length."""
    from audioflow import synth
    x = synth.stream(50, 30.0, 44100, 1)
    p = af.Pipeline(af.pipeline_config(n_mels=0, vad_enable=False))
    a, b = p.run_host([(x, 44100, 1), (x * np.float32(0.25), 44100, 1)])
    assert_bit_equal(a["pcm"] * np.float32(0.25), b["pcm"], "linearity")


@pytest.mark.parametrize("alpha", [0.0, 0.05, 0.3, 0.9, 1.0])
def test_parallel_vad_scan_long_streams(af, orc, alpha):
    """The chunk-parallel EMA / state-machine scan (af_vad_scan_par_kernel) against the sequential oracle
    (vad.rs:97-154): streams longer than one scan block (8192 frames), every smoothing regime (alpha = 0 uses the
    raw energy, 0.05 falls back to the sequential kernel, 1 has no memory), plus a loud-then-nearly-silent stream
    whose EMA takes longer than the speculative warm-up to forget its past -- the verified fallback must give
    the same bits."""
    rng = np.random.default_rng(7)
    n1 = 16000 * 100                                                        # 100 s @ 16 kHz -> 9998 frames, two blocks
    seg = rng.integers(0, 4, n1 // 4000 + 1).repeat(4000)[:n1]
    amp = np.choose(seg, [1e-3, 0.02, 0.08, 0.3]).astype(np.float32)
    x1 = (amp * rng.standard_normal(n1).astype(np.float32)).clip(-1, 1)
    x2 = np.concatenate([0.9 * np.ones(16000 * 3, np.float32),              # loud, then 1e-6: the EMA decays for > 100 frames
                         1e-6 * rng.standard_normal(16000 * 60).astype(np.float32)])
    x3 = x1[:16000 * 7 + 123]
    vc = af.VadConfig(threshold_db=-50.0, smoothing_factor=alpha, silence_timeout_frames=15, min_speech_frames=3)
    oc = orc.default_vad_config()
    oc.threshold_db, oc.smoothing_factor, oc.silence_timeout_frames, oc.min_speech_frames = -50.0, alpha, 15, 3
    streams = [(x1, 16000, 1), (x2, 16000, 1), (x3, 16000, 1)]
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad=vc)).run_host(streams)
    for i, ((x, rate, ch), g) in enumerate(zip(streams, got)):
        ref = orc.pipeline_stream(x, ch, rate, None, oc, 400, 160, "f32")
        assert len(ref["vad"]) == 1 + (len(x) - 400) // 160
        _check_stream(g, ref, f"alpha {alpha} stream {i}")


@pytest.mark.parametrize("rate,ch", [(48000, 1), (48000, 2), (44100, 1), (32000, 1)])
def test_resampler_special_values(af, orc, rate, ch):
    """The frac = 0 shortcut of the 48 kHz path (y = y1 when y1 != 0 and every tap is finite with |x| < 2) and the
    unfused cubic everywhere else must stay bit-exact on the values that break shortcuts: +-0 taps, exact zeros as y1,
    |x| >= 2, huge values, subnormals, +-Inf and NaN (NaN positions must match; payloads are implementation-defined)."""
    from audioflow import synth
    n = int(1.6 * rate) * ch
    x = synth.stream(140 + ch, 1.6, rate, ch)[:n].copy()
    rng = np.random.default_rng(rate + ch)
    specials = np.array([0.0, -0.0, 2.0, -2.0, 1.9999999, 3.5, -7.25, 1e30, -1e30, 1e-40, -1e-41, 1.17549435e-38,
                         np.inf, -np.inf, np.nan, 65504.0, 1.0, -1.0], np.float32)
    # isolated specials, runs of zeros (whole quads with y1 == 0) and runs of large values
    for k in range(400):
        x[int(rng.integers(0, n))] = specials[k % len(specials)]
    for k in range(12):
        s = int(rng.integers(0, n - 200))
        x[s:s + int(rng.integers(3, 120))] = (0.0, -0.0, 2.5, 1e-39)[k % 4]
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad_enable=False)).run_host([(x, rate, ch)])[0]["pcm"]
    ref = orc.resample_stream(orc.to_mono(x, ch), rate)
    assert got.shape == ref.shape
    nan = np.isnan(ref)
    assert nan.any() and (~nan).sum() > 0.9 * len(ref)
    assert np.array_equal(np.isnan(got), nan), "NaN positions differ"
    assert_bit_equal(got[~nan], ref[~nan], f"special values {rate} Hz x{ch}")


@pytest.mark.parametrize("alpha", [0.3, 0.0, 1.0])
def test_many_cta_vad_scan_very_long_streams(af, orc, alpha):
    """Streams of more than two scan blocks (> 16384 frames) take the many-CTA scan: speculative block walks from a fresh
    machine, exact from each block's first sync point (speech after silence_timeout + 1 non-speech frames), a sequential
    carry pass that re-walks the words in front of it.  Against the sequential oracle (vad.rs:97-154), on content built
    to break every shortcut: speech that runs across whole blocks (no sync point), borderline flicker, silence across
    block boundaries (entry state Silence with a non-zero counter), and a loud-to-nearly-silent step right before a
    block boundary (frame 8192 = 81.92 s) so that the EMA warm-up speculation of the next block is wrong."""
    rng = np.random.default_rng(11)
    fs = 16000
    def noise(sec, a): return (a * rng.standard_normal(int(sec * fs))).astype(np.float32).clip(-1, 1)
    # 1: segments of all four classes, 330 s -> 32 998 frames, five blocks
    seg = rng.integers(0, 4, 330 * fs // 4000 + 1).repeat(4000)[:330 * fs]
    x1 = (np.choose(seg, [1e-3, 0.02, 0.08, 0.3]).astype(np.float32) * rng.standard_normal(330 * fs).astype(np.float32)).clip(-1, 1)
    # 2: 200 s of continuous speech (two whole blocks without a sync point), then silence, then flicker
    x2 = np.concatenate([noise(5, 1e-3), noise(200, 0.3), noise(40, 1e-3), noise(60, 0.056)])
    # 3: loud until just before frame 8192, then 1e-6: the next block's EMA warm-up (128 frames) starts inside the loud part
    x3 = np.concatenate([0.9 * np.ones(int(81.7 * fs), np.float32), noise(120, 1e-6), noise(30, 0.3), noise(30, 1e-3)])
    # 4: silence across three block boundaries, speech at the very end
    x4 = np.concatenate([noise(2, 0.3), noise(260, 1e-3), noise(3, 0.3)])
    vc = af.VadConfig(threshold_db=-50.0, smoothing_factor=alpha, silence_timeout_frames=15, min_speech_frames=3)
    oc = orc.default_vad_config()
    oc.threshold_db, oc.smoothing_factor, oc.silence_timeout_frames, oc.min_speech_frames = -50.0, alpha, 15, 3
    streams = [(x, fs, 1) for x in (x1, x2, x3, x4)] + [(x1[:fs * 20], fs, 1)]     # + a short stream in the same batch
    got = af.Pipeline(af.pipeline_config(n_mels=0, vad=vc)).run_host(streams)
    for i, ((x, rate, ch), g) in enumerate(zip(streams, got)):
        ref = orc.pipeline_stream(x, ch, rate, None, oc, 400, 160, "f32")
        assert len(ref["vad"]) == 1 + (len(x) - 400) // 160
        _check_stream(g, ref, f"many-CTA scan, alpha {alpha}, stream {i}")
        assert g["vad_final"]["state"] == int(ref["vad"][-1])
