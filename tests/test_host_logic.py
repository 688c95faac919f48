"""CPU tests of the product's host side: the C-ABI library loads, exports every symbol the header
declares, fails loudly without a GPU, and its planning logic (output lengths, fractional-offset
plan, VAD threshold in the energy domain) agrees with the oracle.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "audioflow_gpu.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AF_API\s+[\w\s\*]+?\b(af_\w+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    import audioflow
    L = audioflow.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 45
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.af_version()


def test_no_gpu_fails_loudly():
    import audioflow
    L = audioflow.load_library()
    n = C.c_int(0)
    L.af_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("GPU present")
    with pytest.raises(audioflow.NoDevice) as e:
        audioflow.init()
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(audioflow.NoDevice):
        audioflow.VoiceActivityDetector()
    with pytest.raises(audioflow.NoDevice):
        audioflow.AudioFrame.new([0.5, 0.25], 16000, 2).to_mono()
    with pytest.raises(audioflow.NoDevice):
        audioflow.AudioResampler(48000, 16000)


@pytest.mark.parametrize("rate", [48000, 44100, 32000, 22050, 8000, 16000, 96000, 11025])
def test_output_len_matches_oracle(orc, rate):
    import audioflow
    rng = np.random.default_rng(rate)
    lens = [0, 1, 127, 128, 129, 255, 256, 1000, 44100, 48000, 123457] + [int(v) for v in rng.integers(1, 300000, 6)]
    for n in lens:
        exp = len(orc.resample_stream(np.zeros(n, np.float32), rate))
        assert audioflow.resample_output_len(rate, 16000, n) == exp, (rate, n)


def test_output_len_tie_chunks_441(orc):
    """chunks = 193 mod 441 make (128 C - 8) * 160 / 441 an integer: the count hinges on f64 rounding."""
    import audioflow
    for chunks in (193, 193 + 441, 193 + 2 * 441):
        n = chunks * 128
        exp = len(orc.resample_stream(np.zeros(n, np.float32), 44100))
        assert audioflow.resample_output_len(44100, 16000, n) == exp


@pytest.mark.parametrize("rate,mode", [(48000, 1), (44100, 2), (32000, 1), (22050, 2), (8000, 1), (24000, 1), (11025, 2)])
def test_resample_plan_fracs_match_oracle(orc, rate, mode):
    import audioflow
    L = audioflow.load_library()
    chunks = 400
    x = np.zeros(chunks * 128, np.float32)
    y, frac = orc.resample_stream(x, rate, return_frac=True)
    buf = np.empty(len(frac) + 16, np.float32)
    m = C.c_int(-1)
    n = L.af_debug_resample_plan(rate, 16000, chunks, buf.ctypes.data_as(C.POINTER(C.c_float)), len(buf), C.byref(m))
    assert n == len(frac)
    assert m.value == mode            # 1 = exact recurrence (dyadic step), 2 = table
    assert np.array_equal(buf[:n], frac)
    # exact positions: k = floor(P), rem/q; the oracle frac must be rem/q (exact) or within 1e-9 of it
    g = np.gcd(rate, 16000)
    p, q = rate // g, 16000 // g
    nn = np.arange(n, dtype=np.int64)
    num = (nn + 1) * p - 4 * q
    rem = np.mod(num, q)
    d = rem / q - frac.astype(np.float64)
    d = d - np.round(d)               # frac ~ 1 when the recurrence sits just below an integer
    assert np.max(np.abs(d)) < 1e-6
    if mode == 1:
        assert np.array_equal((rem / q).astype(np.float32), frac)


@pytest.mark.parametrize("thr", [-50.0, -40.0, -30.0, -60.5, -100.0, 0.0, 10.0, -50.000004])
def test_vad_threshold_in_energy_domain(orc, thr):
    import audioflow
    L = audioflow.load_library()
    e_min = np.float32(L.af_debug_vad_energy_threshold(thr))
    assert e_min > 0
    bits = int(e_min.view(np.uint32))
    # exhaustive neighbourhood: is_speech(e) == (e >= e_min) for 20000 floats either side
    for b in list(range(bits - 2000, bits + 2000)) + [bits - 10 ** 6, bits + 10 ** 6, 1, 0x7f7fffff]:
        e = np.uint32(b).view(np.float32)
        assert (orc.energy_to_dbfs(float(e)) > np.float32(thr)) == bool(e >= e_min), (b, e)
    assert not (orc.energy_to_dbfs(0.0) > thr)


def test_vad_threshold_edge_cases():
    import audioflow
    L = audioflow.load_library()
    assert np.isnan(L.af_debug_vad_energy_threshold(float("nan")))
    assert np.isnan(L.af_debug_vad_energy_threshold(float("inf")))
    tiny = np.float32(L.af_debug_vad_energy_threshold(float("-inf")))
    assert tiny.view(np.uint32) == 1          # any positive energy is speech; 0 is not
