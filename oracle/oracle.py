"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (audio-flow-rs_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

SILENCE, SPEECH, ENDING = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


class VadConfig(C.Structure):
    """vad.rs:21-32"""
    _fields_ = [("threshold_db", C.c_float), ("smoothing_factor", C.c_float),
                ("silence_timeout_frames", C.c_uint64), ("min_speech_frames", C.c_uint64)]


class FeatConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("n_fft", C.c_uint32), ("win_length", C.c_uint32),
                ("hop_length", C.c_uint32), ("n_mels", C.c_uint32), ("f_min", C.c_float),
                ("f_max", C.c_float), ("log_floor", C.c_float), ("log10_flag", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    u8p = C.POINTER(C.c_uint8)
    sz = C.c_size_t
    L.orc_to_mono.restype = sz
    L.orc_to_mono.argtypes = [fp, sz, C.c_uint, fp]
    L.orc_i16_to_f32.argtypes = [C.POINTER(C.c_int16), sz, fp]
    L.orc_resampler_new.restype = C.c_void_p
    L.orc_resampler_new.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_resampler_free.argtypes = [C.c_void_p]
    L.orc_resampler_process.restype = C.c_int
    L.orc_resampler_process.argtypes = [C.c_void_p, fp, sz, fp, sz, C.POINTER(sz)]
    L.orc_batch_new.restype = C.c_void_p
    L.orc_batch_new.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_batch_free.argtypes = [C.c_void_p]
    L.orc_batch_process.restype = C.c_int
    L.orc_batch_process.argtypes = [C.c_void_p, fp, sz, fp, sz, C.POINTER(sz)]
    L.orc_batch_flush.restype = C.c_int
    L.orc_batch_flush.argtypes = [C.c_void_p, fp, sz, C.POINTER(sz)]
    L.orc_resample_stream.restype = sz
    L.orc_resample_stream.argtypes = [C.c_uint32, C.c_uint32, fp, sz, fp, sz, fp]
    L.orc_resample_max_output.restype = sz
    L.orc_resample_max_output.argtypes = [C.c_uint32, C.c_uint32, sz]
    L.orc_vad_default_config.argtypes = [C.POINTER(VadConfig)]
    L.orc_vad_new.restype = C.c_void_p
    L.orc_vad_new.argtypes = [C.POINTER(VadConfig)]
    L.orc_vad_free.argtypes = [C.c_void_p]
    L.orc_frame_energy.restype = C.c_float
    L.orc_frame_energy.argtypes = [fp, sz]
    L.orc_energy_to_dbfs.restype = C.c_float
    L.orc_energy_to_dbfs.argtypes = [C.c_float]
    L.orc_vad_detect.restype = C.c_int
    L.orc_vad_detect.argtypes = [C.c_void_p, fp, sz]
    L.orc_vad_detect_energy.restype = C.c_int
    L.orc_vad_detect_energy.argtypes = [C.c_void_p, C.c_float]
    L.orc_vad_reset.argtypes = [C.c_void_p]
    L.orc_vad_state.restype = C.c_int
    L.orc_vad_state.argtypes = [C.c_void_p]
    L.orc_vad_energy_db.restype = C.c_float
    L.orc_vad_energy_db.argtypes = [C.c_void_p]
    L.orc_vad_is_speaking.restype = C.c_int
    L.orc_vad_is_speaking.argtypes = [C.c_void_p]
    L.orc_vad_speech_frame_count.restype = C.c_uint64
    L.orc_vad_speech_frame_count.argtypes = [C.c_void_p]
    L.orc_vad_silence_frames.restype = C.c_uint64
    L.orc_vad_silence_frames.argtypes = [C.c_void_p]
    L.orc_vad_smoothed_energy.restype = C.c_float
    L.orc_vad_smoothed_energy.argtypes = [C.c_void_p]
    L.orc_vad_stream.restype = sz
    L.orc_vad_stream.argtypes = [C.c_void_p, fp, sz, sz, sz, u8p, fp]
    L.orc_feat_default_config.argtypes = [C.POINTER(FeatConfig)]
    L.orc_hann_window.argtypes = [C.c_uint32, fp]
    L.orc_mel_filterbank.argtypes = [C.POINTER(FeatConfig), fp]
    L.orc_num_frames.restype = sz
    L.orc_num_frames.argtypes = [sz, C.c_uint32, C.c_uint32]
    L.orc_logmel.restype = sz
    L.orc_logmel.argtypes = [fp, sz, C.POINTER(FeatConfig), fp, fp]
    L.orc_feat_plan_new.restype = C.c_void_p
    L.orc_feat_plan_new.argtypes = [C.POINTER(FeatConfig)]
    L.orc_feat_plan_free.argtypes = [C.c_void_p]
    L.orc_logmel_f32.restype = sz
    L.orc_logmel_f32.argtypes = [C.c_void_p, fp, sz, fp]
    L.orc_pcm16_encode.argtypes = [fp, sz, C.POINTER(C.c_int16)]
    L.orc_vad_segments.restype = sz
    L.orc_vad_segments.argtypes = [u8p, sz, C.POINTER(C.c_uint32), sz]
    L.orc_vad_gate.restype = sz
    L.orc_vad_gate.argtypes = [fp, sz, fp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), sz, fp, fp, C.POINTER(C.c_uint32)]
    L.orc_pipeline_stream.restype = sz
    L.orc_pipeline_stream.argtypes = [fp, sz, C.c_uint, C.c_uint32, C.c_void_p, C.POINTER(VadConfig),
                                      C.c_uint32, C.c_uint32, fp, fp, sz, fp, u8p, C.POINTER(sz)]
    L.orc_pipeline_stream_shaped.restype = sz
    L.orc_pipeline_stream_shaped.argtypes = [fp, sz, C.c_uint, C.c_uint32, sz, C.POINTER(VadConfig), C.c_uint32, fp, sz, u8p,
                                             C.POINTER(sz)]
    _lib = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def to_mono(samples, channels: int) -> np.ndarray:
    """AudioFrame::to_mono (capture.rs:30-42)."""
    x = _f32(samples)
    out = np.empty((len(x) + max(channels, 1) - 1) // max(channels, 1) if channels > 1 else len(x), np.float32)
    n = lib().orc_to_mono(_fp(x), len(x), channels, _fp(out))
    return out[:n]


def i16_to_f32(samples) -> np.ndarray:
    x = np.ascontiguousarray(samples, dtype=np.int16)
    out = np.empty(len(x), np.float32)
    lib().orc_i16_to_f32(x.ctypes.data_as(C.POINTER(C.c_int16)), len(x), _fp(out))
    return out


class ResamplingFailed(Exception):
    """AudioError::ResamplingFailed (src-tauri/src/error.rs:109-110)."""


class AudioResampler:
    """resampler.rs:12-112 (intended semantics, see oracle.c header)."""

    def __init__(self, input_rate: int, output_rate: int):
        self._h = lib().orc_resampler_new(input_rate, output_rate)
        self._in, self._out = input_rate, output_rate

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_resampler_free(self._h)
            self._h = None

    def process(self, x) -> np.ndarray:
        x = _f32(x)
        out = np.empty(max(len(x), lib().orc_resample_max_output(self._in, self._out, 128)), np.float32)
        n = C.c_size_t(0)
        st = lib().orc_resampler_process(self._h, _fp(x), len(x), _fp(out), len(out), C.byref(n))
        if st == 1:
            raise ResamplingFailed("Insufficient buffer size %d for input channel 0, expected 128" % len(x))
        assert st == 0
        return out[:n.value].copy()

    def input_rate(self): return self._in
    def output_rate(self): return self._out
    def needs_resampling(self): return self._in != self._out


class BatchResampler:
    """resampler.rs:115-166."""

    def __init__(self, input_rate: int, output_rate: int):
        self._h = lib().orc_batch_new(input_rate, output_rate)
        self._in, self._out = input_rate, output_rate

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_batch_free(self._h)
            self._h = None

    def process(self, x) -> np.ndarray:
        x = _f32(x)
        cap = lib().orc_resample_max_output(self._in, self._out, len(x) + 128)
        out = np.empty(cap, np.float32)
        n = C.c_size_t(0)
        st = lib().orc_batch_process(self._h, _fp(x), len(x), _fp(out), cap, C.byref(n))
        assert st == 0
        return out[:n.value].copy()

    def flush(self) -> np.ndarray:
        out = np.empty(lib().orc_resample_max_output(self._in, self._out, 128), np.float32)
        n = C.c_size_t(0)
        st = lib().orc_batch_flush(self._h, _fp(out), len(out), C.byref(n))
        assert st == 0
        return out[:n.value].copy()


def resample_stream(x, in_rate: int, out_rate: int = 16000, return_frac: bool = False):
    """BatchResampler::process(all) + flush() in one go."""
    x = _f32(x)
    cap = lib().orc_resample_max_output(in_rate, out_rate, len(x))
    out = np.empty(cap, np.float32)
    frac = np.empty(cap, np.float32) if return_frac else None
    n = lib().orc_resample_stream(in_rate, out_rate, _fp(x), len(x), _fp(out), cap,
                                  _fp(frac) if return_frac else None)
    if return_frac:
        return out[:n].copy(), frac[:n].copy()
    return out[:n].copy()


def default_vad_config() -> VadConfig:
    c = VadConfig()
    lib().orc_vad_default_config(C.byref(c))
    return c


class VoiceActivityDetector:
    """vad.rs:60-205."""

    def __init__(self, config: VadConfig | None = None):
        self.config = config if config is not None else default_vad_config()
        self._h = lib().orc_vad_new(C.byref(self.config))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vad_free(self._h)
            self._h = None

    def detect(self, frame) -> int:
        f = _f32(frame)
        return lib().orc_vad_detect(self._h, _fp(f), len(f))

    def detect_energy(self, e: float) -> int:
        return lib().orc_vad_detect_energy(self._h, C.c_float(e))

    def calculate_energy(self, frame) -> float:
        f = _f32(frame)
        return float(lib().orc_frame_energy(_fp(f), len(f)))

    def reset(self): lib().orc_vad_reset(self._h)
    def state(self) -> int: return lib().orc_vad_state(self._h)
    def energy_db(self) -> float: return float(lib().orc_vad_energy_db(self._h))
    def is_speaking(self) -> bool: return bool(lib().orc_vad_is_speaking(self._h))
    def speech_frame_count(self) -> int: return int(lib().orc_vad_speech_frame_count(self._h))
    def silence_frames(self) -> int: return int(lib().orc_vad_silence_frames(self._h))
    def smoothed_energy(self) -> float: return float(lib().orc_vad_smoothed_energy(self._h))

    def stream(self, y, frame_len: int, hop: int):
        """Framed VAD over a whole 16 kHz signal -> (states u8[T], energies f32[T])."""
        y = _f32(y)
        T = lib().orc_num_frames(len(y), frame_len, hop)
        st = np.zeros(T, np.uint8)
        en = np.zeros(T, np.float32)
        n = lib().orc_vad_stream(self._h, _fp(y), len(y), frame_len, hop,
                                 st.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(en))
        assert n == T
        return st, en


def energy_to_dbfs(e: float) -> float:
    return float(lib().orc_energy_to_dbfs(C.c_float(e)))


def default_feat_config(n_mels: int = 80) -> FeatConfig:
    c = FeatConfig()
    lib().orc_feat_default_config(C.byref(c))
    c.n_mels = n_mels
    return c


def hann_window(win: int) -> np.ndarray:
    w = np.empty(win, np.float32)
    lib().orc_hann_window(win, _fp(w))
    return w


def mel_filterbank(cfg: FeatConfig) -> np.ndarray:
    fb = np.empty((cfg.n_fft // 2 + 1, cfg.n_mels), np.float32)
    lib().orc_mel_filterbank(C.byref(cfg), _fp(fb))
    return fb


def num_frames(n: int, win: int = 400, hop: int = 160) -> int:
    return int(lib().orc_num_frames(n, win, hop))


def logmel(y, cfg: FeatConfig | None = None, return_power: bool = False):
    """Spec-defined log-mel (f64 internal) -> f32 [T, n_mels]."""
    cfg = cfg if cfg is not None else default_feat_config()
    y = _f32(y)
    T = num_frames(len(y), cfg.win_length, cfg.hop_length)
    out = np.empty((T, cfg.n_mels), np.float32)
    pw = np.empty((T, cfg.n_fft // 2 + 1), np.float32) if return_power else None
    n = lib().orc_logmel(_fp(y), len(y), C.byref(cfg), _fp(out), _fp(pw) if return_power else None)
    assert n == T
    return (out, pw) if return_power else out


class FeatPlan:
    """f32 CPU feature path, used only as the timed CPU baseline."""

    def __init__(self, cfg: FeatConfig):
        self.cfg = cfg
        self._h = lib().orc_feat_plan_new(C.byref(cfg))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_feat_plan_free(self._h)
            self._h = None

    def logmel(self, y) -> np.ndarray:
        y = _f32(y)
        T = num_frames(len(y), self.cfg.win_length, self.cfg.hop_length)
        out = np.empty((T, self.cfg.n_mels), np.float32)
        lib().orc_logmel_f32(self._h, _fp(y), len(y), _fp(out))
        return out


def pcm16_encode(x) -> np.ndarray:
    """websocket.rs:246-251."""
    x = _f32(x)
    out = np.empty(len(x), np.int16)
    lib().orc_pcm16_encode(_fp(x), len(x), out.ctypes.data_as(C.POINTER(C.c_int16)))
    return out



def pcm16_base64(x) -> bytes:
    """websocket.rs:244-254 / :338-348: base64 (standard alphabet) of the little-endian bytes of pcm16_encode(x)."""
    import base64
    return base64.b64encode(pcm16_encode(x).astype("<i2").tobytes())


def vad_segments(states) -> np.ndarray:
    s = np.ascontiguousarray(states, dtype=np.uint8)
    cap = len(s) // 2 + 2
    seg = np.zeros((cap, 2), np.uint32)
    n = lib().orc_vad_segments(s.ctypes.data_as(C.POINTER(C.c_uint8)), len(s),
                               seg.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
    return seg[:n].copy()


def vad_gate(pcm, logmel, seg, hop: int = 160):
    """Speech-only PCM / log-mel of the segments `seg` ([n, 2] frames), packed: (pcm, logmel, off[n + 1])."""
    seg = np.ascontiguousarray(seg, dtype=np.uint32).reshape(-1, 2)
    kept = int((seg[:, 1] - seg[:, 0]).sum()) if len(seg) else 0
    pcm = _f32(pcm) if pcm is not None else None
    lm = np.ascontiguousarray(logmel, dtype=np.float32) if logmel is not None else None
    M = lm.shape[1] if lm is not None else 0
    out_p = np.zeros(kept * hop, np.float32) if pcm is not None else None
    out_l = np.zeros((kept, M), np.float32) if lm is not None else None
    off = np.zeros(len(seg) + 1, np.uint32)
    n = lib().orc_vad_gate(_fp(pcm) if pcm is not None else None, len(pcm) if pcm is not None else 0,
                           _fp(lm) if lm is not None else None, M, hop, seg.ctypes.data_as(C.POINTER(C.c_uint32)), len(seg),
                           _fp(out_p) if out_p is not None else None, _fp(out_l) if out_l is not None else None,
                           off.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert n == kept
    return out_p, out_l, off


def pipeline_stream(samples, channels: int, in_rate: int, feat: FeatConfig | None, vad: VadConfig | None,
                    vad_len: int = 400, vad_hop: int = 160, fmt: str = "f32"):
    """Whole path for one stream with the PARITY oracle (f64 features).

    returns dict(pcm, logmel, vad, energy)."""
    x = i16_to_f32(samples) if fmt == "i16" else _f32(samples)
    mono = to_mono(x, channels)
    pcm = resample_stream(mono, in_rate, 16000)
    out = {"pcm": pcm, "logmel": None, "vad": None, "energy": None}
    if feat is not None:
        out["logmel"] = logmel(pcm, feat)
    if vad is not None:
        v = VoiceActivityDetector(vad)
        out["vad"], out["energy"] = v.stream(pcm, vad_len, vad_hop)
        out["vad_final"] = dict(state=v.state(), smoothed=v.smoothed_energy(),
                                speech_frames=v.speech_frame_count(), silence_frames=v.silence_frames())
    return out
