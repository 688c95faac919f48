"""Synthetic PCM of SURVEY.md 8(d): concatenated 0.5-2 s segments of
{noise floor, tone complex below / at / above the VAD threshold}.  Test and bench input only."""
from __future__ import annotations

import numpy as np

SEED0 = 0xA0D10F10
PEAKS = (0.0, 0.02, 0.08, 0.3)     # noise-only, below threshold, borderline, clear speech-like


def stream(i: int, seconds: float, rate: int, channels: int = 1, fmt: str = "f32") -> np.ndarray:
    """Interleaved samples of synthetic stream i (numpy, deterministic in i)."""
    rng = np.random.default_rng(SEED0 + i)
    n = int(round(seconds * rate))
    out = np.empty((n, channels), np.float32)
    for c in range(channels):
        x = rng.normal(0.0, 1e-3, n)
        pos = 0
        t = np.arange(n) / rate
        while pos < n:
            seg = int(rng.uniform(0.5, 2.0) * rate)
            peak = PEAKS[int(rng.integers(0, 4))]
            end = min(n, pos + seg)
            if peak > 0:
                freqs = rng.uniform(100.0, 4000.0, 5)
                ph = rng.uniform(0, 2 * np.pi, 5)
                tone = np.zeros(end - pos)
                for f, p in zip(freqs, ph):
                    tone += np.sin(2 * np.pi * f * t[pos:end] + p)
                x[pos:end] += tone * (peak / 5.0)
            pos = end
        out[:, c] = np.clip(x, -1.0, 1.0).astype(np.float32)
    flat = out.reshape(-1)
    if fmt == "i16":
        return np.round(flat * 32767.0).astype(np.int16)
    return flat


def torch_batch(n_streams: int, seconds: float, rate: int, channels: int, device, seed: int = 0, dtype=None):
    """[n_streams, n * channels] synthetic batch generated on `device` with torch (bench input)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED0 + seed)
    n = int(round(seconds * rate))
    x = torch.empty((n_streams, n * channels), device=device, dtype=torch.float32)
    x.normal_(0.0, 1e-3, generator=g)
    seg = rate // 2                                     # 0.5 s envelope granularity
    n_seg = (n + seg - 1) // seg
    cls = torch.randint(0, 4, (n_streams, n_seg), device=device, generator=g)
    peaks = torch.tensor(PEAKS, device=device)[cls]                       # [S, n_seg]
    env = peaks.repeat_interleave(seg, dim=1)[:, :n]                      # [S, n]
    t = torch.arange(n, device=device, dtype=torch.float32) / rate
    for c in range(channels):
        tone = torch.zeros((n_streams, n), device=device)
        freqs = torch.empty((n_streams, 5), device=device).uniform_(100.0, 4000.0, generator=g)
        ph = torch.empty((n_streams, 5), device=device).uniform_(0, 6.2831853, generator=g)
        for k in range(5):
            tone += torch.sin(6.2831853 * freqs[:, k:k + 1] * t[None, :] + ph[:, k:k + 1])
        x[:, c::channels] += tone * env * 0.2
        del tone
    x.clamp_(-1.0, 1.0)
    if dtype is not None and dtype == torch.int16:
        return torch.round(x * 32767.0).to(torch.int16)
    return x
