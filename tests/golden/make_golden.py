"""Regenerates tests/golden/oracle_vectors.npz from the CPU oracle (oracle/oracle.c).

The reference itself cannot be executed in the build image (Rust, no toolchain), so
these vectors pin the ORACLE: they are checked (a) against the oracle on every CPU test run, so an
accidental change of the restatement shows up, and (b) against the CUDA path on the GPU box.
Inputs are regenerated from audioflow.synth (seeded), only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "audio-flow-rs_b200")]
import oracle  # noqa: E402
from audioflow import synth  # noqa: E402

# (stream id, seconds, rate, channels, fmt, n_mels)
CASES = [
    (0, 1.0, 48000, 2, "f32", 80),
    (1, 1.0, 44100, 1, "f32", 80),
    (2, 0.7, 48000, 1, "i16", 128),
    (3, 0.5, 16000, 1, "f32", 80),
]


def main():
    out = {}
    for (sid, sec, rate, ch, fmt, mels) in CASES:
        x = synth.stream(sid, sec, rate, ch, fmt)
        r = oracle.pipeline_stream(x, ch, rate, oracle.default_feat_config(mels), oracle.default_vad_config(), 400, 160, fmt)
        k = f"s{sid}"
        out[k + "_pcm"] = r["pcm"]
        out[k + "_logmel"] = r["logmel"]
        out[k + "_vad"] = r["vad"]
        out[k + "_energy"] = r["energy"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
