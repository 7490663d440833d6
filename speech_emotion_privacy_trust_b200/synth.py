"""Synthetic corpora of the shapes SURVEY.md 8(d) specifies (there is no dataset access).

Speech-shaped audio: white Gaussian noise -> -3 dB/octave spectral tilt (x 1/sqrt(f) in the rFFT
domain) -> 3-5 Hz syllabic amplitude envelope -> peak normalised to 0.3.  16 kHz mono float32.
Host-side numpy only; this is workload generation, not part of the extraction path.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000


def speech_shaped(n_samples: int, rng: np.random.Generator, peak: float = 0.3) -> np.ndarray:
    white = rng.standard_normal(n_samples)
    spec = np.fft.rfft(white)
    f = np.fft.rfftfreq(n_samples, 1.0 / SAMPLE_RATE)
    tilt = np.ones_like(f)
    tilt[1:] = 1.0 / np.sqrt(f[1:])
    tilt[0] = 0.0
    x = np.fft.irfft(spec * tilt, n_samples)
    fm = rng.uniform(3.0, 5.0)
    t = np.arange(n_samples) / SAMPLE_RATE
    x *= 0.5 * (1.0 + np.sin(2.0 * np.pi * fm * t + rng.uniform(0, 2 * np.pi)))
    x *= peak / max(np.abs(x).max(), 1e-12)
    return x.astype(np.float32)


def utterance_lengths(n_utts: int, rng: np.random.Generator, lo_s: float = 2.0, hi_s: float = 10.0) -> np.ndarray:
    return rng.integers(int(lo_s * SAMPLE_RATE), int(hi_s * SAMPLE_RATE) + 1, size=n_utts).astype(np.int64)


def corpus(n_utts: int, seed: int = 1234, lo_s: float = 2.0, hi_s: float = 10.0):
    """Ragged corpus: returns (wav float32 [sum N], utt_off int64 [n+1])."""
    rng = np.random.default_rng(seed)
    lens = utterance_lengths(n_utts, rng, lo_s, hi_s)
    off = np.zeros(n_utts + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    wav = np.empty(int(off[-1]), dtype=np.float32)
    for i in range(n_utts):
        wav[off[i]:off[i + 1]] = speech_shaped(int(lens[i]), rng)
    return wav, off


def cloak_windows(n: int, seed: int = 8, win: int = 200, feat: int = 128, shift: float = 0.5):
    """z-normed-scale feature windows (n,1,win,feat) with class-dependent mean shifts so emotion (bands
    20-60) and gender (bands 0-20) are learnable.  Returns (x, emo, gen, speaker)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 1, win, feat)).astype(np.float32)
    emo = rng.integers(0, 4, size=n)
    gen = rng.integers(0, 2, size=n)
    spk = rng.integers(0, 10, size=n)
    for c in range(4):
        lo = 20 + 10 * c
        x[emo == c, :, :, lo:lo + 10] += shift
    x[gen == 1, :, :, 0:20] += shift
    x[gen == 0, :, :, 0:20] -= shift
    return x, emo.astype(np.int64), gen.astype(np.int64), spk.astype(np.int64)
