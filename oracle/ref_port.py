"""CPU port of the reference's two extraction callables -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (feature_extraction/audio_feature_extraction.py:15-46) builds torchaudio transforms on
every call and applies them to one utterance on the CPU.  torchaudio is the un-vendored third-party
dependency that holds the arithmetic (no version is pinned by the reference; the image has 2.11.0+cu128),
so this port drives the very same dependency the same way: one utterance per call, transforms constructed
inside the call, torch CPU threads as configured by the caller.  It is what `bench.py --impl reference`
and the `cpu_baseline` leg time ("kind": "port"), and one of the two oracles the parity tests compare
against (the other, oracle/restate.py, shares no code with torchaudio).

Pinned bit-for-bit against the real reference by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torchaudio

HOP_MEL = 160  # audio_feature_extraction.py:32
SAMPLE_RATE = 16000


def mel_spectrogram(audio: torch.Tensor, n_fft: int = 1024, feature_len: int = 128) -> torch.Tensor:
    """log-mel dB, (1, N) float32 CPU tensor -> (1, feature_len, 1 + N // 160) (reference :29-46)."""
    to_mel = torchaudio.transforms.MelSpectrogram(
        sample_rate=SAMPLE_RATE, n_fft=n_fft, win_length=n_fft, hop_length=HOP_MEL,
        n_mels=feature_len, window_fn=torch.hann_window)
    to_db = torchaudio.transforms.AmplitudeToDB()  # stype power, top_db None
    return to_db(to_mel(audio).detach())


def mfcc(audio: torch.Tensor) -> np.ndarray:
    """MFCC-40 of the waveform and of its two numerical derivatives -> (1, 120, 1 + N // 200) float32
    ndarray (reference :15-26).  np.gradient's second call uses spacing 2 (half the first)."""
    tf = torchaudio.transforms.MFCC(sample_rate=SAMPLE_RATE, n_mfcc=40)
    wave = audio[0]
    streams = [tf(audio).detach()]
    for spacing in (1, 2):
        d = np.gradient(wave, spacing)[None]
        streams.append(tf(torch.from_numpy(d)).detach())
    return np.concatenate(streams, axis=1)


def extract_all(audio: torch.Tensor) -> dict:
    """What one iteration of the reference's per-file loop computes (reference :185-187)."""
    return {"mfcc": mfcc(audio), "mel1": mel_spectrogram(audio, n_fft=800, feature_len=128),
            "mel2": mel_spectrogram(audio, n_fft=1600, feature_len=128)}


def resample(audio: torch.Tensor, sample_rate: int, new_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """What the reference does to a 44.1 kHz file before extraction (reference :139-141)."""
    return torchaudio.transforms.Resample(sample_rate, new_rate)(audio)
