#!/usr/bin/env python3
"""Time the sinc resampler (44.1 kHz -> 16 kHz) over a synthetic ragged batch (CUDA events, after warm-up)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_emotion_privacy_trust_b200 import extraction

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
rng = np.random.default_rng(7)
lens = rng.integers(2 * 44100, 10 * 44100 + 1, size=n)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
wav = torch.randn(int(off[-1]), device="cuda") * 0.1
batch = extraction.RaggedAudio(wav, off)
for _ in range(3):
    out = extraction.resample(batch, 44100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = extraction.resample(batch, 44100)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
hours = float(lens.sum()) / 44100 / 3600
print(f"resample 44.1k->16k: {ms:.3f} ms for {hours:.2f} audio-hours = {hours / (ms * 1e-3):.0f} audio-hours/s "
      f"({(wav.numel() + out.wav.numel()) * 4 / ms / 1e6:.0f} GB/s of HBM traffic)")
