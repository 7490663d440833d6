#!/usr/bin/env python3
"""Smallest run that touches every kernel once (for compute-sanitizer memcheck / racecheck)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_emotion_privacy_trust_b200 import dropin, extraction, normalization, synth

rng = np.random.default_rng(0)
waves = [synth.speech_shaped(n, rng) for n in (1700, 4001, 9000)]
b = extraction.RaggedAudio.from_list(waves)
for n_fft, hop in ((800, 160), (1600, 160), (400, 200)):
    m, lay = extraction.logmel(b, n_fft=n_fft, hop=hop)
    extraction.logmel(b, n_fft=n_fft, hop=hop, band_major=True)
mf, _ = extraction.mfcc(b)
r = extraction.resample(extraction.RaggedAudio.from_list([synth.speech_shaped(5000, rng)]), 44100)
m, lay = extraction.logmel(b, n_fft=800)
st = normalization.speaker_stats(m, lay, ["a", "b", "a"])
normalization.normalize(m, lay, st)
wu, wt = normalization.window_table(lay)
x = normalization.normalized_windows(m, lay, st, wu, wt)
dropin.install()
import cloak_models
layer = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, "cuda").cuda()
xin = x[:2].clone().requires_grad_(True)
ya, yb = layer.forward_with_reversed_twin(xin, None, 0.1)
(ya.sum() + (yb * 2).sum()).backward()
torch.cuda.synchronize()
print("sanitize_small ok", float(m.mean()), float(mf.mean()), float(layer.locs.grad.abs().sum()))
