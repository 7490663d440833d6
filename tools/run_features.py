#!/usr/bin/env python3
"""Run every extraction / normalisation kernel once over a synthetic corpus (profiling helper: wrap in ncu)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from speech_emotion_privacy_trust_b200 import extraction, normalization

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
dev = torch.device("cuda", 0)
lengths = bench.corpus_lengths(n, 1234)
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
batch = extraction.RaggedAudio(bench.synth_corpus_device(lengths, 4321, dev), off)
for rep in range(2):
    torch.cuda.nvtx.range_push("features" if rep else "warm")
    mel, lay = extraction.logmel(batch, n_fft=800)
    mel2, _ = extraction.logmel(batch, n_fft=1600)
    mf, _ = extraction.mfcc(batch)
    spk = [u % 10 for u in range(n)]
    st = normalization.speaker_stats(mel, lay, spk)
    z = normalization.normalize(mel, lay, st)
    wu, wt = normalization.window_table(lay)
    w = normalization.normalized_windows(mel, lay, st, wu, wt)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("frames", lay.total_frames, "windows", len(wu), "ok")
