#!/usr/bin/env python3
"""Time the other features of the reference's extraction script over the bench corpus (n_fft 1600, MFCC, all three);
`python tools/time_features.py mfcc` runs only MFCC a few times (profiling helper: wrap in ncu)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from speech_emotion_privacy_trust_b200 import extraction

dev = torch.device("cuda", 0)
lengths = bench.corpus_lengths(bench.CORPUS_UTTS, 1234)
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
batch = extraction.RaggedAudio(bench.synth_corpus_device(lengths, 4321, dev), off)
hours = float(lengths.sum()) / 16000 / 3600
if len(sys.argv) > 1 and sys.argv[1] == "mfcc":
    out = torch.empty(batch.layout(400, 200).total_frames * 120, dtype=torch.float32, device=dev)
    for _ in range(3):
        extraction.mfcc(batch, out=out)
    torch.cuda.synchronize()
else:
    r = bench.other_features(batch, hours, dev)
    print({k: (round(v["ms"], 3), round(v["fp32_frac"], 3)) for k, v in r.items()})
