#!/usr/bin/env python3
"""Where a warp of extract_kernel spends its cycles (diagnostic build only).

    SEPT_NVCC_EXTRA=-DSEPT_PHASE_CLOCKS python tools/build_variant.py variants/clk.so
    SEPT_LIB_PATH=variants/clk.so python tools/phase_clocks.py [n_fft]

The variant accumulates clock64() deltas between the phase boundaries of every warp (lane 0, atomicAdd) into a device
array; this script runs the bench corpus once and prints cycles per item per phase.  Numbers from this build are
diagnostics, never bench values."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from speech_emotion_privacy_trust_b200 import _lib, extraction

NAMES = ["wait stage", "pass 1", "syncwarp", "locate+prefetch", "pass 2 + split", "syncwarp", "store P", "syncwarp", "mel+log+store", "syncwarp"]
n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 800
hop = 200 if n_fft == 400 else 160
dev = torch.device("cuda", 0)
lengths = bench.corpus_lengths(bench.CORPUS_UTTS, 1234)
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
batch = extraction.RaggedAudio(bench.synth_corpus_device(lengths, 4321, dev), off)
lay = batch.layout(n_fft, hop)
out = torch.empty((lay.total_frames, 128), device=dev)
lib = _lib.lib()
fn = lib.sept_debug_phase_clocks
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
for _ in range(3):
    extraction.logmel(batch, n_fft=n_fft, hop=hop, out=out)
buf = (ctypes.c_ulonglong * 16)()
assert fn(ctypes.addressof(buf), 1) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
extraction.logmel(batch, n_fft=n_fft, hop=hop, out=out)
e1.record()
assert fn(ctypes.addressof(buf), 1) == 0
items = int(lay.item_off[-1])
tot = sum(buf[:10])
print(f"n_fft {n_fft}: {items} items, {e0.elapsed_time(e1):.3f} ms (instrumented), {tot / items:.0f} cycles per item per warp")
for i, name in enumerate(NAMES):
    print(f"  {i} {name:18s} {buf[i] / items:8.0f} cycles/item  {100.0 * buf[i] / tot:5.1f} %")
