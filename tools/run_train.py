#!/usr/bin/env python3
"""Stand-alone cloak+GRL training-throughput run (profiling helper)."""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import benchmarks_train
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for cl, gr in ((False, False), (True, False), (True, True)):
    r = benchmarks_train.train_throughput(torch.device("cuda", 0), 0, 1, steps=steps, warmup=8, channels_last=cl, graphs=gr)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "memory_format", "cuda_graphs", "final_loss")}))
