#!/usr/bin/env python3
"""Pinned-memory PCIe copy bandwidth of the box: H2D alone, D2H alone, both at once (the ceiling of bench.py's e2e).

    python tools/pcie_peak.py                                     # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_peak.py [--bind]

Under torchrun every rank drives its own GPU and all ranks copy AT THE SAME TIME (barrier before each measurement), so the
sum over ranks is what the host's memory / PCIe fabric sustains with N concurrent pinned streams.  --bind pins each rank's
host thread and pinned buffers to the CPUs / NUMA node NVML reports as local to its GPU (parallel.bind_host_to_gpu)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_emotion_privacy_trust_b200 import parallel

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
placement = parallel.bind_host_to_gpu(local) if "--bind" in sys.argv else None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29                                            # 2 GiB of fp32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
h_in.zero_(); h_out.zero_()                            # first touch by this (possibly bound) thread
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)     # all ranks copy concurrently: the slowest one defines the rate
    return float(ms)


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


gb = n * 4 / 1e9
rows = [("H2D alone", timed(h2d)), ("D2H alone", timed(d2h)), ("both at once", timed(lambda: (h2d(), d2h())))]
if rank == 0:
    print(f"{world} rank(s) copying concurrently, 2 GiB per direction per rank, pinned host memory, placement: {placement or 'default'}")
    for name, t in rows:
        per = gb / t * 1e3
        print(f"{name}: {per:.1f} GB/s per rank each way, {per * world:.1f} GB/s over {world} rank(s) ({t:.1f} ms)")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
