#!/usr/bin/env python3
"""Pinned-memory PCIe copy bandwidth of the box: H2D alone, D2H alone, both at once (the ceiling of bench.py's e2e)."""
import torch
n = 1 << 29                                            # 2 GiB of fp32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


gb = n * 4 / 1e9
t = timed(h2d); print(f"H2D alone: {gb / t * 1e3:.1f} GB/s")
t = timed(d2h); print(f"D2H alone: {gb / t * 1e3:.1f} GB/s")
t = timed(lambda: (h2d(), d2h())); print(f"both at once: {gb / t * 1e3:.1f} GB/s each way ({t:.1f} ms for {gb:.2f} GB each)")
