#!/usr/bin/env python3
"""Per-kernel time of ONE eager cloak+GRL training step (channels_last, B = 32): run under
`ncu --metrics gpu__time_duration.sum --profile-from-start off` to see where the 5.3 ms of the graphed step go."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import benchmarks_train
from speech_emotion_privacy_trust_b200 import losses, synth

dev = torch.device("cuda", 0)
model = benchmarks_train.build_model(dev).train().to(memory_format=torch.channels_last)
torch.backends.cudnn.benchmark = True
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4)
x, emo, gen, _ = synth.cloak_windows(32, seed=8)
xb, eb, gb = torch.from_numpy(x).to(dev), torch.from_numpy(emo).to(dev), torch.from_numpy(gen).to(dev)
w = torch.ones(32, device=dev)
for i in range(6):
    if i == 5:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()          # ncu --profile-from-start off: every thread's kernels (backward runs on autograd's)
    p1, p2, _ = model(xb, pooling="mean")
    loss = losses.cloak_grl_loss(p1, p2, eb, gb, w, 0.1)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
