#!/usr/bin/env python3
"""Cloak forward + fused cloak/GRL backward at B=64, W=200, F=128 (config 2 shape), a few iterations (profiling helper)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from speech_emotion_privacy_trust_b200 import dropin
dropin.install()
import cloak_models
torch.manual_seed(0)
layer = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, "cuda").cuda()
xs = [torch.randn(64, 1, 200, 128, device="cuda") for _ in range(12)]        # 12 x 6.5 MB inputs + outputs + grads > L2
for i in range(12):
    ya, yb = layer.forward_with_reversed_twin(xs[i], None, 0.1)
    ((ya * xs[(i + 1) % 12]).sum() + (yb * xs[(i + 2) % 12]).sum()).backward()
torch.cuda.synchronize()
print("ok", float(layer.locs.grad.abs().mean()))
