"""Deterministic, torch-version-independent model weights and inputs for the model-level golden vectors --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Checkpoints are far too large to commit (1.2 M parameters per classifier), so golden vectors pin OUTPUTS only and both
sides -- oracle/make_golden_models.py running the real reference, and the tests running the drop-ins / train_port --
rebuild the same weights from (seed, parameter name) with numpy's PCG64 streams.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def fill_state(module: torch.nn.Module, seed: int) -> None:
    """Overwrite every floating tensor of module.state_dict() in place, keyed by its name (not by iteration order)."""
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if not t.is_floating_point():
                continue
            r = _rng(seed, name)
            shape = tuple(t.shape)
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_var":
                v = r.uniform(0.5, 1.5, shape)
            elif leaf == "running_mean":
                v = 0.1 * r.standard_normal(shape)
            elif leaf in ("locs",):
                v = 0.05 * r.standard_normal(shape)
            elif leaf in ("rhos",):
                v = -2.0 + 0.8 * r.standard_normal(shape)
            elif t.dim() == 1 and leaf == "weight":             # batch-norm scale
                v = 1.0 + 0.1 * r.standard_normal(shape)
            elif t.dim() == 1 or "bias" in leaf:
                v = 0.05 * r.standard_normal(shape)
            elif leaf.startswith("att_mat"):
                v = r.uniform(0.0, 1.0, shape)
            else:
                fan_in = int(np.prod(shape[1:]))
                v = r.standard_normal(shape) / np.sqrt(fan_in)
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)).to(t.device))


def case_inputs(seed: int, batch: int = 3, win: int = 200, feat: int = 128):
    """(x, eps, mask, global_feature, g_emotion, g_gender): float32 numpy arrays for one golden case."""
    r = _rng(seed, "inputs")
    x = r.standard_normal((batch, 1, win, feat)).astype(np.float32)
    eps = (0.1 * r.standard_normal((1, win, feat))).astype(np.float32)
    mask = (r.uniform(size=(1, win, feat)) > 0.3).astype(np.float32)
    glob = r.standard_normal((batch, 88)).astype(np.float32)
    g1 = r.standard_normal((batch, 4)).astype(np.float32)
    g2 = r.standard_normal((batch, 2)).astype(np.float32)
    return x, eps, mask, glob, g1, g2


def dropout_off(model: torch.nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
        if isinstance(m, torch.nn.RNNBase):
            m.dropout = 0.0
