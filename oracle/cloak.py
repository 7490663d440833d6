"""numpy restatement of the cloak noise layer and gradient reversal -- TEST INFRASTRUCTURE ONLY.

Follows model/cloak_models.py:24-58 (cloak_noise) and model/reversal_gradient.py:5-23.  float32
arithmetic in the reference's operation order; pinned against the real modules by
tests/test_oracle_golden.py (golden vectors from oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def scales(rhos, min_scale, max_scale):
    """sigma = (1 + tanh(rho)) / 2 * (max - min) + min   (cloak_models.py:41-43)."""
    rhos = np.asarray(rhos, F32)
    return ((F32(1.0) + np.tanh(rhos)) / F32(2) * F32(max_scale - min_scale) + F32(min_scale)).astype(F32)


def forward(x, locs, rhos, eps, min_scale, max_scale, mask=None):
    """out = x[*mask] + (locs + sigma * eps[*mask])   (cloak_models.py:45-58).  x (B,1,W,F); rest (1,W,F)."""
    x = np.asarray(x, F32)
    e = np.asarray(eps, F32)
    if mask is not None:
        e = (e * np.asarray(mask, F32)).astype(F32)
    noise = (np.asarray(locs, F32) + scales(rhos, min_scale, max_scale) * e).astype(F32)
    if mask is None:
        return (x + noise).astype(F32)
    return (x * np.asarray(mask, F32) + noise).astype(F32)


def backward(g_a, rhos, eps, min_scale, max_scale, mask=None, g_b=None, lambda_=0.0, dtype=np.float64):
    """Gradients of forward() for upstream grad g_a on the noisy output, optionally joined with a second
    upstream grad g_b that arrives through a gradient-reversal layer (reversal_gradient.py:19-23):
        g = g_a - lambda * g_b
        dlocs = sum_b g ; drhos = sum_b g * eps[*mask] * (1 - tanh(rho)^2) / 2 * (max - min) ; dx = g[*mask]
    Accumulated in `dtype` (float64 = ground truth for the batch reduction)."""
    g = np.asarray(g_a, dtype)
    if g_b is not None:
        g = g - dtype(lambda_) * np.asarray(g_b, dtype)
    e = np.asarray(eps, dtype)
    m = None if mask is None else np.asarray(mask, dtype)
    if m is not None:
        e = e * m
    t = np.tanh(np.asarray(rhos, dtype))
    dsig = (1.0 - t * t) / 2.0 * (max_scale - min_scale)
    dlocs = g.sum(axis=0)
    drhos = (g * e).sum(axis=0) * dsig
    dx = g if m is None else g * m
    return dlocs.reshape(np.shape(rhos)), drhos.reshape(np.shape(rhos)), dx


def grl_backward(g, lambda_):
    """dx = -lambda * g   (reversal_gradient.py:19-23)."""
    g = np.asarray(g, F32)
    return (-F32(lambda_) * g).astype(F32)
