"""The loss of the cloak + GRL training step as ONE vectorised expression.

The reference accumulates it sample by sample in Python (training/training_cloak_with_grl.py:141-154): for every batch
element two `nn.CrossEntropyLoss` calls on a (1, C) slice, each scaled by the speaker weight and divided by the batch
size -- 2B tiny forward launches (and 2B backward ones) per step.  The same number, to fp32 rounding of the summation
order, is

    sum_i w_i * ( CE(emotion_i) + gender_lambda * CE(gender_i) ) / B      -  scale_lamda * log(mean(sigma))

which is what `cloak_grl_loss` evaluates with a constant number of launches, independent of B, and which CUDA-graph
capture can record (no Python-side dict lookups or .item() calls inside).
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch
import torch.nn.functional as F


def speaker_weight_vector(weights: Mapping[str, float], speaker_ids: Sequence, datasets: Sequence[str], device=None) -> torch.Tensor:
    """Per-sample weights w_i = weights[f"{speaker}_{dataset}"] (the dict get_class_weight builds, reference :311-318,
    looked up per sample at :146) as a tensor, so the lookup happens once per batch on the host."""
    w = torch.tensor([float(weights[str(s) + "_" + d]) for s, d in zip(speaker_ids, datasets)], dtype=torch.float32)
    return w if device is None else w.to(device, non_blocking=True)


def weighted_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, w: torch.Tensor | None = None) -> torch.Tensor:
    """sum_i CE(logits_i, labels_i) * w_i / B  -- the reference's loop over one head (training_cloak.py:139-143)."""
    ce = F.cross_entropy(logits, labels.reshape(-1), reduction="none")
    if w is not None:
        ce = ce * w
    return ce.sum() / logits.shape[0]


def cloak_grl_loss(p_emo: torch.Tensor, p_gen: torch.Tensor, emo: torch.Tensor, gen: torch.Tensor, w: torch.Tensor | None,
                   gender_lambda: float, sigma: torch.Tensor | None = None, scale_lamda: float = 0.0) -> torch.Tensor:
    """training_cloak_with_grl.py:141-160.  w=None is the validate-mode branch (:151-154, no speaker weights);
    sigma = cloak_model.intermed.scales() adds the - scale_lamda * log(mean sigma) term of :158-160."""
    ce = F.cross_entropy(p_emo, emo.reshape(-1), reduction="none") + float(gender_lambda) * F.cross_entropy(p_gen, gen.reshape(-1), reduction="none")
    if w is not None:
        ce = ce * w
    total = ce.sum() / p_emo.shape[0]
    if sigma is not None and float(scale_lamda) != 0.0:
        total = total - float(scale_lamda) * torch.log(torch.mean(sigma))
    return total
