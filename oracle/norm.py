"""numpy restatement of windowing + per-speaker normalisation -- TEST INFRASTRUCTURE ONLY.

The reference has no callable for this: it is an inline block of
preprocess_data/preprocess_adversary_data.py (:41-48 window count, :26-27 frame accumulation,
:29-35 NaN/zero padding of short utterances, :357-385 statistics and normalisation).  Restated
here as functions; parity is unpinned by any reference test and anchored instead on
oracle/make_golden.py running the same numpy calls the reference makes.
"""
from __future__ import annotations

import numpy as np

WIN_LEN = 200    # training_data_preprocess.sh:6-8  --win_len 200
SHIFT_LEN = 50   # preprocess_adversary_data.py:131


def n_windows(n_frames: int, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN) -> int:
    """:44-45 with --shift 1."""
    return 1 if n_frames < win_len else int((n_frames - win_len) / shift_len) + 1


def frame_multiplicity(n_frames: int, whole_utterance: bool = False, win_len: int = WIN_LEN,
                       shift_len: int = SHIFT_LEN) -> np.ndarray:
    """How many times each frame of one utterance is appended to its speaker's list (:26-27):
    once per window that contains it; test-split utterances are appended whole, once (:56-60)."""
    if whole_utterance or n_frames < win_len:
        return np.ones(n_frames, dtype=np.int64)
    mult = np.zeros(n_frames, dtype=np.int64)
    for i in range(n_windows(n_frames, win_len, shift_len)):
        mult[i * shift_len:i * shift_len + win_len] += 1
    return mult


def speaker_stats(feats, speaker_of_utt, whole_utterance=None, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN):
    """feats: list of (T_u, F) float32.  Returns {speaker: dict(mean,std,min,max)} exactly as :358-364:
    every window's rows are stacked (duplicates included) and reduced by numpy in float32."""
    rows = {}
    for u, f in enumerate(feats):
        whole = bool(whole_utterance[u]) if whole_utterance is not None else False
        lst = rows.setdefault(speaker_of_utt[u], [])
        T = len(f)
        if whole or T < win_len:
            lst.extend(f[i] for i in range(T))
        else:
            for i in range(n_windows(T, win_len, shift_len)):
                w = f[i * shift_len:i * shift_len + win_len]
                lst.extend(w[j] for j in range(len(w)))
    out = {}
    for s, lst in rows.items():
        a = np.array(lst).reshape(-1, feats[0].shape[1])
        out[s] = {"mean": np.nanmean(a, axis=0), "std": np.nanstd(a, axis=0),
                  "min": np.nanmin(a, axis=0), "max": np.nanmax(a, axis=0)}
    return out


def speaker_stats_f64(feats, speaker_of_utt, whole_utterance=None, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN):
    """Same statistics from multiplicity weights in float64 (ground truth for the reductions)."""
    acc = {}
    for u, f in enumerate(feats):
        whole = bool(whole_utterance[u]) if whole_utterance is not None else False
        w = frame_multiplicity(len(f), whole, win_len, shift_len).astype(np.float64)
        a = acc.setdefault(speaker_of_utt[u], {"n": 0.0, "s": 0.0, "ss": 0.0, "mn": np.inf, "mx": -np.inf})
        x = np.asarray(f, np.float64)
        a["n"] += w.sum()
        a["s"] = a["s"] + (w[:, None] * x).sum(0)
        a["ss"] = a["ss"] + (w[:, None] * x * x).sum(0)
        used = x[w > 0]
        if len(used):
            a["mn"] = np.minimum(a["mn"], used.min(0))
            a["mx"] = np.maximum(a["mx"], used.max(0))
    out = {}
    for s, a in acc.items():
        mean = a["s"] / a["n"]
        var = np.maximum(a["ss"] / a["n"] - mean * mean, 0.0)
        out[s] = {"mean": mean, "std": np.sqrt(var), "min": a["mn"], "max": a["mx"]}
    return out


def normalize(x, st, mode: str = "znorm"):
    """:377-381.  x (len, F); st one speaker's dict."""
    if mode == "znorm":
        return (x - st["mean"]) / (st["std"] + 1e-5)
    if mode == "min_max":
        return (x - st["min"]) / (st["max"] - st["min"]) * 2 - 1
    raise ValueError(mode)


def windows_of(feat, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN):
    """Training-split windows of one utterance (:55-81 + :29-35): list of (win_len, F) arrays; a short
    utterance yields one window zero-padded to win_len (the padding is applied BEFORE normalisation,
    so padded rows become (0 - mean) / (std + 1e-5))."""
    T, F = feat.shape
    if T < win_len:
        w = np.zeros((win_len, F), dtype=np.float64)
        w[:T] = feat
        return [w]
    return [feat[i * shift_len:i * shift_len + win_len] for i in range(n_windows(T, win_len, shift_len))]
