"""Per-speaker normalisation + training-window assembly on the device (csrc/norm.cu via the C ABI).

The reference does this inline in preprocess_data/preprocess_adversary_data.py (:26-27 frame accumulation, :41-48 window
rule, :29-35 zero padding, :357-385 statistics and normalisation); there is no callable to mirror, so this module
defines the API.  Inputs are the frame-major features extraction.logmel() returns.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .extraction import Layout, _stream

WIN_LEN, SHIFT_LEN = 200, 50      # training_data_preprocess.sh:6-8 ; preprocess_adversary_data.py:131
MODES = {"znorm": 0, "min_max": 1}


@dataclass
class SpeakerStats:
    speakers: list                 # speaker labels, row order of `stats`
    stats: torch.Tensor            # device (n_spk, 5, F): count, mean, std (ddof 0), min, max
    spk_of_utt: torch.Tensor       # device int32 [n_utts] row index per utterance

    def as_dict(self) -> dict:
        h = self.stats.cpu().numpy()
        return {s: {"count": h[i, 0], "mean": h[i, 1], "std": h[i, 2], "min": h[i, 3], "max": h[i, 4]}
                for i, s in enumerate(self.speakers)}


def n_windows(n_frames: int, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN) -> int:
    return 1 if n_frames < win_len else (n_frames - win_len) // shift_len + 1


def speaker_stats(feat: torch.Tensor, lay: Layout, speaker_of_utt: Sequence, whole_utterance: Sequence[bool] | None = None,
                  win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN) -> SpeakerStats:
    """Weighted per-speaker mean/std/min/max: a frame inside k training windows counts k times (:26-27), utterances
    flagged whole (test split, :56-60) count each frame once."""
    _lib.require_cuda(feat)
    if feat.dtype != torch.float32 or feat.dim() != 2 or not feat.is_contiguous():
        raise ValueError("feat must be a contiguous (total_frames, F) float32 tensor")
    n_utts = len(lay.frame_off_host) - 1
    if len(speaker_of_utt) != n_utts:
        raise ValueError("one speaker label per utterance is required")
    speakers = list(dict.fromkeys(speaker_of_utt))
    row = {s: i for i, s in enumerate(speakers)}
    spk_idx = np.fromiter((row[s] for s in speaker_of_utt), dtype=np.int32, count=n_utts)
    order = np.argsort(spk_idx, kind="stable").astype(np.int32)
    ptr = np.zeros(len(speakers) + 1, dtype=np.int32)
    np.cumsum(np.bincount(spk_idx, minlength=len(speakers)), out=ptr[1:])
    dev = feat.device
    F = feat.shape[1]
    d_ptr, d_order = torch.from_numpy(ptr).to(dev), torch.from_numpy(order).to(dev)
    d_whole = None if whole_utterance is None else torch.from_numpy(np.asarray(whole_utterance, dtype=np.uint8)).to(dev)
    partial = torch.empty((n_utts, 5, F), dtype=torch.float32, device=dev)
    stats = torch.empty((len(speakers), 5, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sept_speaker_stats_f32(
            feat.data_ptr(), lay.frame_off.data_ptr(), 0 if d_whole is None else d_whole.data_ptr(), n_utts, F, win_len,
            shift_len, d_ptr.data_ptr(), d_order.data_ptr(), len(speakers), partial.data_ptr(), stats.data_ptr(), _stream(dev)))
    return SpeakerStats(speakers, stats, torch.from_numpy(spk_idx).to(dev))


def normalize(feat: torch.Tensor, lay: Layout, st: SpeakerStats, mode: str = "znorm") -> torch.Tensor:
    """Frame-wise normalisation with each utterance's speaker statistics (:377-381); same shape as feat."""
    _lib.require_cuda(feat)
    out = torch.empty_like(feat)
    dev = feat.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sept_normalize_f32(feat.data_ptr(), lay.frame_off.data_ptr(), st.spk_of_utt.data_ptr(),
                                                 st.stats.data_ptr(), len(lay.frame_off_host) - 1, feat.shape[1], MODES[mode],
                                                 out.data_ptr(), _stream(dev)))
    return out


def window_table(lay: Layout, utts: Sequence[int] | None = None, win_len: int = WIN_LEN, shift_len: int = SHIFT_LEN):
    """(win_utt, win_t0) int32 arrays of every training window of the given utterances (:44-48)."""
    fo = lay.frame_off_host
    utts = range(len(fo) - 1) if utts is None else utts
    wu, wt = [], []
    for u in utts:
        k = n_windows(int(fo[u + 1] - fo[u]), win_len, shift_len)
        wu.append(np.full(k, u, dtype=np.int32))
        wt.append(np.arange(k, dtype=np.int32) * shift_len)
    if not wu:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    return np.concatenate(wu), np.concatenate(wt)


def normalized_windows(feat: torch.Tensor, lay: Layout, st: SpeakerStats, win_utt: np.ndarray, win_t0: np.ndarray,
                       mode: str = "znorm", win_len: int = WIN_LEN) -> torch.Tensor:
    """(n_windows, 1, win_len, F) float32 batches for the cloak layer: normalised windows, short utterances zero-padded
    BEFORE normalisation like :29-35."""
    _lib.require_cuda(feat)
    dev = feat.device
    n = len(win_utt)
    F = feat.shape[1]
    out = torch.empty((n, 1, win_len, F), dtype=torch.float32, device=dev)
    d_wu = torch.from_numpy(np.ascontiguousarray(win_utt, dtype=np.int32)).to(dev)
    d_wt = torch.from_numpy(np.ascontiguousarray(win_t0, dtype=np.int32)).to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sept_normalize_windows_f32(
            feat.data_ptr(), lay.frame_off.data_ptr(), st.spk_of_utt.data_ptr(), st.stats.data_ptr(), d_wu.data_ptr(),
            d_wt.data_ptr(), n, win_len, F, MODES[mode], out.data_ptr(), _stream(dev)))
    return out


def speaker_stats_distributed(feat: torch.Tensor, lay: Layout, speaker_of_utt: Sequence, all_speakers: Sequence,
                              whole_utterance: Sequence[bool] | None = None, group=None, win_len: int = WIN_LEN,
                              shift_len: int = SHIFT_LEN) -> SpeakerStats:
    """Per-speaker statistics when a speaker's utterances are spread over several ranks (they are under length-bucketed
    sharding): every rank reduces its own utterances with the kernels, then ONE all-gather of the
    (count, mean, M2, min, max) partials and a Chan merge in rank order give every rank the global statistics.
    `all_speakers` fixes the row order on every rank (speakers a rank does not hold contribute empty partials)."""
    from . import parallel
    local = speaker_stats(feat, lay, speaker_of_utt, whole_utterance, win_len, shift_len)
    dev, F = feat.device, feat.shape[1]
    row = {s: i for i, s in enumerate(all_speakers)}
    n_spk = len(all_speakers)
    cnt = torch.zeros((n_spk, F), dtype=torch.float32, device=dev)
    mean, m2 = torch.zeros_like(cnt), torch.zeros_like(cnt)
    mn, mx = torch.full_like(cnt, float("inf")), torch.full_like(cnt, float("-inf"))
    if local.speakers:
        idx = torch.tensor([row[s] for s in local.speakers], device=dev)
        st = local.stats
        cnt[idx], mean[idx], m2[idx] = st[:, 0], st[:, 1], st[:, 2] * st[:, 2] * st[:, 0]
        mn[idx], mx[idx] = st[:, 3], st[:, 4]
    n, mu, sd, lo, hi = parallel.merge_speaker_partials(cnt, mean, m2, mn, mx, group=group)
    stats = torch.stack([n, mu, sd, lo, hi], dim=1).contiguous()
    spk_idx = torch.tensor([row[s] for s in speaker_of_utt], dtype=torch.int32, device=dev)
    return SpeakerStats(list(all_speakers), stats, spk_idx)
