"""GPU parity of the per-speaker statistics / normalisation / window kernels against the golden fixture produced by the
reference's own write_data_dict/save_data_dict + numpy block (oracle/make_golden.py: norm_fixture)."""
import numpy as np
import pytest
import torch

from oracle import norm as onorm

pytestmark = pytest.mark.gpu


def _load(golden_norm):
    g = golden_norm
    n = int(g["n_utts"])
    feats = [g[f"feat{u}"] for u in range(n)]
    spk = [str(s) for s in g["speakers"]]
    return g, feats, spk


def _layout(feats):
    from speech_emotion_privacy_trust_b200.extraction import Layout
    fo = np.concatenate([[0], np.cumsum([len(f) for f in feats])]).astype(np.int64)
    return Layout(fo, torch.from_numpy(fo).cuda(), torch.zeros(len(fo), dtype=torch.int32, device="cuda"))


def test_speaker_stats_and_znorm_golden(golden_norm):
    from speech_emotion_privacy_trust_b200 import normalization as nz
    g, feats, spk = _load(golden_norm)
    whole = [s == "s2" for s in spk]                     # s2 is the test-split speaker of the fixture
    lay = _layout(feats)
    feat = torch.from_numpy(np.concatenate(feats)).cuda()
    st = nz.speaker_stats(feat, lay, spk, whole)
    d = st.as_dict()
    st64 = onorm.speaker_stats_f64(feats, spk, whole)
    for s in ("s0", "s1", "s2"):
        assert int(d[s]["count"][0]) == int(g[f"{s}_count"])
        for k in ("mean", "std", "min", "max"):
            # the reference's float32 numpy reduction is itself ~5e-5 from exact; both must sit within 2e-4 of each
            # other, and the kernel within 2e-5 of the fp64 ground truth
            assert np.max(np.abs(d[s][k] - g[f"{s}_{k}"])) < 2e-4, (s, k)
            assert np.max(np.abs(d[s][k] - st64[s][k])) < 2e-5, (s, k)
    # training windows of utterances 0 (short: zero padded before normalisation), 2 and 5
    wu, wt = nz.window_table(lay, [0, 2, 5])
    assert list(wu) == [0, 2, 2, 5] and list(wt) == [0, 0, 50, 0]
    win = nz.normalized_windows(feat, lay, st, wu, wt).cpu().numpy()
    assert win.shape == (4, 1, 200, feats[0].shape[1])
    for j, (u, t0) in enumerate(zip(wu, wt)):
        ref = np.asarray(g[f"z|utt{u}_{t0 // 50}"]).reshape(-1, feats[0].shape[1])
        assert len(ref) == 200
        assert np.max(np.abs(win[j, 0] - ref)) < 1e-4, (u, t0)
    # whole utterances, frame-wise
    z = nz.normalize(feat, lay, st).cpu().numpy()
    fo = lay.frame_off_host
    for u in (3, 6):                                      # test-split utterances are stored whole
        ref = np.asarray(g[f"z|utt{u}_0"]).reshape(-1, feats[0].shape[1])
        T = fo[u + 1] - fo[u]                           # a short test utterance is stored zero padded to 200 rows (:29-35)
        assert np.max(np.abs(z[fo[u]:fo[u + 1]] - ref[:T])) < 1e-4


def test_min_max_and_weights_vs_oracle():
    from speech_emotion_privacy_trust_b200 import normalization as nz
    rng = np.random.default_rng(5)
    frames = [199, 200, 249, 250, 251, 1001, 37]
    spk = ["a", "b", "a", "b", "a", "b", "a"]
    feats = [(rng.standard_normal((T, 128)) * 9 - 40).astype(np.float32) for T in frames]
    lay = _layout(feats)
    feat = torch.from_numpy(np.concatenate(feats)).cuda()
    st = nz.speaker_stats(feat, lay, spk)
    ref = onorm.speaker_stats_f64(feats, spk)
    d = st.as_dict()
    for s in ("a", "b"):
        n_ref = sum(onorm.frame_multiplicity(T).sum() for T, q in zip(frames, spk) if q == s)
        assert int(d[s]["count"][0]) == n_ref
        for k in ("mean", "std", "min", "max"):
            assert np.max(np.abs(d[s][k] - ref[s][k])) < 5e-5, (s, k)
    z = nz.normalize(feat, lay, st, mode="min_max").cpu().numpy()
    fo = lay.frame_off_host
    for u in range(len(frames)):
        want = onorm.normalize(feats[u].astype(np.float64), ref[spk[u]], "min_max")
        assert np.max(np.abs(z[fo[u]:fo[u + 1]] - want)) < 1e-5
    assert nz.n_windows(199) == 1 and nz.n_windows(250) == 2 and nz.n_windows(1001) == 17
