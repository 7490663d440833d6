"""Pin the oracle (oracle/) against golden vectors produced by the REAL reference
(oracle/make_golden.py, run where /root/reference exists).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cloak as ocloak
from oracle import norm as onorm
from oracle import ref_port, restate

N_UTTS = 7


@pytest.mark.parametrize("i", range(N_UTTS))
def test_ref_port_is_bit_identical_to_reference(golden_extraction, i):
    torch.set_num_threads(1)
    wav = torch.from_numpy(golden_extraction[f"wav{i}"])[None]
    assert np.array_equal(ref_port.mel_spectrogram(wav, 800, 128).numpy(), golden_extraction[f"mel1_{i}"])
    assert np.array_equal(ref_port.mel_spectrogram(wav, 1600, 128).numpy(), golden_extraction[f"mel2_{i}"])
    got = ref_port.mfcc(wav)
    assert got.dtype == np.float32 and np.array_equal(got, golden_extraction[f"mfcc_{i}"])


@pytest.mark.parametrize("i", range(5))          # speech-shaped utterances
@pytest.mark.parametrize("key,n_fft", [("mel1", 800), ("mel2", 1600)])
def test_restatement_logmel_within_tolerance(golden_extraction, i, key, n_fft):
    """north_star tolerance: log-mel within 1e-3 dB max abs (fp64 restatement vs fp32 reference)."""
    wav = golden_extraction[f"wav{i}"][None]
    ref = golden_extraction[f"{key}_{i}"]
    got = restate.mel_spectrogram(wav, n_fft, 128, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 1e-3
    got32 = restate.mel_spectrogram(wav, n_fft, 128, dtype=np.float32)
    assert np.max(np.abs(got32 - ref)) < 1e-3


@pytest.mark.parametrize("i", range(5))
def test_restatement_mfcc_within_tolerance(golden_extraction, i):
    """north_star tolerance: MFCC within 1e-4 relative (to the utterance's max |coefficient|)."""
    wav = golden_extraction[f"wav{i}"][None]
    ref = golden_extraction[f"mfcc_{i}"]
    got = restate.mfcc(wav, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-4


def test_restatement_silence_hits_floor(golden_extraction):
    ref = golden_extraction["mel1_5"]
    assert np.all(ref == -100.0)
    got = restate.mel_spectrogram(golden_extraction["wav5"][None], 800, 128, dtype=np.float32)
    assert np.max(np.abs(got + 100.0)) < 1e-4      # numpy's float32 log10 lands 1 ulp off torch's here


@pytest.mark.parametrize("n_fft", [400, 800, 1600])
def test_restatement_filterbank(golden_constants, n_fft):
    fb_ref = np.zeros((n_fft // 2 + 1, 128), np.float32)
    fb_ref[golden_constants[f"fb{n_fft}_k"], golden_constants[f"fb{n_fft}_m"]] = golden_constants[f"fb{n_fft}_v"]
    fb64 = restate.melscale_fbanks_htk(n_fft // 2 + 1, 128, dtype=np.float64)
    assert np.max(np.abs(fb64 - fb_ref)) < 2e-5
    assert np.array_equal(fb64 > 1e-6, fb_ref > 1e-6)
    # the identically-zero filters of the 201-bin bank
    if n_fft == 400:
        assert [m for m in range(128) if not fb_ref[:, m].any()] == [0, 3, 6, 13]


def test_restatement_dct_and_window(golden_constants):
    # torch evaluates both tables with float32 angles: its own tables sit ~1e-6 / 3e-7 off the exact formula
    assert np.max(np.abs(restate.create_dct_ortho(40, 128) - golden_constants["dct"])) < 2e-6
    assert np.max(np.abs(restate.hann_periodic(800) - golden_constants["hann800"])) < 5e-7


@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_cloak_oracle(golden_cloak, tag):
    g = golden_cloak
    mask = g["mask_mask"] if tag == "mask" else None
    mn, mx = float(g["min_scale"]), float(g["max_scale"])
    assert np.max(np.abs(ocloak.scales(g[f"{tag}_rhos"], mn, mx) - g[f"{tag}_sigma"])) < 1e-6
    out = ocloak.forward(g[f"{tag}_x"], g[f"{tag}_locs"], g[f"{tag}_rhos"], g[f"{tag}_eps"], mn, mx, mask)
    assert np.max(np.abs(out - g[f"{tag}_out"])) < 1e-6
    dlocs, drhos, dx = ocloak.backward(g[f"{tag}_g_a"], g[f"{tag}_rhos"], g[f"{tag}_eps"], mn, mx, mask,
                                       g_b=g[f"{tag}_g_b"], lambda_=float(g["lambda"]))
    assert np.max(np.abs(dlocs - g[f"{tag}_dlocs"])) < 1e-5
    assert np.max(np.abs(drhos - g[f"{tag}_drhos"])) < 1e-5
    assert np.max(np.abs(dx - g[f"{tag}_dx"])) < 1e-6


def test_grl_oracle(golden_cloak):
    g = golden_cloak
    assert np.array_equal(ocloak.grl_backward(g["grl_g"], float(g["grl_lambda"])), g["grl_dx"])


def test_norm_oracle(golden_norm):
    g = golden_norm
    n = int(g["n_utts"])
    feats = [g[f"feat{u}"] for u in range(n)]
    spk = [str(s) for s in g["speakers"]]
    whole = [s == "s2" for s in spk]                       # s2 is the test-split speaker in the fixture
    st = onorm.speaker_stats(feats, spk, whole)
    st64 = onorm.speaker_stats_f64(feats, spk, whole)
    for s in ("s0", "s1", "s2"):
        for k in ("mean", "std", "min", "max"):
            assert np.array_equal(st[s][k], g[f"{s}_{k}"]), (s, k)
            # numpy reduces axis 0 of float32 rows sequentially: the reference's own statistics sit
            # ~5e-5 off the exact value at this size; the fp64 restatement is the ground truth
            assert np.max(np.abs(st64[s][k] - g[f"{s}_{k}"])) < 2e-4, (s, k)
    # windows: training-split utterance 2 (260 frames -> 2 windows), short utterance 0 (padded)
    for u in (0, 2, 5):
        for i, w in enumerate(onorm.windows_of(feats[u])):
            z = onorm.normalize(w, st[spk[u]])
            assert np.allclose(z, g[f"z|utt{u}_{i}"], rtol=0, atol=1e-12)
    assert onorm.n_windows(199) == 1 and onorm.n_windows(200) == 1 and onorm.n_windows(250) == 2
    assert list(onorm.frame_multiplicity(300)[[0, 49, 50, 99, 100, 199, 200, 299]]) == [1, 1, 2, 2, 3, 3, 2, 1]
