"""Generate tests/golden/*.npz by running the REAL reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container, where /root/reference (read-only) and torchaudio 2.11.0 exist:

    python oracle/make_golden.py

The reference's own tree holds no golden vectors or tests for this path, so the oracle is
pinned on outputs of the reference itself: its callables are imported unchanged (the three unused,
uninstalled top-level imports of audio_feature_extraction.py -- python_speech_features, moviepy.editor,
opensmile -- are stubbed in sys.modules) and run on small seeded inputs.  The fixtures travel to the GPU
box; /root/reference does not.
"""
from __future__ import annotations

import ast
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))

from speech_emotion_privacy_trust_b200 import synth  # noqa: E402


def import_reference():
    for name in ("python_speech_features", "moviepy", "moviepy.editor", "opensmile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    for sub in ("feature_extraction", "model", "utils"):
        sys.path.insert(0, str(REF / sub))
    import audio_feature_extraction as afe
    import baseline_models
    import cloak_models
    import reversal_gradient
    return afe, cloak_models, reversal_gradient, baseline_models


def extraction_fixture(afe):
    rng = np.random.default_rng(20261018)
    lengths = [801, 1600, 4000, 7777, 16000]
    waves = [synth.speech_shaped(n, rng) for n in lengths]
    waves.append(np.zeros(1000, dtype=np.float32))                        # silence: the 1e-10 floor
    waves.append((0.25 * np.sin(2 * np.pi * 440.0 * np.arange(3000) / 16000)).astype(np.float32))  # pure tone
    out = {"n_utts": np.int64(len(waves))}
    torch.set_num_threads(1)
    for i, w in enumerate(waves):
        a = torch.from_numpy(w)[None]
        out[f"wav{i}"] = w
        out[f"mel1_{i}"] = afe.mel_spectrogram(a, n_fft=800, feature_len=128).numpy()
        out[f"mel2_{i}"] = afe.mel_spectrogram(a, n_fft=1600, feature_len=128).numpy()
        out[f"mfcc_{i}"] = afe.mfcc(a)
    np.savez_compressed(OUT / "extraction.npz", **out)
    print("extraction.npz:", {k: v.shape for k, v in out.items() if k.startswith(("mel1", "mfcc"))})


def constants_fixture():
    import torchaudio.functional as AF
    out = {}
    for n_fft in (400, 800, 1600):
        fb = AF.melscale_fbanks(n_fft // 2 + 1, 0.0, 8000.0, 128, 16000, None, "htk").numpy()
        k, m = np.nonzero(fb)
        out[f"fb{n_fft}_k"] = k.astype(np.int32)
        out[f"fb{n_fft}_m"] = m.astype(np.int32)
        out[f"fb{n_fft}_v"] = fb[k, m]
    out["dct"] = AF.create_dct(40, 128, "ortho").numpy()
    out["hann800"] = torch.hann_window(800).numpy()
    np.savez_compressed(OUT / "constants.npz", **out)
    print("constants.npz nnz:", {n: len(out[f"fb{n}_v"]) for n in (400, 800, 1600)})


def cloak_fixture(cloak_models, reversal_gradient):
    torch.manual_seed(7)
    B, W, F = 5, 12, 16
    out = {}
    for tag, use_mask in (("nomask", False), ("mask", True)):
        locs0 = 0.05 * torch.randn(1, W, F)
        layer = cloak_models.cloak_noise(locs0, torch.ones(1, W, F), 0.01, 10.0, "cpu")
        with torch.no_grad():
            layer.rhos.add_(0.8 * torch.randn(1, W, F))
        eps = 0.1 * torch.randn(1, W, F)
        layer.normal.sample = lambda shape, _e=eps: _e.clone()      # "eps supplied externally"
        mask = (torch.rand(1, W, F) > 0.3).float() if use_mask else None
        x = torch.randn(B, 1, W, F, requires_grad=True)
        g_a, g_b, lam = torch.randn(B, 1, W, F), torch.randn(B, 1, W, F), 0.1
        y = layer(x, mask) if use_mask else layer(x)
        y_rev = reversal_gradient.GradientReversalFunction.apply(y, lam)
        ((y * g_a).sum() + (y_rev * g_b).sum()).backward()
        out.update({f"{tag}_x": x.detach().numpy(), f"{tag}_locs": locs0.numpy(), f"{tag}_rhos": layer.rhos.detach().numpy(),
                    f"{tag}_eps": eps.numpy(), f"{tag}_g_a": g_a.numpy(), f"{tag}_g_b": g_b.numpy(),
                    f"{tag}_out": y.detach().numpy(), f"{tag}_sigma": layer.scales().detach().numpy(),
                    f"{tag}_dlocs": layer.locs.grad.numpy(), f"{tag}_drhos": layer.rhos.grad.numpy(),
                    f"{tag}_dx": x.grad.numpy()})
        if use_mask:
            out["mask_mask"] = mask.numpy()
    out["lambda"] = np.float32(0.1)
    out["min_scale"], out["max_scale"] = np.float32(0.01), np.float32(10.0)
    # stand-alone gradient reversal
    g = torch.randn(3, 1, 7, 9)
    z = torch.randn(3, 1, 7, 9, requires_grad=True)
    reversal_gradient.GradientReversal(0.37)(z).backward(g)
    out["grl_g"], out["grl_dx"], out["grl_lambda"] = g.numpy(), z.grad.numpy(), np.float32(0.37)
    np.savez_compressed(OUT / "cloak.npz", **out)
    print("cloak.npz keys:", len(out))


def model_keys_fixture(cloak_models, baseline_models):
    """state_dict key -> shape of the classes whose checkpoints are loaded strictly
    (training_cloak_with_grl.py:395,403)."""
    def keys(m):
        return {k: list(v.shape) for k, v in m.state_dict().items()}
    spec = {}
    for cls in ("two_d_cnn_lstm", "deep_two_d_cnn_lstm"):
        for att in (None, "self_att"):
            m = getattr(baseline_models, cls)(1, 128, 5, lstm_hidden_size=64, num_layers_lstm=2, pred="emotion",
                                              bidirectional=True, rnn_cell="gru", global_feature=0, att=att)
            spec[f"{cls}|att={att}"] = keys(m)
    mk = lambda pred: baseline_models.two_d_cnn_lstm(1, 128, 5, lstm_hidden_size=64, pred=pred, global_feature=0)
    noise = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, "cpu")
    spec["two_d_cnn_lstm_syn"] = keys(cloak_models.two_d_cnn_lstm_syn(mk("emotion"), noise))
    noise = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, "cpu")
    spec["two_d_cnn_lstm_syn_with_grl"] = keys(
        cloak_models.two_d_cnn_lstm_syn_with_grl(mk("emotion"), mk("gender"), noise, 0.1))
    (OUT / "state_dict_keys.json").write_text(json.dumps(spec, indent=0, sort_keys=True))
    print("state_dict_keys.json:", {k: len(v) for k, v in spec.items()})


NORM_F = 8   # the block is feature-width agnostic; 8 keeps the fixture small


def norm_fixture():
    """Runs the reference's own write_data_dict/save_data_dict (preprocess_adversary_data.py:20-83), lifted
    out of the script by AST (it has no importable entry point), then the numpy calls of :358-381."""
    src = (REF / "preprocess_data" / "preprocess_adversary_data.py").read_text()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("write_data_dict", "save_data_dict")]
    import pandas as pd
    ns = {"np": np, "pd": pd, "win_len": 200, "shift_len": 50, "args": types.SimpleNamespace(shift="1", aug=None),
          "training_norm_dict": {}, "training_global_norm_dict": {}, "global_data": np.zeros((1, 88)),
          "training_dict": {}, "valid_dict": {}, "adv_training_dict": {}, "adv_valid_dict": {}, "test_dict": {},
          "train_speaker_id_arr": ["s0"], "validation_speaker_id_arr": ["s1"], "adv_train_speaker_id_arr": [],
          "adv_validation_speaker_id_arr": [], "test_speaker_id_arr": ["s2"], "train_label_list": []}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "ref_preprocess_functions", "exec"), ns)
    rng = np.random.default_rng(99)
    frames = [120, 200, 260, 431, 777, 203, 90, 350]
    spk = ["s0", "s1", "s0", "s2", "s1", "s0", "s2", "s1"]
    stats_dict = {k: {"neu": 0} for k in ("training", "valid", "adv_train", "adv_valid", "test")}
    out = {"n_utts": np.int64(len(frames)), "speakers": np.array(spk)}
    for u, (T, s) in enumerate(zip(frames, spk)):
        feat = (rng.standard_normal((T, NORM_F)) * 12.0 - 30.0 + 3.0 * (u % 3)).astype(np.float32)
        out[f"feat{u}"] = feat
        ns["sentence_file"] = f"utt{u}"
        ns["save_data_dict"](feat, stats_dict, "neu", "F", s)
    for s, lst in ns["training_norm_dict"].items():
        a = np.array(lst).reshape(-1, NORM_F)
        out[f"{s}_count"] = np.int64(a.shape[0])
        out[f"{s}_mean"], out[f"{s}_std"] = np.nanmean(a, axis=0), np.nanstd(a, axis=0)
        out[f"{s}_min"], out[f"{s}_max"] = np.nanmin(a, axis=0), np.nanmax(a, axis=0)
    n_out = 0
    for split in ("training_dict", "valid_dict", "test_dict"):
        for key, d in ns[split].items():
            st = {k: out[f"{d['speaker_id']}_{k}"] for k in ("mean", "std")}
            z = (d["data"] - st["mean"]) / (st["std"] + 1e-5)          # :378
            out[f"z|{key}"] = np.asarray(z, np.float64)
            n_out += 1
    np.savez_compressed(OUT / "norm.npz", **out)
    print("norm.npz: windows", n_out, "speakers", list(ns["training_norm_dict"]))


def resample_fixture():
    """torchaudio.transforms.Resample(44100, 16000), the call of audio_feature_extraction.py:140-141, on synthetic
    44.1 kHz audio (third-party dependency of the reference; version 2.11.0 in this container)."""
    import torchaudio
    rng = np.random.default_rng(4410)
    out = {}
    for i, n in enumerate((441, 4410, 10007, 30000)):
        w = synth.speech_shaped(n, rng)
        out[f"in{i}"] = w
        out[f"out{i}"] = torchaudio.transforms.Resample(44100, 16000)(torch.from_numpy(w)[None])[0].numpy()
    w = synth.speech_shaped(9000, rng)
    out["in_48k"] = w
    out["out_48k"] = torchaudio.transforms.Resample(48000, 16000)(torch.from_numpy(w)[None])[0].numpy()
    np.savez_compressed(OUT / "resample.npz", **out)
    print("resample.npz:", {k: v.shape for k, v in out.items() if k.startswith("out")})


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    afe, cloak_models, reversal_gradient, baseline_models = import_reference()
    extraction_fixture(afe)
    constants_fixture()
    cloak_fixture(cloak_models, reversal_gradient)
    model_keys_fixture(cloak_models, baseline_models)
    norm_fixture()
    resample_fixture()
