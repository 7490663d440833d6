"""Harness that drives the reference's OWN train() / test() functions -- TEST INFRASTRUCTURE ONLY.

The caller scripts under tests/ref_fixture/training/ are verbatim copies of the reference's (tests/ref_fixture/vendor.py).
They are imported as modules -- their `__main__` blocks never run -- once with the reference's model/ directory on
sys.path ("reference") and once with the drop-in directory ahead of it ("dropin"), so the very same `train()` /
`test()` bodies run against both implementations.  The harness supplies what the scripts' `__main__` would have set up,
working around the latent defects SURVEY.md Appendix B lists WITHOUT editing the scripts:

  * an 8-tuple collate (the scripts read sampled_batch[7] = speaker id; the shipped collate returns 6 items),
  * data dicts that carry 'dataset' and 'speaker_id', run with args.dataset = 'combine',
  * the module-level globals the functions read: weights, cloak_model, scheduler, baseline_model, adversary_model,
  * np.Inf (removed in NumPy 2) for EarlyStopping,
  * an argparse-free Namespace with the types the function bodies expect.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from argparse import Namespace
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
FIXTURE = HERE / "ref_fixture"
REPO = HERE.parent
FLAT_MODULES = ("baseline_models", "cloak_models", "reversal_gradient", "training_tools", "audio_feature_extraction")
EMO = ["neu", "hap", "sad", "ang"]
GEN = ["F", "M"]
DATASETS = ["iemocap", "crema-d", "msp-improv"]

if not hasattr(np, "Inf"):
    np.Inf = np.inf                      # utils/training_tools.py:104


def load_driver(script: str, variant: str) -> types.ModuleType:
    """Import tests/ref_fixture/training/<script>.py as a module whose flat imports (`from cloak_models import ...`)
    resolve to the reference's model/ directory (variant 'reference') or to the drop-ins (variant 'dropin')."""
    assert variant in ("reference", "dropin")
    saved_path = list(sys.path)
    saved_mods = {m: sys.modules.pop(m) for m in FLAT_MODULES if m in sys.modules}
    try:
        sys.path.insert(0, str(FIXTURE / "utils"))
        sys.path.insert(0, str(FIXTURE / "model"))
        if variant == "dropin":
            if str(REPO) not in sys.path:
                sys.path.insert(0, str(REPO))
            from speech_emotion_privacy_trust_b200 import dropin
            dropin.install()             # puts the drop-in directory at sys.path[0]
        name = f"refdriver_{script}_{variant}"
        spec = importlib.util.spec_from_file_location(name, FIXTURE / "training" / f"{script}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)     # runs the script's own import lines; `__name__` != '__main__'
        origin = Path(sys.modules["baseline_models"].__file__).resolve().parent
        want = FIXTURE / "model" if variant == "reference" else REPO / "speech_emotion_privacy_trust_b200" / "dropin"
        assert origin == want, (origin, want)
        mod.variant = variant
        mod.flat_origins = {m: Path(sys.modules[m].__file__).resolve().parent for m in FLAT_MODULES if m in sys.modules}
        return mod
    finally:
        sys.path[:] = saved_path
        for m in FLAT_MODULES:
            sys.modules.pop(m, None)
        sys.modules.update(saved_mods)


def make_args(**over) -> Namespace:
    """What the scripts' argparse produces, with the types their function bodies need (Appendix B: untyped argparse)."""
    base = dict(dataset="combine", feature_type="mel_spec", input_channel=1, input_spec_size="128", batch_size=8, aug="emotion",
                num_epochs=2, model_type="2d-cnn-lstm", pred="emotion", global_feature=0, norm="znorm", win_len=200,
                optimizer="sgd", shift=1, att=None, adv=0, suppression_ratio=0, scale_lamda=0, grl_lambda=0.1,
                gender_lambda=0.1, grl=1)
    base.update(over)
    return Namespace(**base)


def synthetic_split(n: int, seed: int, frames=(200, 200), shift: float = 0.5, n_speakers: int = 6, emo_shift: float = 1.0,
                    nuisance: float = 0.0) -> dict:
    """A data dict of the shape preprocess_adversary_data.py:20-38 writes (+ 'dataset', which combine_data adds :102):
    z-normed-scale windows with class-dependent mean shifts so that emotion and gender are learnable.
    `nuisance` adds a per-utterance random offset to every band (std `nuisance`): it overlaps the class shifts, so the
    classes are only partly separable and UAR / accuracy land strictly between chance and 1."""
    rng = np.random.default_rng(seed)
    out = {}
    for i in range(n):
        T = int(rng.integers(frames[0], frames[1] + 1))
        emo, gen = int(rng.integers(4)), int(rng.integers(2))
        spk = int(rng.integers(n_speakers))
        x = rng.standard_normal((1, T, 128))
        if nuisance:
            x += nuisance * rng.standard_normal((1, 1, 128))
        x[:, :, 20 + 10 * emo: 30 + 10 * emo] += emo_shift
        x[:, :, 0:20] += shift if gen == 1 else -shift
        out[f"utt{seed}_{i}"] = {"data": x.astype(np.float64), "global_data": np.zeros((1, 88)), "label": EMO[emo],
                                 "gender": GEN[gen], "speaker_id": spk, "dataset": DATASETS[i % 3]}   # every corpus present:
        # ReturnResultDict scores each of the three when dataset == 'combine' (training_tools.py:152-170)
    return out


def make_loader(mod, data: dict, batch_size: int, shuffle: bool, seed: int = 0):
    """The reference's SpeechDataGenerator + speech_collate, extended to the 8-tuple train() indexes (Appendix B row 1)."""
    base_cls, collate6 = mod.SpeechDataGenerator, mod.speech_collate

    class WithSpeaker(base_cls):
        def __getitem__(self, idx):
            sample = super().__getitem__(idx)
            sample["speaker_id"] = self.data_dict[self.dict_keys[idx]]["speaker_id"]
            return sample

    def collate8(batch):
        return (*collate6(batch), [None] * len(batch), [s["speaker_id"] for s in batch])

    ds = WithSpeaker(data, list(data.keys()), mode="train", input_channel=1)
    gen = torch.Generator().manual_seed(seed)
    return torch.utils.data.DataLoader(ds, batch_size=batch_size, num_workers=0, shuffle=shuffle, collate_fn=collate8, generator=gen)


def speaker_weights(mod_tools_get_class_weight, train: dict) -> dict:
    """training_cloak_with_grl.py:311-318."""
    w = {}
    for key in train:
        sid = str(train[key]["speaker_id"]) + "_" + train[key]["dataset"]
        w[sid] = w.get(sid, 0) + 1
    return mod_tools_get_class_weight(w)


def dropout_off(model: torch.nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
            m.p = 0.0
        if isinstance(m, torch.nn.RNNBase):
            m.dropout = 0.0


class EpsTape:
    """Replaces `noise_model.normal.sample` -- the reference's own draw site (cloak_models.py:47,49) -- with a seeded
    tape so both variants see the same eps sequence ("eps supplied externally")."""

    def __init__(self, seed: int, std: float = 0.1, ulp: bool = False):
        self.gen = torch.Generator().manual_seed(seed)
        self.std = std
        self.draws = 0
        self.ulp = ulp                   # shift every sample by one unit in the last place: a rounding-sized perturbation

    def __call__(self, shape):
        self.draws += 1
        eps = self.std * torch.randn(tuple(shape), generator=self.gen)
        return torch.nextafter(eps, torch.full_like(eps, float("inf"))) if self.ulp else eps


def build_grl_model(mod, device, state=None, model_type="2d-cnn-lstm", att=None, hidden=64, max_scale=10.0, grl_lambda=0.1):
    """The model section of training_cloak_with_grl.py:330-399 with the classes the driver module imported."""
    cls = mod.deep_two_d_cnn_lstm if "deep" in model_type else mod.two_d_cnn_lstm
    mk = lambda pred: cls(input_channel=1, input_spec_size=128, cnn_filter_size=64, pred=pred, lstm_hidden_size=hidden,
                          num_layers_lstm=2, attention_size=128, att=att, global_feature=0)
    mus = torch.zeros((1, 200, 128)).to(device)
    scale = torch.ones((1, 200, 128)).to(device)
    noise = mod.cloak_noise(mus, scale, torch.tensor(0.01).to(device), torch.tensor(max_scale).to(device), device).to(device)
    pre, gender = mk("emotion").to(device), mk("gender").to(device)
    model = mod.two_d_cnn_lstm_syn_with_grl(pre, gender, noise, float(grl_lambda)).to(device)
    if state is not None:
        model.load_state_dict(state, strict=True)
    return model


def run_grl_training(mod, device, train: dict, valid: dict, test: dict, *, state=None, epochs=2, batch_size=8, eps_seed=5,
                     loader_seed=3, model_type="2d-cnn-lstm", att=None, hidden=64, record=None, cloak_lr=0.001, scale_lamda=0,
                     eps_ulp=False):
    """epochs x [train(training), train(validate), test()] of training_cloak_with_grl.py:430-436 on the given splits.
    Returns (per-epoch result dicts, final model).  `record`, if a list, receives every (preds, preds_grl) the model
    returned in training mode, in call order."""
    args = make_args(batch_size=batch_size, num_epochs=epochs, model_type=model_type, att=att, scale_lamda=scale_lamda)
    model = build_grl_model(mod, device, state, model_type, att, hidden)
    dropout_off(model)
    tape = EpsTape(eps_seed, ulp=eps_ulp) if eps_seed is not None else None   # None: the layer's own sampler (CPU normal / device Philox)
    if tape is not None:
        model.intermed.normal.sample = tape
    mod.weights = speaker_weights(mod.get_class_weight, train)
    mod.cloak_model = model
    loss = torch.nn.CrossEntropyLoss().to(device)
    # :417 (SGD 1e-3, momentum 0.9, wd 1e-4); the cloak parameters may get their own learning rate so that a test of a few
    # dozen steps moves them by more than rounding noise (their gradients are ~1e-5 per element)
    cloak_params = list(model.intermed.parameters())
    others = [p for p in model.parameters() if p.requires_grad and all(p is not q for q in cloak_params)]
    optimizer = torch.optim.SGD([{"params": cloak_params, "lr": cloak_lr}, {"params": others}], lr=0.001, momentum=0.9, weight_decay=1e-4)
    mod.scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=10, gamma=0.5)
    hook = None
    if record is not None:
        def keep(_m, _inp, out):
            if model.training:
                record.append((out[0].detach().float().cpu().numpy().copy(), out[1].detach().float().cpu().numpy().copy()))
        hook = model.register_forward_hook(keep)
    results = []
    for epoch in range(epochs):
        dl_train = make_loader(mod, train, batch_size, True, loader_seed + epoch)
        dl_val = make_loader(mod, valid, batch_size, True, loader_seed + 100 + epoch)
        dl_test = make_loader(mod, test, 1, False)
        r_train = mod.train(model, device, dl_train, optimizer, loss, epoch, args, mode="training", pred=args.pred, mask=None)
        r_val = mod.train(model, device, dl_val, optimizer, loss, epoch, args, mode="validate", pred=args.pred, mask=None)
        r_test = mod.test(model, device, dl_test, optimizer, loss, epoch, args, pred=args.pred, mask=None)
        results.append({"train": r_train, "validate": r_val, "test": r_test})
    if hook is not None:
        hook.remove()
    return results, model, tape


def build_syn_model(mod, device, state=None, att=None, hidden=64, max_scale=10.0):
    """training_cloak.py:330-362: only the deep model is wired up there (Appendix B: pre_trained_model undefined else)."""
    pre = mod.deep_two_d_cnn_lstm(input_channel=1, input_spec_size=128, cnn_filter_size=64, pred="emotion", lstm_hidden_size=hidden,
                                  num_layers_lstm=2, attention_size=128, att=att, global_feature=0).to(device)
    mus = torch.zeros((1, 200, 128)).to(device)
    scale = torch.ones((1, 200, 128)).to(device)
    noise = mod.cloak_noise(mus, scale, torch.tensor(0.01).to(device), torch.tensor(max_scale).to(device), device).to(device)
    model = mod.two_d_cnn_lstm_syn(pre, noise).to(device)
    if state is not None:
        model.load_state_dict(state, strict=True)
    return model


def run_cloak_training(mod, device, train: dict, valid: dict, test: dict, *, state=None, epochs=1, batch_size=8, eps_seed=5,
                       loader_seed=3, att=None, hidden=64, record=None, cloak_lr=0.001, scale_lamda=0):
    """training_cloak.py's train()/test() (no GRL; two_d_cnn_lstm_syn over deep_two_d_cnn_lstm, pooling None)."""
    args = make_args(batch_size=batch_size, num_epochs=epochs, model_type="deep-2d-cnn-lstm", att=att, scale_lamda=scale_lamda)
    model = build_syn_model(mod, device, state, att, hidden)
    dropout_off(model)
    tape = EpsTape(eps_seed)
    model.intermed.normal.sample = tape
    mod.weights = speaker_weights(mod.get_class_weight, train)
    loss = torch.nn.CrossEntropyLoss().to(device)
    optimizer = torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=cloak_lr, momentum=0.9, weight_decay=1e-4)
    mod.scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=10, gamma=0.5)
    hook = None
    if record is not None:
        def keep(_m, _inp, out):
            if model.training:
                record.append(out[0].detach().float().cpu().numpy().copy())
        hook = model.register_forward_hook(keep)
    results = []
    for epoch in range(epochs):
        r_train = mod.train(model, device, make_loader(mod, train, batch_size, True, loader_seed + epoch), optimizer, loss, epoch, args,
                            mode="training", pred="emotion", mask=None)
        r_val = mod.train(model, device, make_loader(mod, valid, batch_size, True, loader_seed + 100 + epoch), optimizer, loss, epoch,
                          args, mode="validate", pred="emotion", mask=None)
        r_test = mod.test(model, device, make_loader(mod, test, 1, False), optimizer, loss, epoch, args, pred="emotion", mask=None)
        results.append({"train": r_train, "validate": r_val, "test": r_test})
    if hook is not None:
        hook.remove()
    return results, model, tape


def run_baseline_training(mod, device, train: dict, valid: dict, test: dict, *, pred="emotion", model_type="2d-cnn-lstm", att=None,
                          hidden=64, epochs=3, batch_size=8, lr=5e-4, loader_seed=11):
    """training_adversary_baselines.py's train()/test() (:44-235): produces the `model.pt` the cloak scripts load
    (stock PyTorch classifier; the harness picks Adam with a test-sized learning rate)."""
    args = make_args(batch_size=batch_size, num_epochs=epochs, model_type=model_type, att=att, pred=pred, optimizer="sgd")
    cls = mod.deep_two_d_cnn_lstm if "deep" in model_type else mod.two_d_cnn_lstm
    model = cls(input_channel=1, input_spec_size=128, cnn_filter_size=64, pred=pred, lstm_hidden_size=hidden, num_layers_lstm=2,
                attention_size=128, att=att, global_feature=0).to(device)
    # this script indexes `weights` in validate mode too (:178), so the validation speakers need an entry
    mod.weights = speaker_weights(mod.get_class_weight, {**train, **valid})
    loss = torch.nn.CrossEntropyLoss().to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-4)
    mod.scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=50, gamma=0.5)
    results = []
    for epoch in range(epochs):
        r_train = mod.train(model, device, make_loader(mod, train, batch_size, True, loader_seed + epoch), optimizer, loss, epoch, args,
                            mode="training", pred=pred)
        r_val = mod.train(model, device, make_loader(mod, valid, batch_size, True, loader_seed + 100 + epoch), optimizer, loss, epoch,
                          args, mode="validate", pred=pred)
        r_test = mod.test(model, device, make_loader(mod, test, 1, False), optimizer, loss, epoch, args, pred=pred)
        results.append({"train": r_train, "validate": r_val, "test": r_test})
    return results, model


def run_adversary_evaluation(mod, device, cloak_model, baseline_model, adversary_model, test: dict, *, mask=None, eps_seed=9, grl=1):
    """adversary_cloak_evaluation.test() (:40-110) with the globals its body reads."""
    args = make_args(grl=grl)
    mod.baseline_model, mod.adversary_model = baseline_model, adversary_model
    tape = EpsTape(eps_seed)
    cloak_model.intermed.normal.sample = tape
    with torch.no_grad():
        emo, adv = mod.test(cloak_model, device, make_loader(mod, test, 1, False), args, mask=mask)
    return emo, adv, tape
