"""Generate tests/golden/models.npz + models.json by running the REAL reference model classes -- TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_models.py        # build container: imports /root/reference/model unchanged, CPU, 1 thread

Pins rows a9 / a10 / a12 of SURVEY.md 8(a): `two_d_cnn_lstm`, `deep_two_d_cnn_lstm` (model/baseline_models.py:143-385)
and the cloak wrappers `two_d_cnn_lstm_syn`, `two_d_cnn_lstm_syn_with_grl` (model/cloak_models.py:61-226), for
att in {None, 'self_att'} x pooling in {None, 'mean'} (+ mask, + global_feature).  Combinations on which the reference
itself raises are recorded as such: the drop-ins must raise there too.

Weights are not stored: oracle/weights.fill_state rebuilds them from (seed, parameter name); inputs likewise.  Stored
per case: logits in eval mode; logits, mu/rho gradients (sub-sampled) and the first conv's weight gradient in train mode
with dropout off, for the loss sum(preds * g1) + sum(preds_grl * g2); eps injected at the reference's own draw site.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))

from oracle import weights as W  # noqa: E402

SEED = 2026
HIDDEN, FILTER, ATT = 64, 64, 128
SUB = (slice(None), slice(0, 200, 5), slice(0, 128, 4))       # (1, 40, 32) view of a (1, 200, 128) gradient


def classifier_cases():
    for cls in ("two_d_cnn_lstm", "deep_two_d_cnn_lstm"):
        for att in (None, "self_att"):
            for pred in ("emotion", "gender", "multitask"):
                yield cls, att, pred, 0
        yield cls, None, "emotion", 1
        yield cls, "self_att", "emotion", 1


def wrapper_cases():
    for wrapper in ("syn", "grl"):
        for cls in ("two_d_cnn_lstm", "deep_two_d_cnn_lstm"):
            for att in (None, "self_att"):
                for pooling in (None, "mean"):
                    yield wrapper, cls, att, pooling, False, 0
        yield wrapper, "two_d_cnn_lstm", None, "mean", True, 0          # suppression mask
        yield wrapper, "two_d_cnn_lstm", None, "mean", False, 1         # global feature (88-d) concatenated
        yield wrapper, "deep_two_d_cnn_lstm", None, None, True, 0


def case_name(*parts):
    return "|".join(str(p) for p in parts)


def build_classifier(bm, cls, att, pred, glob):
    return getattr(bm, cls)(input_channel=1, input_spec_size=128, cnn_filter_size=FILTER, pred=pred, lstm_hidden_size=HIDDEN,
                            num_layers_lstm=2, attention_size=ATT, att=att, global_feature=glob)


def build_wrapper(bm, cm, wrapper, cls, att, glob, device="cpu", max_scale=10.0, grl_lambda=0.1):
    noise = cm.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, max_scale, device)
    emo = build_classifier(bm, cls, att, "emotion", glob)
    if wrapper == "syn":
        model = cm.two_d_cnn_lstm_syn(emo, noise)
    else:
        model = cm.two_d_cnn_lstm_syn_with_grl(emo, build_classifier(bm, cls, att, "gender", glob), noise, grl_lambda)
    W.fill_state(model, SEED)
    return model.to(device)


def run_wrapper(model, wrapper, pooling, use_mask, glob, device="cpu"):
    """eval-mode logits, then train-mode (dropout off) logits + gradients.  Returns a dict of numpy arrays."""
    x, eps, mask, g, g1, g2 = (torch.from_numpy(a).to(device) for a in W.case_inputs(SEED))
    model.intermed.normal.sample = lambda shape: eps.clone()           # "eps supplied externally" (cloak_models.py:47)
    kw = {"pooling": pooling}
    if use_mask:
        kw["mask"] = mask
    if glob:
        kw["global_feature"] = g
    out = {}
    model.eval()
    with torch.no_grad():
        res = model(x, **kw)
    out["eval_preds"] = res[0].cpu().numpy()
    if wrapper == "grl":
        out["eval_preds_grl"] = res[1].cpu().numpy()
    out["noisy_sub"] = res[-1][:, 0][SUB].cpu().numpy()
    model.train()
    W.dropout_off(model)
    model.zero_grad()
    res = model(x, **kw)
    loss = (res[0] * g1).sum()
    if wrapper == "grl":
        loss = loss + (res[1] * g2).sum()
    loss.backward()
    out["train_preds"] = res[0].detach().cpu().numpy()
    if wrapper == "grl":
        out["train_preds_grl"] = res[1].detach().cpu().numpy()
        conv0 = model.gender_model.conv[1][0]
        out["gender_conv0_wgrad"] = conv0.weight.grad.cpu().numpy()
        out["gender_head_wgrad"] = model.gender_model.pred_gender_layer.weight.grad.cpu().numpy()
    out["dlocs_sub"] = model.intermed.locs.grad[SUB].cpu().numpy()
    out["drhos_sub"] = model.intermed.rhos.grad[SUB].cpu().numpy()
    out["dlocs_sum"] = np.float64(model.intermed.locs.grad.double().sum().item())
    out["drhos_abs_sum"] = np.float64(model.intermed.rhos.grad.double().abs().sum().item())
    return out


def main():
    for sub in ("model",):
        sys.path.insert(0, str(REF / sub))
    import baseline_models as bm
    import cloak_models as cm
    assert Path(bm.__file__).resolve().parent == REF / "model"
    torch.set_num_threads(1)
    arrays, meta = {}, {"seed": SEED, "hidden": HIDDEN, "classifiers": [], "wrappers": [], "raises": []}

    x, eps, mask, g, g1, g2 = (torch.from_numpy(a) for a in W.case_inputs(SEED))
    for cls, att, pred, glob in classifier_cases():
        name = case_name("clf", cls, att, pred, glob)
        m = build_classifier(bm, cls, att, pred, glob)
        W.fill_state(m, SEED)
        m.eval()
        try:
            with torch.no_grad():
                res = m(x, global_feature=g) if glob else m(x)
        except RuntimeError as e:
            meta["raises"].append(name)
            print("raises  ", name, str(e)[:60])
            continue
        res = res if isinstance(res, tuple) else (res,)
        for i, r in enumerate(res):
            arrays[f"{name}#out{i}"] = r.numpy()
        meta["classifiers"].append(name)
        print("ok      ", name, [tuple(r.shape) for r in res])

    for wrapper, cls, att, pooling, use_mask, glob in wrapper_cases():
        name = case_name(wrapper, cls, att, pooling, int(use_mask), glob)
        model = build_wrapper(bm, cm, wrapper, cls, att, glob)
        try:
            out = run_wrapper(model, wrapper, pooling, use_mask, glob)
        except RuntimeError as e:
            meta["raises"].append(name)
            print("raises  ", name, str(e)[:60])
            continue
        for k, v in out.items():
            arrays[f"{name}#{k}"] = v
        meta["wrappers"].append(name)
        print("ok      ", name)
    np.savez_compressed(OUT / "models.npz", **arrays)
    (OUT / "models.json").write_text(json.dumps(meta, indent=1))
    print(f"models.npz: {len(arrays)} arrays, {sum(a.nbytes for a in arrays.values()) / 1e6:.2f} MB raw; "
          f"{len(meta['classifiers'])} classifier cases, {len(meta['wrappers'])} wrapper cases, {len(meta['raises'])} raising")


if __name__ == "__main__":
    main()
