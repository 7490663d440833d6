"""Generate tests/golden/augment.npz by EXECUTING the reference's own augmentation statements -- TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_augment.py        (build container: needs /root/reference)

The block lives inline in the reference's preprocessing script (preprocess_data/preprocess_adversary_data.py, the body of
`if args.aug is not None:` at :392-421), not in a function, so it is read from the reference tree at generation time,
dedented and exec'd on a small synthetic training_dict; nothing of it is copied into this repository.  torch.normal is
wrapped to record every noise sample, so the CUDA path can be checked with the noise supplied externally."""
from __future__ import annotations

import sys
import textwrap
import types
from collections import Counter
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference/preprocess_data/preprocess_adversary_data.py")
OUT = REPO / "tests" / "golden" / "augment.npz"


def reference_block() -> str:
    lines = SRC.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == "if args.aug is not None:" and i > 300)
    end = next(i for i in range(start, len(lines)) if "['data'] = augmented_audio" in lines[i])
    return textwrap.dedent("\n".join(lines[start:end + 1]))


def run_case(seed: int, n_win: int, shape, field: str, probs):
    rng = np.random.RandomState(seed)
    classes = ["neu", "hap", "sad", "ang"] if field == "emotion" else ["F", "M"]
    labels = list(rng.choice(classes, size=n_win, p=probs))
    data = rng.standard_normal((n_win,) + shape).astype(np.float32)
    training_dict = {}
    for i in range(n_win):
        training_dict[f"utt{i // 3}_{i % 3}"] = {"data": data[i].copy(), "label": labels[i] if field == "emotion" else "neu",
                                                 "gender": labels[i] if field != "emotion" else "F"}
    keys0 = list(training_dict)
    noises = []
    real_normal = torch.normal

    def recording_normal(*a, **k):
        t = real_normal(*a, **k)
        noises.append(t.numpy().copy())
        return t

    np.random.seed(seed + 1)
    torch.manual_seed(seed + 2)
    torch.normal = recording_normal
    try:
        env = {"args": types.SimpleNamespace(aug=field), "Counter": Counter, "np": np, "torch": torch,
               "training_dict": training_dict, "train_label_list": list(labels)}
        exec(reference_block(), env)
    finally:
        torch.normal = real_normal
    keys = list(training_dict)
    # every key -> index of the ORIGINAL window whose dict it aliases
    owner = {id(training_dict[k]): i for i, k in enumerate(keys0)}
    alias_of = np.array([owner[id(training_dict[k])] for k in keys], np.int64)
    final = np.stack([training_dict[k]["data"] for k in keys0]).astype(np.float64)
    return {"labels": np.array(labels), "data": data, "alias_of": alias_of, "final": final,
            "noise": np.stack(noises).astype(np.float32) if noises else np.zeros((0,) + shape, np.float32),
            "np_seed": np.int64(seed + 1), "key_names": np.array(keys)}


def main():
    out = {}
    cases = [(11, 40, (8, 16), "emotion", [0.4, 0.3, 0.2, 0.1]), (12, 25, (4, 8), "gender", [0.7, 0.3]),
             (13, 12, (4, 8), "emotion", [0.25, 0.25, 0.25, 0.25])]
    for c, (seed, n, shape, field, probs) in enumerate(cases):
        for k, v in run_case(seed, n, shape, field, probs).items():
            out[f"c{c}_{k}"] = v
        out[f"c{c}_field"] = np.array(field)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("alias_of")})


if __name__ == "__main__":
    sys.exit(main())
