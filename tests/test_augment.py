"""Class-balance noise augmentation (preprocess_adversary_data.py:392-421): the oracle restatement and
the host plan against vectors produced by executing the reference's own statements (oracle/make_golden_augment.py); the
CUDA kernel through the C ABI against the same vectors with the noise supplied externally."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

CASES = [0, 1, 2]


@pytest.fixture(scope="module")
def golden():
    return np.load(REPO / "tests" / "golden" / "augment.npz")


def _dict_for(g, c):
    labels, data, field = list(g[f"c{c}_labels"]), g[f"c{c}_data"], str(g[f"c{c}_field"])
    d = {}
    for i in range(len(labels)):
        d[f"utt{i // 3}_{i % 3}"] = {"data": data[i].copy(), "label": labels[i] if field == "emotion" else "neu",
                                     "gender": labels[i] if field != "emotion" else "F"}
    return d, labels, field


@pytest.mark.parametrize("c", CASES)
def test_oracle_restatement_matches_the_reference_block(golden, c):
    from oracle import augment as oracle_augment
    d, labels, field = _dict_for(golden, c)
    keys0 = list(d)
    noise = iter(golden[f"c{c}_noise"])
    np.random.seed(int(golden[f"c{c}_np_seed"]))
    oracle_augment.class_balance(d, labels, field, np.random.randint, lambda shape: next(noise))
    assert list(d) == list(golden[f"c{c}_key_names"])
    owner = {id(d[k]): i for i, k in enumerate(keys0)}
    assert np.array_equal([owner[id(d[k])] for k in d], golden[f"c{c}_alias_of"])
    final = np.stack([d[k]["data"] for k in keys0]).astype(np.float64)
    assert np.array_equal(final, golden[f"c{c}_final"])           # same operations in the same order: bit exact


@pytest.mark.parametrize("c", CASES)
def test_host_plan_reproduces_the_reference_draws(golden, c):
    from speech_emotion_privacy_trust_b200 import augmentation
    labels = list(golden[f"c{c}_labels"])
    np.random.seed(int(golden[f"c{c}_np_seed"]))
    plan = augmentation.class_balance_plan(labels)                 # NumPy's global generator, like the reference
    assert np.array_equal(plan.alias_of, golden[f"c{c}_alias_of"])
    assert plan.n_aug == len(golden[f"c{c}_noise"])
    counts = {}
    for lab in plan.labels:
        counts[lab] = counts.get(lab, 0) + 1
    assert len(set(counts.values())) == 1                          # every class as large as the largest
    # rows never drawn keep their data: final == data there
    untouched = np.setdiff1d(np.arange(len(labels)), plan.draw_rows)
    assert np.array_equal(golden[f"c{c}_final"][untouched], golden[f"c{c}_data"][untouched].astype(np.float64))


def test_balanced_input_needs_no_augmentation():
    from speech_emotion_privacy_trust_b200 import augmentation
    plan = augmentation.class_balance_plan(["a", "b", "a", "b"])
    assert plan.n_aug == 0 and list(plan.alias_of) == [0, 1, 2, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES)
def test_kernel_with_supplied_noise_matches_the_reference(golden, c):
    from speech_emotion_privacy_trust_b200 import augmentation
    labels = list(golden[f"c{c}_labels"])
    np.random.seed(int(golden[f"c{c}_np_seed"]))
    plan = augmentation.class_balance_plan(labels)
    x = torch.from_numpy(golden[f"c{c}_data"]).cuda()
    noise = torch.from_numpy(golden[f"c{c}_noise"]).cuda()
    out = augmentation.apply_plan(x, plan, noise=noise)
    ref = golden[f"c{c}_final"]                                   # float64 sums of float32 terms
    assert np.max(np.abs(out.cpu().numpy().astype(np.float64) - ref)) <= 1e-6 * max(1.0, np.max(np.abs(ref)))
    assert torch.equal(x, torch.from_numpy(golden[f"c{c}_data"]).cuda())        # out of place by default


@pytest.mark.gpu
def test_kernel_philox_noise_statistics_and_determinism():
    from speech_emotion_privacy_trust_b200 import augmentation
    rng = np.random.RandomState(5)
    labels = ["maj"] * 300 + ["min"] * 40
    plan = augmentation.class_balance_plan(labels, rng.randint)
    assert plan.n_aug == 260
    x = torch.zeros(340, 1, 200, 128, device="cuda")
    a = augmentation.apply_plan(x, plan, std=0.05, seed=77)
    b = augmentation.apply_plan(x, plan, std=0.05, seed=77)
    assert torch.equal(a, b)
    assert not torch.equal(a, augmentation.apply_plan(x, plan, std=0.05, seed=78))
    assert float(a[:300].abs().max()) == 0.0                       # majority windows untouched
    m = np.bincount(plan.draw_rows, minlength=340)                 # draws per row: the sum of m samples has variance m * std^2
    rows = np.nonzero(m)[0]
    z = (a[rows].reshape(len(rows), -1).cpu().numpy() / (0.05 * np.sqrt(m[rows])[:, None])).ravel()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.05                        # Gaussian kurtosis
    out, alias_of, key_labels = augmentation.class_balance(x, labels, seed=1, randint=np.random.RandomState(5).randint)
    assert alias_of.shape[0] == 600 and key_labels.count("min") == 300
