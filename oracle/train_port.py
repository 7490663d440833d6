"""Plain-PyTorch restatement of the cloak + GRL training step -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference module by module, in its own operation order, with stock torch ops only (no product code):
  cloak_noise                    model/cloak_models.py:24-58      eps ~ N(0, 0.1) drawn on the CPU, then .to(device)
  GradientReversal               model/reversal_gradient.py:5-32  clone forward, -lambda * g backward
  two_d_cnn_lstm                 model/baseline_models.py:143-260 (att None, global_feature 0)
  two_d_cnn_lstm_syn_with_grl    model/cloak_models.py:133-226
  train step                     training/training_cloak_with_grl.py:122-169: float64 batch -> device, per-sample weighted
                                 cross-entropy loop (2B tiny launches), -scale_lamda * log(mean sigma), zero_grad/backward/step
It is the CPU baseline of bench.py's secondary metric and the independent checker of the fused GPU path
(tests/test_train_parity_gpu.py).  Parameter names equal the reference's, so state_dicts are interchangeable.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Reverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = lambda_
        return x.clone()

    @staticmethod
    def backward(ctx, grads):
        return -grads.new_tensor(ctx.lambda_) * grads, None


class GradientReversal(nn.Module):
    def __init__(self, lambda_=1):
        super().__init__()
        self.lambda_ = lambda_

    def forward(self, x):
        return _Reverse.apply(x, self.lambda_)


class CloakNoise(nn.Module):
    def __init__(self, given_locs, given_scales, min_scale, max_scale, device):
        super().__init__()
        self.min_scale, self.max_scale, self.device = min_scale, max_scale, device
        self.locs = nn.Parameter(given_locs.clone().float())
        self.rhos = nn.Parameter(torch.ones(given_scales.shape) - 3)
        self.normal = torch.distributions.normal.Normal(0, 0.1)

    def scales(self):
        return (1.0 + torch.tanh(self.rhos)) / 2 * (self.max_scale - self.min_scale) + self.min_scale

    def sample_noise(self, mask=None):
        eps = self.normal.sample(self.rhos.shape).to(self.rhos.device)
        if mask is not None:
            eps = eps * mask
        return self.locs + self.scales() * eps

    def forward(self, x, mask=None):
        noise = self.sample_noise(mask)
        return x + noise if mask is None else x * mask + noise


class Classifier(nn.Module):
    """two_d_cnn_lstm with att None, global_feature 0 (the configuration the cloak scripts load)."""

    def __init__(self, pred="emotion", hidden=64, p_drop=0.2):
        super().__init__()
        self.pred = pred
        layers = []
        for c_in, c_out in ((1, 32), (32, 64), (64, 128)):
            layers += [nn.Conv2d(c_in, c_out, kernel_size=5, padding=2), nn.BatchNorm2d(c_out), nn.ReLU(),
                       nn.MaxPool2d(kernel_size=(2, 2), stride=(2, 2)), nn.Dropout2d(p_drop)]
        self.conv = nn.Sequential(*layers)
        self.rnn = nn.GRU(input_size=128 * 128 // 8, hidden_size=hidden, num_layers=2, batch_first=True, dropout=p_drop,
                          bidirectional=True)
        self.dropout = nn.Dropout(p_drop)
        self.dense1 = nn.Linear(2 * hidden, 128)
        self.dense_relu1 = nn.ReLU()
        self.pred_emotion_layer = nn.Linear(128, 4)
        self.pred_gender_layer = nn.Linear(128, 2)

    def features(self, x, conv=None):
        x = (self.conv if conv is None else conv)(x.float())
        x = x.transpose(1, 2).contiguous()
        size = x.size()
        x, _ = self.rnn(x.reshape(-1, size[1], size[2] * size[3]))
        return torch.mean(x, dim=1)                                   # pooling='mean' (training_cloak_with_grl.py:137)

    def head(self, z, pred):
        z = self.dropout(self.dense_relu1(self.dense1(z)))
        return self.pred_emotion_layer(z) if pred == "emotion" else self.pred_gender_layer(z)


class CloakGRLModel(nn.Module):
    def __init__(self, original_model: Classifier, gender_model: Classifier, noise_model: CloakNoise, grl_lambda: float):
        super().__init__()
        self.intermed, self.original_model, self.gender_model = noise_model, original_model, gender_model
        for p in self.original_model.parameters():
            p.requires_grad = False
        self.gender_model.conv = nn.Sequential(GradientReversal(grl_lambda), gender_model.conv)

    def forward(self, input_var, mask=None):
        x = self.intermed(input_var.float(), mask)
        noisy = x.detach()
        p1 = self.original_model.head(self.original_model.features(x), "emotion")
        p2 = self.gender_model.head(self.gender_model.features(x), "gender")
        return p1, p2, noisy


def reference_loss(model, preds, preds_grl, labels_emo, labels_gen, weights, gender_lambda, scale_lamda):
    """The per-sample loop of training_cloak_with_grl.py:141-160."""
    ce = nn.CrossEntropyLoss()
    total = 0
    n = len(preds)
    for i in range(n):
        total = total + (ce(preds[i].unsqueeze(dim=0), labels_emo[i:i + 1]) * weights[i]) / n
        total = total + (float(gender_lambda) * ce(preds_grl[i].unsqueeze(dim=0), labels_gen[i:i + 1]) * weights[i]) / n
    return total - float(scale_lamda) * torch.log(torch.mean(model.intermed.scales()))


def train_step(model, optimizer, features64, labels_emo, labels_gen, weights, device, gender_lambda=0.1, scale_lamda=0.0):
    """One iteration of the reference loop (:122-169): float64 features to the device, forward, loss loop, step."""
    features = features64.to(device)
    labels_emo, labels_gen = labels_emo.to(device), labels_gen.to(device)
    p1, p2, _ = model(features)
    loss = reference_loss(model, p1, p2, labels_emo, labels_gen, weights, gender_lambda, scale_lamda)
    value = loss.item()
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return value


def build(device="cpu", grl_lambda=0.1, seed=8, min_scale=0.01, max_scale=10.0):
    torch.manual_seed(seed)
    noise = CloakNoise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), min_scale, max_scale, device)
    return CloakGRLModel(Classifier("emotion"), Classifier("gender"), noise, grl_lambda).to(device)


@torch.no_grad()
def evaluate_utterance(noise_model: CloakNoise, baseline_model: Classifier, adversary_model: Classifier, feat, device="cpu",
                       mask=None, win_len=200, shift_len=50):
    """The per-utterance body of adversary_cloak_evaluation.test() (:69-93): batch 1, one window at a time, two softmax
    rows copied to the host per window, mean + argmax.  feat: (T, 128) normalised features, T >= win_len."""
    import numpy as np
    sm = nn.Softmax(dim=1)
    pe, pg = [], []
    for i in range(int((feat.shape[0] - win_len) / shift_len) + 1):
        x = feat[None, None, i * shift_len:i * shift_len + win_len, :].to(device)
        noisy = noise_model(x.float(), mask).detach()
        pe.append(sm(baseline_model.head(baseline_model.features(noisy), "emotion")).cpu().numpy()[0])
        pg.append(sm(adversary_model.head(adversary_model.features(noisy), "gender")).cpu().numpy()[0])
    return int(np.argmax(np.mean(np.array(pe), axis=0))), int(np.argmax(np.mean(np.array(pg), axis=0))), len(pe)
