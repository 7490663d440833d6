"""Drop-in for model/cloak_models.py: cloak_noise, two_d_cnn_lstm_syn, two_d_cnn_lstm_syn_with_grl with the reference's
signatures, attributes and state_dict keys (intermed.locs / intermed.rhos, original_model.*, gender_model.conv.1.*).

The noise layer runs as ONE fused kernel forward (device Philox eps, no CPU RNG / H2D copy) and ONE fused kernel backward
that also folds in the gradient reversal of the gender branch (sept_cloak_fwd_f32 / sept_cloak_grl_bwd_f32)."""
import torch
import torch.nn as nn

from baseline_models import classify, self_attention_pool, sequence_features
from reversal_gradient import GradientReversal
from speech_emotion_privacy_trust_b200 import cloak_ops

EPS_STD = 0.1          # reference: torch.distributions.normal.Normal(0, 0.1)  (:37)


class cloak_noise(nn.Module):
    def __init__(self, given_locs, given_scales, min_scale, max_scale, device):
        super().__init__()
        size = given_scales.shape
        self.min_scale = min_scale
        self.max_scale = max_scale
        self.given_locs = given_locs
        self.given_scales = given_scales
        self.locs = nn.Parameter(torch.empty(size).copy_(given_locs), requires_grad=True)
        self.rhos = nn.Parameter(torch.full(size, -2.0), requires_grad=True)          # ones - 3 (:33)
        self.device = device
        self.normal = torch.distributions.normal.Normal(0, EPS_STD)
        # the reference's drivers pass CUDA scalars here (torch.tensor(0.01).to(device), training_cloak_with_grl.py:334);
        # the kernels take plain floats: convert ONCE (a float() per forward would be a device sync, illegal in capture)
        self._scale_bounds = (float(min_scale), float(max_scale))
        self._workspace = None        # backward partial sums, owned by the layer (allocated on the first forward)
        self.external_eps = None      # set to a (1, W, F) tensor to supply eps instead of drawing it on the device
        self._draws = None            # device counter of samples drawn: the Philox offset lives on the GPU, so a captured
                                      # CUDA graph draws a fresh eps per replay; identical on every data-parallel rank

    def scales(self):
        return (1.0 + torch.tanh(self.rhos)) / 2 * (self.max_scale - self.min_scale) + self.min_scale

    # ---- eps source -------------------------------------------------------------------------------------------
    def _eps_source(self):
        """(eps | None, seed, draw counter).  eps is external when the caller set `external_eps` or replaced
        `self.normal.sample` (the reference's own hook for injecting a known sample, cloak_models.py:47)."""
        if self.external_eps is not None:
            return self.external_eps, 0, None
        if 'sample' in vars(self.normal):
            return self.normal.sample(self.rhos.shape), 0, None
        if self._draws is None or self._draws.device != self.rhos.device:
            self._draws = torch.zeros(1, dtype=torch.int64, device=self.rhos.device)
        return None, torch.initial_seed(), self._draws

    def _bounds(self):
        lo, hi = self._scale_bounds
        if not torch.is_tensor(self.min_scale) and not torch.is_tensor(self.max_scale):
            lo, hi = float(self.min_scale), float(self.max_scale)       # plain numbers: follow later reassignment for free
        return lo, hi

    def _ws(self):
        if self._workspace is None or self._workspace.device != self.rhos.device:
            self._workspace = cloak_ops.new_workspace(self.rhos.device, self.rhos.numel())
        return self._workspace

    def sample_noise(self, mask=None):
        eps, seed, draw = self._eps_source()
        if eps is None:
            lo, hi = self._bounds()
            _, eps, _ = cloak_ops.cloak_forward_raw(None, self.locs.detach(), self.rhos.detach(), None, None, seed, 0,
                                                    EPS_STD, lo, hi, draw=draw)
            eps = eps.view(self.rhos.shape)
        eps = eps.to(self.rhos.device)
        if mask is not None:
            eps = eps * mask
        return self.locs + self.scales() * eps

    def forward(self, input, mask=None):
        eps, seed, draw = self._eps_source()
        lo, hi = self._bounds()
        return cloak_ops.CloakNoiseFunction.apply(input, self.locs, self.rhos, mask, eps, seed, 0, EPS_STD, lo, hi, False, 0.0,
                                                  draw, self._ws())

    def forward_with_reversed_twin(self, input, mask, grl_lambda):
        """(y, y_rev): y_rev carries the same values and reverses its gradient by -grl_lambda inside the fused backward."""
        eps, seed, draw = self._eps_source()
        lo, hi = self._bounds()
        return cloak_ops.CloakNoiseFunction.apply(input, self.locs, self.rhos, mask, eps, seed, 0, EPS_STD, lo, hi, True,
                                                  float(grl_lambda), draw, self._ws())


def _freeze(model):
    for param in model.parameters():
        param.requires_grad = False
    # the reference also tries to switch BatchNorm off here but tests parameters, not modules (:73-79): BN keeps
    # running in whatever mode the caller sets, and so it does here


def _pooled(model, att_from, x, pooling):
    if att_from.att is None:
        return x.reshape(x.shape[0], -1) if pooling is None else x.mean(dim=1)
    if att_from.att == 'self_att':
        return self_attention_pool(model, x)
    raise ValueError("Unsupported attention: {0}".format(att_from.att))


class two_d_cnn_lstm_syn(nn.Module):
    def __init__(self, original_model, noise_model):
        super().__init__()
        self.intermed = noise_model
        self.original_model = original_model
        _freeze(self.original_model)

    def forward(self, input_var, global_feature=None, mask=None, pooling=None):
        x = self.intermed(input_var.float(), mask)
        noisy = x.detach()
        m = self.original_model
        z = _pooled(m, m, sequence_features(m, x), pooling)
        return classify(m, z, global_feature), noisy


class two_d_cnn_lstm_syn_with_grl(nn.Module):
    def __init__(self, original_model, gender_model, noise_model, grl_lambda):
        super().__init__()
        self.intermed = noise_model
        self.original_model = original_model
        self.gender_model = gender_model
        _freeze(self.original_model)
        self.gender_model.conv = nn.Sequential(GradientReversal(grl_lambda), gender_model.conv)

    def forward(self, input_var, global_feature=None, mask=None, grl=False, pooling=None):
        x = input_var.float()
        stack = self.gender_model.conv
        fused = isinstance(stack, nn.Sequential) and len(stack) == 2 and isinstance(stack[0], GradientReversal) and x.is_cuda
        if fused:
            x, x_rev = self.intermed.forward_with_reversed_twin(x, mask, stack[0].lambda_)
        else:
            x = self.intermed(x, mask)
        noisy = x.detach()

        m = self.original_model
        z1 = _pooled(m, m, sequence_features(m, x), pooling)
        preds1 = classify(m, z1, global_feature, pred='emotion')

        g = self.gender_model
        if fused:
            x2 = stack[1](x_rev)
            x2 = x2.transpose(1, 2).contiguous()
            x2, _ = g.rnn(x2.reshape(x2.shape[0], x2.shape[1], -1))
        else:
            x2 = sequence_features(g, x)
        z2 = _pooled(g, m, x2, pooling)
        preds2 = classify(g, z2, global_feature, pred='gender')
        return preds1, preds2, noisy
