"""The fused GPU training path (drop-in modules: one cloak forward kernel, one cloak + gradient-reversal backward kernel,
vectorised loss) against the plain-PyTorch restatement of the reference's step (oracle/train_port.py), same weights,
same batch, same eps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_step_matches_reference_style_step():
    from oracle import train_port
    from speech_emotion_privacy_trust_b200 import dropin, losses, synth
    dropin.install()
    import baseline_models
    import cloak_models
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        ref = train_port.build(dev)
        with torch.no_grad():
            ref.intermed.locs.add_(0.05 * torch.randn_like(ref.intermed.locs))
            ref.intermed.rhos.add_(0.5 * torch.randn_like(ref.intermed.rhos))
        mk = lambda pred: baseline_models.two_d_cnn_lstm(1, 128, 5, lstm_hidden_size=64, pred=pred, global_feature=0)
        noise = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, dev)
        new = cloak_models.two_d_cnn_lstm_syn_with_grl(mk("emotion"), mk("gender"), noise, 0.1).to(dev)
        missing = new.load_state_dict(ref.state_dict(), strict=False)          # the port has no attention/dense2 layers
        assert not missing.unexpected_keys and all("att_" in k or "dense2" in k or "pred_" in k for k in missing.missing_keys)
        for m in (ref, new):
            m.train()
            for mod in m.modules():                                           # same dropout-free, running-stat-free forward
                if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
                    mod.p = 0.0
                if isinstance(mod, torch.nn.GRU):
                    mod.dropout = 0.0
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.eval()
        B = 6
        x, emo, gen, spk = synth.cloak_windows(B, seed=3)
        w = torch.tensor(np.linspace(0.5, 2.0, B), dtype=torch.float32, device=dev)
        eps = 0.1 * torch.randn(1, 200, 128)
        ref.intermed.normal.sample = lambda shape: eps.clone()                # the reference's own hook (CPU sample)
        new.intermed.external_eps = eps.to(dev)
        x64 = torch.from_numpy(x).double()                                    # the reference feeds float64 windows
        emo_t, gen_t = torch.from_numpy(emo).to(dev), torch.from_numpy(gen).to(dev)

        p1r, p2r, noisy_r = ref(x64.to(dev))
        loss_r = train_port.reference_loss(ref, p1r, p2r, emo_t, gen_t, w, 0.1, 0.05)
        loss_r.backward()

        p1n, p2n, noisy_n = new(x64.to(dev), pooling="mean")
        loss_n = losses.cloak_grl_loss(p1n, p2n, emo_t, gen_t, w, 0.1, sigma=new.intermed.scales(), scale_lamda=0.05)
        loss_n.backward()

        assert float((noisy_n - noisy_r).abs().max()) < 1e-6                  # cloak forward, north_star tolerance
        assert float((p1n - p1r).abs().max()) < 1e-4 and float((p2n - p2r).abs().max()) < 1e-4
        assert abs(float(loss_n) - float(loss_r)) < 1e-5
        pairs = [(new.intermed.locs.grad, ref.intermed.locs.grad, "dlocs"), (new.intermed.rhos.grad, ref.intermed.rhos.grad, "drhos"),
                 (new.gender_model.conv[1][0].weight.grad, ref.gender_model.conv[1][0].weight.grad, "gender conv1 wgrad"),
                 (new.gender_model.pred_gender_layer.weight.grad, ref.gender_model.pred_gender_layer.weight.grad, "gender head wgrad")]
        for got, want, name in pairs:
            scale = max(float(want.abs().max()), 1e-6)
            assert float((got - want).abs().max()) < 2e-4 * scale + 1e-7, name
        assert all(p.grad is None for p in new.original_model.parameters())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_fused_losses_equal_the_reference_per_sample_loop():
    """losses.cloak_grl_loss / weighted_cross_entropy against the literal loops of training_cloak_with_grl.py:141-160 and
    training_cloak.py:139-147 (value and gradients), training and validate branches, labels shaped (B, 1) as the
    reference's collate delivers them."""
    from speech_emotion_privacy_trust_b200 import losses
    dev = torch.device("cuda")
    torch.manual_seed(3)
    B = 32
    weights = {f"{s}_{d}": 1.0 + 0.37 * s for s in range(5) for d in ("iemocap", "crema-d")}
    spk = [int(i % 5) for i in range(B)]
    dsets = ["iemocap" if i % 3 else "crema-d" for i in range(B)]
    emo = torch.randint(0, 4, (B, 1), device=dev)
    gen = torch.randint(0, 2, (B, 1), device=dev)
    rhos = (torch.randn(1, 200, 128, device=dev) - 2).requires_grad_()
    ce = torch.nn.CrossEntropyLoss().to(dev)
    for training in (True, False):
        for scale_lamda in (0.0, 0.05):
            outs = []
            for fused in (False, True):
                p1 = torch.randn(B, 4, device=dev, generator=torch.Generator(dev).manual_seed(1)).requires_grad_()
                p2 = torch.randn(B, 2, device=dev, generator=torch.Generator(dev).manual_seed(2)).requires_grad_()
                rhos.grad = None
                sigma = (1.0 + torch.tanh(rhos)) / 2 * (10.0 - 0.01) + 0.01
                if fused:
                    w = losses.speaker_weight_vector(weights, spk, dsets, dev) if training else None
                    total = losses.cloak_grl_loss(p1, p2, emo, gen, w, 0.1, sigma=sigma, scale_lamda=scale_lamda)
                else:
                    total = 0
                    for i in range(B):                                            # reference :143-154
                        sid = str(spk[i]) + "_" + dsets[i]
                        if training:
                            total += (ce(p1[i].unsqueeze(dim=0), emo[i]) * weights[sid]) / B
                            total += (float(0.1) * ce(p2[i].unsqueeze(dim=0), gen[i]) * weights[sid]) / B
                        else:
                            total += (ce(p1[i].unsqueeze(dim=0), emo[i])) / B
                            total += (float(0.1) * ce(p2[i].unsqueeze(dim=0), gen[i])) / B
                    total = total - float(scale_lamda) * torch.log(torch.mean(sigma))     # :158-160
                total.backward()
                outs.append((total.detach(), p1.grad.clone(), p2.grad.clone(), None if rhos.grad is None else rhos.grad.clone()))
            (la, g1a, g2a, gra), (lb, g1b, g2b, grb) = outs
            assert abs(float(la) - float(lb)) < 2e-6 * max(1.0, abs(float(la)))
            assert float((g1a - g1b).abs().max()) < 1e-7 and float((g2a - g2b).abs().max()) < 1e-7
            if scale_lamda:
                assert float((gra - grb).abs().max()) < 1e-9
    # single head (training_cloak.py:139-143)
    p = torch.randn(B, 4, device=dev)
    w = losses.speaker_weight_vector(weights, spk, dsets, dev)
    want = sum(ce(p[i].unsqueeze(0), emo[i]) * weights[str(spk[i]) + "_" + dsets[i]] / B for i in range(B))
    assert abs(float(losses.weighted_cross_entropy(p, emo, w)) - float(want)) < 2e-6
