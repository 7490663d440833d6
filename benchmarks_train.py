"""Secondary figure of BASELINE.json: cloak + GRL training utterances/sec (config 3), measured by bench.py.

One step = what training_cloak_with_grl.train() does per batch (reference :122-169): H2D of the batch, cloak forward,
frozen emotion classifier + gender adversary through gradient reversal, speaker-weighted cross-entropies
(emotion + gender_lambda * gender, /B), backward, one flat-bucket gradient all-reduce (N > 1), SGD step, loss read back.
B = 32 per GPU (reference default :212), two_d_cnn_lstm h=64, random-init weights, synthetic z-normed windows."""
from __future__ import annotations

import numpy as np
import torch


def build_model(dev, grl_lambda=0.1, seed=8):
    from speech_emotion_privacy_trust_b200 import dropin
    dropin.install()
    import baseline_models
    import cloak_models
    torch.manual_seed(seed)
    mk = lambda pred: baseline_models.two_d_cnn_lstm(1, 128, 5, lstm_hidden_size=64, num_layers_lstm=2, pred=pred,
                                                     bidirectional=True, rnn_cell="gru", global_feature=0)
    noise = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 10.0, dev)
    return cloak_models.two_d_cnn_lstm_syn_with_grl(mk("emotion"), mk("gender"), noise, grl_lambda).to(dev)


def train_throughput(dev, rank, world, steps=20, warmup=5, batch=32, channels_last=True, graphs=True, overlap=None):
    """overlap: reduce the adversary's recurrent / dense / head gradients on a side stream while backward runs through the
    convolutions (two buckets); None = SEPT_TRAIN_OVERLAP (default off: measured 2.3 % SLOWER than one bucket on two GPUs,
    5.37 vs 5.24 ms -- the forked NCCL kernel takes SMs from the convolution backward it overlaps)."""
    import os
    import torch.distributed as dist
    if overlap is None:
        overlap = os.environ.get("SEPT_TRAIN_OVERLAP", "0") == "1"
    from speech_emotion_privacy_trust_b200 import losses, parallel, synth
    model = build_model(dev).train()
    if channels_last:
        # stock cuDNN picks its NHWC tensor-core convolutions and the fast NHWC batch-norm; the (B,1,200,128) input of
        # the cloak kernels is the same memory in either format (C = 1)
        model = model.to(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
    parallel.broadcast_parameters(model)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4)      # reference :417
    x, emo, gen, spk = synth.cloak_windows(batch * 4, seed=8 + rank)
    hx = torch.from_numpy(x).pin_memory()
    hemo, hgen = torch.from_numpy(emo).pin_memory(), torch.from_numpy(gen).pin_memory()
    w = torch.ones(batch, device=dev)
    loss_host = torch.empty(1, pin_memory=True)

    def loss_fn(m, xb, eb, gb, wb):
        p1, p2, _ = m(xb, pooling="mean")
        return losses.cloak_grl_loss(p1, p2, eb, gb, wb, 0.1)

    graphed = None
    if graphs:
        from speech_emotion_privacy_trust_b200.train_step import GraphedTrainStep
        example = [hx[:batch].to(dev), hemo[:batch].to(dev), hgen[:batch].to(dev), w]
        # the layers nearest the loss (recurrent, dense, heads: 80 % of the gradient bytes) finish first in backward
        early = [p for n, p in model.gender_model.named_parameters() if p.requires_grad and not n.startswith("conv.")] if overlap else []
        graphed = GraphedTrainStep(model, opt, loss_fn, example, data_parallel=world > 1, early_params=early)

    def step(i):
        s = (i % 4) * batch
        if graphed is not None:
            loss = graphed(hx[s:s + batch], hemo[s:s + batch], hgen[s:s + batch], w)
        else:
            xb = hx[s:s + batch].to(dev, non_blocking=True)
            eb, gb = hemo[s:s + batch].to(dev, non_blocking=True), hgen[s:s + batch].to(dev, non_blocking=True)
            loss = loss_fn(model, xb, eb, gb, w)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            parallel.allreduce_gradients(params)
            opt.step()
            loss = loss.detach()
        loss_host.copy_(loss.reshape(1), non_blocking=True)

    for i in range(warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("train_timed")
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    torch.cuda.nvtx.range_pop()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    n_params = sum(p.numel() for p in params)
    result = _train_line(graphed, batch, world, steps, ms, n_params, channels_last, graphs, float(loss_host[0]))
    if graphed is not None:
        graphed.close()               # the graph holds NCCL kernels: it must be gone before destroy_process_group()
    return result


def _train_line(graphed, batch, world, steps, ms, n_params, channels_last, graphs, final_loss):
    return {"metric": "cloak+GRL train utterances/sec", "value": batch * world * steps / (ms * 1e-3), "unit": "utterances/s",
            "ms_per_step": ms / steps, "per_gpu_batch": batch, "global_batch": batch * world, "steps": steps,
            "model": "two_d_cnn_lstm_syn_with_grl(two_d_cnn_lstm h=64 x2)", "graphs_per_step": (1 if graphed is not None and graphed.single_graph else (2 if graphed is not None else 0)),
            "allreduce": (("NCCL AVG of the flat gradient buffer in two buckets, captured inside the step graph; the early bucket "
                           f"({graphed.grads.early_numel * 4} B: recurrent / dense / head gradients) overlaps the convolution backward"
                           if graphed.overlap else "NCCL AVG of the flat gradient buffer, captured inside the step graph")
                          if (graphed is not None and graphed.single_graph) else "flat gradient buffer, eager") if world > 1 else None, "memory_format": "channels_last" if channels_last else "contiguous", "cuda_graphs": bool(graphs), "trainable_params": n_params,
            "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0, "final_loss": final_loss,
            "h2d_bytes_per_step": int(batch * 200 * 128 * 4 + batch * 16), "includes": "H2D batch, fwd, bwd, all-reduce, SGD, loss D2H"}


def eval_throughput(dev, n_utts=64, steps=5, seed=8):
    """Config 2 / 5 of BASELINE.json: cloak evaluation (adversary_cloak_evaluation.test semantics) on 64 synthetic test
    utterances of 2-10 s: sliding 200/50 windows, per-window noise, emotion classifier + gender adversary, softmax mean
    and argmax per utterance; windows/s including the window gather/normalisation and the prediction read-back."""
    from speech_emotion_privacy_trust_b200 import dropin, evaluation, normalization
    from speech_emotion_privacy_trust_b200.extraction import Layout
    dropin.install()
    import baseline_models
    import cloak_models
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    frames = rng.integers(201, 1002, size=n_utts)
    fo = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    lay = Layout(fo, torch.from_numpy(fo).to(dev), torch.zeros(len(fo), dtype=torch.int32, device=dev))
    feat = (torch.randn((int(fo[-1]), 128), device=dev) * 8 - 40).contiguous()
    st = normalization.speaker_stats(feat, lay, [u % 10 for u in range(n_utts)], whole_utterance=[True] * n_utts)
    mk = lambda pred: baseline_models.two_d_cnn_lstm(1, 128, 5, lstm_hidden_size=64, pred=pred, global_feature=0).to(dev).eval()
    base, adv = mk("emotion").to(memory_format=torch.channels_last), mk("gender").to(memory_format=torch.channels_last)
    layer = cloak_models.cloak_noise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 5.0, dev).to(dev)
    mask = evaluation.suppression_mask(layer, 0)
    n_win = len(evaluation.eval_window_table(lay)[0])
    for _ in range(2):
        evaluation.cloak_evaluate(layer, base, adv, feat, lay, st, mask=mask)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        evaluation.cloak_evaluate(layer, base, adv, feat, lay, st, mask=mask)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "cloak evaluation windows/sec", "value": n_win / (ms * 1e-3), "unit": "windows/s", "utterances_per_s": n_utts / (ms * 1e-3),
            "utterances": n_utts, "windows": n_win, "ms_per_pass": ms, "max_windows_per_forward": 512}


def cloak_kernel_bandwidth(dev, batch=64, reps=64):
    """Achieved HBM GB/s of the fused cloak forward and cloak+GRL backward kernels at B=64, W=200, F=128 (config 2).
    16 distinct input/output sets (16 x 13 MB > the 126 MB L2) are cycled so that the traffic comes from HBM."""
    from speech_emotion_privacy_trust_b200 import cloak_ops
    W, F, sets = 200, 128, 16
    wf = W * F
    xs = [torch.randn(batch, 1, W, F, device=dev) for _ in range(sets)]
    gb = [torch.randn(batch, 1, W, F, device=dev) for _ in range(sets)]
    locs, rhos = torch.zeros(wf, device=dev), torch.full((wf,), -2.0, device=dev)
    eps = 0.1 * torch.randn(wf, device=dev)
    lib = cloak_ops._lib.lib()
    ws = cloak_ops.new_workspace(dev, wf)
    outs = [torch.empty_like(xs[0]) for _ in range(sets)]
    dlocs, drhos = torch.empty(wf, device=dev), torch.empty(wf, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def fwd(i):
        cloak_ops._lib.check(lib.sept_cloak_fwd_f32(xs[i].data_ptr(), locs.data_ptr(), rhos.data_ptr(), 0, eps.data_ptr(), 0, 0, 0, 0,
                                                    0.1, 0.01, 10.0, batch, wf, outs[i].data_ptr(), 0, 0, st))

    def bwd(i):
        cloak_ops._lib.check(lib.sept_cloak_grl_bwd_f32(xs[i].data_ptr(), gb[i].data_ptr(), 0.1, eps.data_ptr(), rhos.data_ptr(), 0,
                                                        0.01, 10.0, batch, wf, ws.data_ptr(), dlocs.data_ptr(), drhos.data_ptr(), 0, st))
    res = {}
    for name, fn, nbytes in (("cloak_fwd", fwd, 2 * batch * wf * 4 + 3 * wf * 4), ("cloak_grl_bwd", bwd, 2 * batch * wf * 4 + 4 * wf * 4)):
        for i in range(sets):
            fn(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            fn(r % sets)
        e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / reps
        res[name] = {"us_per_launch": us, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / us * 1e-3}
    return res


def train_cpu_baseline(batch=32, steps=2, warmup=1):
    """The reference's training step on the host cores (oracle/train_port.py: CPU eps, float64 batch, per-sample loss
    loop, SGD), the baseline of the secondary metric; a bounded sample of `steps` batches."""
    import os
    import time
    from oracle import train_port
    from speech_emotion_privacy_trust_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = train_port.build("cpu").train()
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-3, momentum=0.9, weight_decay=1e-4)
    x, emo, gen, _ = synth.cloak_windows(batch, seed=8)
    x64, emo, gen, w = torch.from_numpy(x).double(), torch.from_numpy(emo), torch.from_numpy(gen), torch.ones(batch)
    for _ in range(warmup):
        train_port.train_step(model, opt, x64, emo, gen, w, "cpu")
    t0 = time.perf_counter()
    for _ in range(steps):
        train_port.train_step(model, opt, x64, emo, gen, w, "cpu")
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": "utterances/s", "cores": threads, "kind": "port", "ms_per_step": dt * 1e3,
            "sample": f"{steps} steps of B={batch} after {warmup} warm-up, reference-style step (oracle/train_port.py)"}


def eval_cpu_baseline(n_utts=4, seed=8):
    """The reference's evaluation loop on the host cores (oracle/train_port.evaluate_utterance: batch 1 per window),
    windows/s on a bounded sample of synthetic test utterances."""
    import os
    import time
    from oracle import train_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    noise = train_port.CloakNoise(torch.zeros(1, 200, 128), torch.ones(1, 200, 128), 0.01, 5.0, "cpu")
    base, adv = train_port.Classifier("emotion").eval(), train_port.Classifier("gender").eval()
    feats = [torch.randn(int(rng.integers(201, 1002)), 128) for _ in range(n_utts)]
    train_port.evaluate_utterance(noise, base, adv, feats[0][:250])
    t0 = time.perf_counter()
    n_win = sum(train_port.evaluate_utterance(noise, base, adv, f)[2] for f in feats)
    dt = time.perf_counter() - t0
    return {"value": n_win / dt, "unit": "windows/s", "cores": threads, "kind": "port",
            "sample": f"{n_utts} utterances, {n_win} windows, batch 1 per window (oracle/train_port.py)"}
