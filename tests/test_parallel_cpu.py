"""Host-side multi-GPU logic on CPU: world_size-2 gloo groups exercise the utterance sharding, the flat-bucket
gradient all-reduce and the per-speaker partial merge that the B200 path runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_emotion_privacy_trust_b200 import parallel


def test_shard_by_length_balances_and_partitions():
    rng = np.random.default_rng(0)
    lens = rng.integers(32000, 160001, size=5531)
    for world in (1, 2, 4, 8):
        parts = parallel.shard_by_length(lens, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(lens)))              # a partition: every utterance exactly once
        loads = np.array([lens[p].sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.001                         # balanced to 0.1 % of the audio
        if world > 1:                                                     # every rank sees the same mix of lengths
            means = [lens[p].mean() for p in parts]
            assert max(means) / min(means) < 1.02
    assert [len(p) for p in parallel.shard_by_length([5, 4, 3], 4)] == [1, 1, 1, 0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        frozen = torch.nn.Linear(3, 3)
        for p in frozen.parameters():
            p.requires_grad = False
        if rank == 1:                                                     # ranks start different; broadcast fixes it
            for p in model.parameters():
                p.data.add_(1.0)
        parallel.broadcast_parameters(model)
        g = torch.Generator().manual_seed(100)
        x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        xs, ys = x_all[rank * 4:(rank + 1) * 4], y_all[rank * 4:(rank + 1) * 4]
        loss = ((frozen(model(xs)) - ys) ** 2).mean()
        loss.backward()
        model[2].bias.grad = None                                         # a parameter without a gradient on this rank
        n = parallel.allreduce_gradients(list(model.parameters()) + list(frozen.parameters()))
        grads = [p.grad.clone() for p in model.parameters()]
        # reference: the same model on the whole batch in one process
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ((frozen(ref(x_all)) - y_all) ** 2).mean().backward()
        ok = n == sum(p.numel() for p in model.parameters())
        for i, (gp, rp) in enumerate(zip(grads, ref.parameters())):
            want = torch.zeros_like(rp.grad) if i == 3 else rp.grad      # the dropped bias gradient reduces to zero
            ok = ok and torch.allclose(gp, want, atol=1e-6)
        # per-speaker partial merge: each rank holds half the frames of 3 speakers
        rng = np.random.default_rng(7)
        data = rng.standard_normal((3, 40, 4)) * 5 + 2
        mine = torch.from_numpy(data[:, rank * 20:(rank + 1) * 20])
        cnt = torch.full((3, 4), 20.0, dtype=torch.float64)
        mean = mine.mean(1)
        m2 = ((mine - mean[:, None]) ** 2).sum(1)
        n_, mu, sd, lo, hi = parallel.merge_speaker_partials(cnt, mean, m2, mine.amin(1), mine.amax(1))
        ok = ok and torch.allclose(mu, torch.from_numpy(data.mean(1))) and torch.allclose(sd, torch.from_numpy(data.std(1)))
        ok = ok and torch.equal(lo, torch.from_numpy(data.min(1))) and torch.equal(hi, torch.from_numpy(data.max(1))) and float(n_[0, 0]) == 40.0
        # FlatGradients: .grad are views of one buffer; two data-parallel SGD steps == two full-batch steps in one process
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 3))
        net = net.to(memory_format=torch.channels_last)
        torch.manual_seed(1)
        full = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 3))
        xi, yi = torch.randn(8, 1, 6, 6, generator=g), torch.randn(8, 3, generator=g)
        fg = parallel.FlatGradients(net.parameters())
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
        opt_full = torch.optim.SGD(full.parameters(), lr=0.1, momentum=0.9)
        ok = ok and fg.check() and all(p.grad.stride() == p.stride() for p in net.parameters())
        for _ in range(2):
            fg.zero()
            ((net(xi[rank * 4:(rank + 1) * 4]) - yi[rank * 4:(rank + 1) * 4]) ** 2).mean().backward()
            ok = ok and fg.check()                                        # autograd accumulated in place: still views
            ok = ok and fg.allreduce() == sum(p.numel() for p in net.parameters())
            opt.step()
            opt_full.zero_grad()
            ((full(xi) - yi) ** 2).mean().backward()
            opt_full.step()
        for a, b in zip(net.parameters(), full.parameters()):
            ok = ok and torch.allclose(a, b, atol=1e-6)
        # two buckets: the layer nearest the loss sits at the front of the buffer and is reduced on its own
        # (train_step.GraphedTrainStep overlaps that exchange with the rest of backward); early + late == one bucket
        torch.manual_seed(1)
        net2 = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 3))
        for p in net2.parameters():
            p.grad = None
        fg2 = parallel.FlatGradients(net2.parameters(), early=list(net2[3].parameters()))
        ok = ok and fg2.check() and fg2.n_early_params == 2 and fg2.early_numel == 4 * 6 * 6 * 3 + 3
        ok = ok and fg2.params[0] is net2[3].weight and fg2.params[2] is net2[0].weight
        fg2.zero()
        ((net2(xi[rank * 4:(rank + 1) * 4]) - yi[rank * 4:(rank + 1) * 4]) ** 2).mean().backward()
        ok = ok and fg2.allreduce_early() == fg2.early_numel and fg2.allreduce_late() == fg2.flat.numel() - fg2.early_numel
        torch.manual_seed(1)
        full2 = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 3))
        ((full2(xi) - yi) ** 2).mean().backward()
        for a, b in zip(net2.parameters(), full2.parameters()):
            ok = ok and torch.allclose(a.grad, b.grad, atol=1e-6)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_partial_merge_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(out) == {0: True, 1: True}
