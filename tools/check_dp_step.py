#!/usr/bin/env python3
"""Data-parallel correctness of the graphed training step on real GPUs (run under torchrun, >= 2 ranks):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/check_dp_step.py

Three copies of the cloak + GRL model start from the same weights on every rank and take the same 6 steps on rank-specific
batches with the same device-drawn eps: (a) eager step + parallel.allreduce_gradients (the plain reference of the exchange),
(b) GraphedTrainStep with ONE bucket captured in the graph, (c) GraphedTrainStep with the early bucket overlapped on a side
stream.  All three must end with the same parameters (fp32 rounding of the reduction order aside) on every rank, and every
rank must hold identical parameters."""
import faulthandler
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import benchmarks_train
from speech_emotion_privacy_trust_b200 import losses, parallel, synth
from speech_emotion_privacy_trust_b200.train_step import GraphedTrainStep

faulthandler.dump_traceback_later(int(os.environ.get("SEPT_CHECK_TIMEOUT", "150")), exit=True)   # a stuck collective must not hold the box
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic = True
B, STEPS = 16, 6
x, emo, gen, _ = synth.cloak_windows(B * STEPS, seed=100 + rank)
xs, es, gs = torch.from_numpy(x).to(dev), torch.from_numpy(emo).to(dev), torch.from_numpy(gen).to(dev)
w = torch.ones(B, device=dev)


def loss_fn(m, xb, eb, gb, wb):
    p1, p2, _ = m(xb, pooling="mean")
    return losses.cloak_grl_loss(p1, p2, eb, gb, wb, 0.1)


def fresh():
    torch.manual_seed(8)                                  # same weights AND same Philox seed for eps on every rank / copy
    m = benchmarks_train.build_model(dev).train()
    for mod in m.modules():                               # dropout draws from the CUDA generator: keep the copies comparable
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
        if isinstance(mod, torch.nn.RNNBase):
            mod.dropout = 0.0
    parallel.broadcast_parameters(m)
    params = [p for p in m.parameters() if p.requires_grad]
    return m, params, torch.optim.SGD(params, lr=1e-2, momentum=0.9, weight_decay=1e-4)


results = {}
# (a) eager
m, params, opt = fresh()
for i in range(STEPS + 3):                                # 3 extra = the warm-up steps the graphed variants take on the example batch
    j = 0 if i < 3 else i - 3
    sl = slice(j * B, (j + 1) * B)
    opt.zero_grad(set_to_none=True)
    loss_fn(m, xs[sl], es[sl], gs[sl], w).backward()
    parallel.allreduce_gradients(params)
    opt.step()
results["eager"] = [p.detach().clone() for p in params]
# (b), (c) graphed
for tag, use_early in (("graph_one_bucket", False), ("graph_overlap", True)):
    m, params, opt = fresh()
    early = [p for n, p in m.gender_model.named_parameters() if p.requires_grad and not n.startswith("conv.")] if use_early else []
    step = GraphedTrainStep(m, opt, loss_fn, [xs[:B], es[:B], gs[:B], w], data_parallel=True, early_params=early, warmup=3)
    assert step.single_graph and step.overlap == use_early, (step.single_graph, step.overlap)
    for j in range(STEPS):
        sl = slice(j * B, (j + 1) * B)
        step(xs[sl], es[sl], gs[sl], w)
    torch.cuda.synchronize()
    results[tag] = [p.detach().clone() for p in params]
    step.close()                                          # a live graph holding NCCL kernels blocks destroy_process_group()

worst = {}
for tag in ("graph_one_bucket", "graph_overlap"):
    worst[tag] = max(float((a - b).abs().max() / a.abs().max().clamp_min(1e-12)) for a, b in zip(results["eager"], results[tag]))
flat = torch.cat([p.reshape(-1) for p in results["graph_overlap"]])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
across = max(float((g - flat).abs().max()) for g in gathered)
moved = float((flat - torch.cat([p.reshape(-1) for p in fresh()[1]]).detach()).abs().max())
if rank == 0:
    print(f"ranks {world}: graphed vs eager data-parallel step after {STEPS} steps, worst relative parameter difference: {worst}; "
          f"largest difference between ranks {across:.3e}; parameters moved by up to {moved:.3e}")
assert across == 0.0, "ranks diverged"
assert all(v < 5e-3 for v in worst.values()), worst      # the warm-up of the graphed variants replays the example batch: same steps as (a)
dist.barrier()
dist.destroy_process_group()
