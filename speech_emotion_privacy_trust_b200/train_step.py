"""One cloak + GRL training step captured into ONE CUDA graph (no tracing compiler: plain stream capture).

The eager step of training_cloak_with_grl.train() (reference :122-169) launches ~400 small kernels; at B = 32 the host
cannot feed them fast enough.  Capturing the whole step once and replaying it removes the launch overhead.  Everything on
the path is capture safe: the cloak kernels take the current stream, allocate nothing after warm-up, and draw eps from a
DEVICE-side Philox counter, so every replay uses a fresh noise sample; the gradients live in one persistent flat buffer
(parallel.FlatGradients), so zeroing them is one memset and the data-parallel exchange is one NCCL all-reduce of that
buffer -- captured INSIDE the graph, between backward and the optimizer:

    zero flat grads | forward | backward | all-reduce(flat) | SGD          = one graph, one replay per step

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs, data_parallel=True)
    loss = step(x, emo, gen, w)          # copies the batch into the static buffers, replays, returns the loss tensor
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist

from . import parallel


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable, example_inputs: Sequence[torch.Tensor],
                 data_parallel: bool = False, warmup: int = 3, group=None):
        self.model, self.opt, self.loss_fn, self.group = model, optimizer, loss_fn, group
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static = [t.clone() for t in example_inputs]
        self.world = dist.get_world_size(group) if (data_parallel and dist.is_initialized()) else 1
        dev = self.static[0].device
        self.opt.zero_grad(set_to_none=True)
        self.grads = parallel.FlatGradients(self.params)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                              # warm-up outside capture: cuDNN autotune, workspaces,
            for _ in range(warmup):                                # NCCL communicator + its buffers
                self._step_body().item()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        assert self.grads.check(), "a gradient was re-allocated during warm-up: .grad no longer aliases the flat buffer"
        self.single_graph = True
        self.graph = torch.cuda.CUDAGraph()
        try:
            # thread_local: the NCCL watchdog thread may query events while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                self.loss = self._step_body()
        except RuntimeError:
            if self.world == 1:
                raise
            # a process group that cannot be captured: keep the collective eager between two graphs
            torch.cuda.synchronize(dev)
            self.single_graph = False
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.grads.zero()
                self.loss = self._forward_backward()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b):
                self.opt.step()
        assert self.grads.check()

    def _forward_backward(self) -> torch.Tensor:
        loss = self.loss_fn(self.model, *self.static)
        loss.backward()
        return loss.detach()

    def _step_body(self) -> torch.Tensor:
        self.grads.zero()
        loss = self._forward_backward()
        if self.world > 1:
            self.grads.allreduce(self.group)
        self.opt.step()
        return loss

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if not self.single_graph:
            self.grads.allreduce(self.group)
            self.graph_b.replay()
        return self.loss
