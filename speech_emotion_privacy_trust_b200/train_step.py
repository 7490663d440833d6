"""One cloak + GRL training step captured into CUDA graphs (no tracing compiler: plain stream capture).

The eager step of training_cloak_with_grl.train() (reference :122-169) launches ~400 small kernels; at B = 32 the host
cannot feed them fast enough.  Capturing forward + backward (+ optimizer) once and replaying removes the launch
overhead.  Everything on the path is capture safe: the cloak kernels take the current stream, allocate nothing after
warm-up, and draw eps from a DEVICE-side Philox counter, so every replay uses a fresh noise sample.

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs, allreduce=parallel.allreduce_gradients)
    loss = step(x, emo, gen, w)          # copies the batch into the static buffers, replays, returns the loss tensor
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable, example_inputs: Sequence[torch.Tensor],
                 allreduce: Callable | None = None, warmup: int = 3):
        self.model, self.opt, self.loss_fn, self.allreduce = model, optimizer, loss_fn, allreduce
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static = [t.clone() for t in example_inputs]
        dev = self.static[0].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                              # warm-up outside capture: cuDNN autotune, workspaces
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._forward_backward().item()
                if self.allreduce:
                    self.allreduce(self.params)
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.opt.zero_grad(set_to_none=True)
        self.fused = allreduce is None
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.loss = self._forward_backward()
            if self.fused:
                self.opt.step()
        if not self.fused:
            # the gradient all-reduce (NCCL) stays between two graphs: backward | all-reduce | optimizer
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b):
                self.opt.step()

    def _forward_backward(self) -> torch.Tensor:
        loss = self.loss_fn(self.model, *self.static)
        loss.backward()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph_a.replay()
        if not self.fused:
            self.allreduce(self.params)
            self.graph_b.replay()
        return self.loss
