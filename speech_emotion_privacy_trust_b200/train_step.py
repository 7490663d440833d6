"""One cloak + GRL training step captured into ONE CUDA graph (no tracing compiler: plain stream capture).

The eager step of training_cloak_with_grl.train() (reference :122-169) launches ~400 small kernels; at B = 32 the host
cannot feed them fast enough.  Capturing the whole step once and replaying it removes the launch overhead.  Everything on
the path is capture safe: the cloak kernels take the current stream, allocate nothing after warm-up, and draw eps from a
DEVICE-side Philox counter, so every replay uses a fresh noise sample; the gradients live in one persistent flat buffer
(parallel.FlatGradients), so zeroing them is one memset and the data-parallel exchange is one NCCL all-reduce of that
buffer -- captured INSIDE the graph, between backward and the optimizer:

    zero flat grads | forward | backward | all-reduce(flat) | SGD          = one graph, one replay per step

With `early_params` (the layers nearest the loss: recurrent, dense and head weights, 80 % of the adversary's gradient
bytes) the exchange is split: a hook fires when the last of their gradients has been accumulated and reduces that bucket
on a side stream WHILE backward continues through the convolution stack; only the small remainder (convolution + cloak
gradients) is reduced after backward:

    ... | backward (head, dense, GRU) -+- backward (convolutions, cloak) -+- all-reduce(late) | SGD
                                       +--- all-reduce(early), side stream -+

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs, data_parallel=True)
    loss = step(x, emo, gen, w)          # copies the batch into the static buffers, replays, returns the loss tensor
"""
from __future__ import annotations

import weakref
from typing import Callable, Sequence

import torch
import torch.distributed as dist

from . import parallel


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable, example_inputs: Sequence[torch.Tensor],
                 data_parallel: bool = False, warmup: int = 3, group=None, early_params: Sequence[torch.nn.Parameter] = ()):
        self.model, self.opt, self.loss_fn, self.group = model, optimizer, loss_fn, group
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static = [t.clone() for t in example_inputs]
        self.world = dist.get_world_size(group) if (data_parallel and dist.is_initialized()) else 1
        dev = self.static[0].device
        self.opt.zero_grad(set_to_none=True)
        self.grads = parallel.FlatGradients(self.params, early=early_params if self.world > 1 else ())
        self.overlap = self.world > 1 and self.grads.n_early_params > 0
        self._side = torch.cuda.Stream(dev) if self.overlap else None
        self._expected = None         # early parameters that really receive a gradient (unused heads never fire): counted in warm-up
        self._fired = 0
        self._hooks = []
        if self.overlap:
            me = weakref.ref(self)                                 # no parameter -> hook -> step -> model cycle: the step (and the
                                                                   # NCCL kernels its graph holds) must die by reference count

            def hook(_param, me=me):
                step = me()
                if step is not None:
                    step._early_grad_ready()
            for p in self.grads.params[:self.grads.n_early_params]:
                self._hooks.append(p.register_post_accumulate_grad_hook(hook))
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
            # a process group that cannot be captured: keep the collective eager (one bucket) between two graphs
            torch.cuda.synchronize(dev)
            self.single_graph = False
            self.overlap = False
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.grads.zero()
                self.loss = self._forward_backward()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b):
                self.opt.step()
        assert self.grads.check()

    def close(self) -> None:
        """Drop the hooks and the captured graph(s).  Call it (or let the object die) BEFORE
        `dist.destroy_process_group()`: NCCL's communicator teardown waits for ever on a live graph that holds its kernels."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self.overlap = False
        for name in ("graph", "graph_b"):
            g = getattr(self, name, None)
            if g is not None:
                g.reset()
                setattr(self, name, None)

    def _early_grad_ready(self) -> None:
        """Fires once per early parameter during backward; the last one launches the early bucket's all-reduce on the side
        stream (forked from the stream backward runs on, joined again in `_step_body`)."""
        if not self.overlap:
            return
        self._fired += 1
        if self._expected is not None and self._fired == self._expected:
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.grads.allreduce_early(self.group)

    def _forward_backward(self) -> torch.Tensor:
        loss = self.loss_fn(self.model, *self.static)
        self._fired = 0
        loss.backward()
        return loss.detach()

    def _step_body(self) -> torch.Tensor:
        self.grads.zero()
        loss = self._forward_backward()
        if self.overlap:
            if self._expected is None:                             # first warm-up step: learn how many hooks fire, reduce afterwards
                self._expected = self._fired
                self.grads.allreduce_early(self.group)
            elif self._fired != self._expected:
                raise RuntimeError(f"{self._fired} early gradients arrived, {self._expected} expected: the early bucket was not reduced")
            else:
                torch.cuda.current_stream().wait_stream(self._side)
            self.grads.allreduce_late(self.group)
        elif self.world > 1:
            self.grads.allreduce(self.group)
        self.opt.step()
        return loss

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if self.graph is None:
            raise RuntimeError("GraphedTrainStep was closed")
        for dst, src in zip(self.static, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if not self.single_graph:
            self.grads.allreduce(self.group)
            self.graph_b.replay()
        return self.loss
