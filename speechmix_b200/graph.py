"""Whole-step CUDA graph: forward + loss + backward + optimizer step of one fixed-shape batch captured once and
replayed, so the ~700 kernel launches of a step cost one graph launch (no Python / driver launch gaps).

Every kernel of libspeechmix_sm100 is enqueued on the caller's stream with host-encoded TMA descriptors and raw
device pointers, so a capture records them verbatim; the captured step owns a private memory pool (activations,
gradients, the zero arena), the batch is copied into static input tensors before each replay.
"""
import torch

from . import kernels as K
from . import ops


class GraphedTrainStep:
    def __init__(self, model, optimizer, x, y, warmup=3, reducer=None, forward_kwargs=None):
        self.model, self.opt, self.reducer = model, optimizer, reducer
        self.kw = dict(forward_kwargs or {})
        self.x = x.clone()
        self.y = y.clone()
        # At least one eager step must run BEFORE the capture: optimizers create their state (moments, step
        # counters) lazily in the first step(), and a capture that contains that initialisation would reset the
        # state on every replay.
        warmup = max(int(warmup), 2)   # the second step also builds the multi-tensor cast table outside the capture
        self.warmup_losses = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up on a side stream (lazy init: weight-cache table, func attributes)
            for _ in range(warmup):
                self.warmup_losses.append(self._eager().detach().clone())
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.zero_grad(set_to_none=True)
        K._ARENA.reset()
        launches0 = K.LAUNCHES[0]
        self.graph = torch.cuda.CUDAGraph()
        if self.reducer is None:
            with torch.cuda.graph(self.graph):
                self.loss = self._eager(zero=False)
        else:
            # data parallel: NCCL collectives stay OUT of the capture.  Forward + backward are one graph that also
            # copies each finished gradient bucket into its flat buffer and records an external event; after the
            # replay the bucketed all-reduce runs on a side stream gated by those events (overlapping the rest of
            # the replayed backward) and the optimizer step runs eagerly.
            self.reducer.begin_capture()
            with torch.cuda.graph(self.graph):
                out = self.model(self.x, labels=self.y, return_model_detail=False, **self.kw)
                out["loss"].backward()
                self.reducer.end_capture()
                self.loss = out["loss"]
        self.launches_per_step = K.LAUNCHES[0] - launches0
        K._ARENA.reset()                       # the arena chunk carved during capture belongs to the graph's pool
        ops.CACHE.invalidate()

    def _eager(self, zero=True):
        if zero:
            self.opt.zero_grad(set_to_none=True)
        out = self.model(self.x, labels=self.y, return_model_detail=False, **self.kw)
        out["loss"].backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return out["loss"]

    def __call__(self, x, y):
        """Copy the batch into the static inputs (H2D or D2D, asynchronous) and replay; returns the loss tensor."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        K.LAUNCHES[0] += self.launches_per_step
        if self.reducer is not None:
            self.reducer.reduce_after_replay()
            self.opt.step()
        return self.loss
