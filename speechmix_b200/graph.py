"""Whole-step CUDA graph: forward + loss + backward + optimizer step of one fixed-shape batch captured once and
replayed, so the ~700 kernel launches of a step cost one graph launch (no Python / driver launch gaps).

Every kernel of libspeechmix_sm100 is enqueued on the caller's stream with host-encoded TMA descriptors and raw
device pointers, so a capture records them verbatim; the captured step owns a private memory pool (activations,
gradients, the zero arena), the batch is copied into static input tensors before each replay.
"""
import torch

from . import kernels as K
from . import ops


def _capturable(opt):
    """True when ``opt.step()`` may be recorded into a CUDA graph: torch optimizers and ``optim.FusedAdafactor`` built
    with ``capturable=True`` keep their step counters on the device; anything else (host-side step counts or
    schedule values passed as kernel arguments) must run eagerly after the replay."""
    return bool(opt.param_groups) and all(bool(g.get("capturable", False)) for g in opt.param_groups)


def _check_no_host_randomness(model):
    """Host-drawn randomness would be frozen into the graph (the same layers skipped / the same frames masked on
    every replay): refuse instead of training silently wrong."""
    if hasattr(model, "update_count") and hasattr(model, "des_update"):
        raise RuntimeError("GraphedTrainStep: SpeechMixGAN switches its update phase with host-side counters "
                           "(ref:speechmix/hf_model.py:609-626: which family's .grad is cleared before the step); a "
                           "captured step would replay one phase forever -- run it eagerly")
    enc = getattr(model, "encoder_model", None)
    cfg = getattr(enc, "config", None)
    if enc is None or cfg is None or not enc.training:
        return
    if float(getattr(cfg, "layerdrop", 0.0) or 0.0) > 0.0:
        raise RuntimeError("GraphedTrainStep: config.layerdrop=%g draws its per-layer decisions on the host; a captured "
                           "step would skip the same layers on every replay -- set layerdrop=0 or run eagerly"
                           % cfg.layerdrop)


class GraphedTrainStep:
    def __init__(self, model, optimizer, x, y, warmup=3, reducer=None, forward_kwargs=None):
        self.model, self.opt, self.reducer = model, optimizer, reducer
        self.kw = dict(forward_kwargs or {})
        self.x = x.clone()
        self.y = y.clone()
        _check_no_host_randomness(model)
        # the optimizer step is captured only when it is capturable and nothing (a collective) sits between
        # backward and the step; otherwise it runs eagerly after every replay
        self.opt_in_graph = reducer is None and _capturable(optimizer)
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
                out = self.model(self.x, labels=self.y, return_model_detail=False, **self.kw)
                out["loss"].backward()
                if self.opt_in_graph:
                    self.opt.step()
                self.loss = out["loss"]
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
        # The capture holds raw pointers into the weight cache (the bf16 working copies refreshed by the captured
        # multi-tensor cast, and its pointer table): keep them alive for the lifetime of the graph, and let eager
        # calls (eval / generate between replays) build their own fresh copies.
        self._pinned = ([ent[1] for ent in ops.CACHE._store.values()] + list(ops.CACHE._tables.values()))
        ops.CACHE._store = {}
        ops.CACHE._tables = {}
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
        if self.opt_in_graph:          # the replayed optimizer step moved the masters (no Python hook runs in a replay)
            ops.CACHE.dirty = True
            ops.CACHE.epoch += 1
        K.LAUNCHES[0] += self.launches_per_step
        if self.reducer is not None:
            self.reducer.reduce_after_replay()
        if not self.opt_in_graph:
            self.opt.step()
        return self.loss
