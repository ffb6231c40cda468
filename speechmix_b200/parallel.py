"""Data parallelism for the SpeechMix hot path: one process per GPU, weights
replicated, the global batch split by sample, ONE collective per step -- a bucketed gradient
all-reduce over NCCL (NVLink 5 / NVSwitch) that is launched from backward as soon as every
gradient of a bucket has been produced, so it overlaps the remaining backward kernels.

The reference gets this implicitly from HF Trainer (nn.DataParallel / DDP, hf:trainer.py:2405-2430).
"""
import math

import torch
import torch.distributed as dist
from torch import nn


class GradientAllReducer:
    """Bucketed, backward-overlapped gradient averaging.

    Buckets are filled in reverse parameter-registration order (~ the order backward produces
    gradients: LM head / decoder first, conv stack last).  When the last gradient of a bucket
    arrives its gradients are packed into one flat buffer and ``all_reduce`` is issued
    asynchronously on the communicator's stream and ``param.grad`` is re-pointed at its slice of the
    bucket (no copy back); ``finish()`` only waits.  Works with any backend (``gloo`` in the CPU tests).

    Collectives are always issued in bucket-index order (bucket i only after bucket i-1), whatever order the
    gradients arrive in: which parameters receive a gradient may differ between ranks (LayerDrop draws per
    process), and NCCL requires every rank to issue the same sequence of collectives.

    ``payload="bf16"`` halves the bytes on the wire: gradients are packed into a bf16 bucket, summed by NCCL in
    bf16 and unpacked into the fp32 gradient slices (one rounding per rank + log2(N) in the ring; the optimizer
    state stays fp32).  ``payload="fp32"`` (default) is bit-comparable with a single-process run."""

    def __init__(self, module, world_size=None, bucket_mb=64, group=None, payload="fp32"):
        self.world = world_size if world_size is not None else dist.get_world_size()
        self.group = group
        self.bucket_mb = bucket_mb
        self.handles = []
        self._build(module, payload)

    def rebuild(self, module):
        """Re-bucket after the trainable set changed (gradual unfreezing, training.FreezingPolicy): hooks follow
        ``requires_grad`` as it is NOW.  Must be called at the same point on every rank."""
        self.remove()
        self._build(module, self.payload)

    def _build(self, module, payload):
        bucket_mb = self.bucket_mb
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG -> scale after the wait
        self.avg_in_collective = dist.get_backend(self.group) == "nccl"
        self.enabled = True
        self.graph_mode = False      # set by graph.GraphedTrainStep while it captures forward + backward
        self.events, self.order, self.comm_stream = None, [], None
        params = [p for p in module.parameters() if p.requires_grad]
        params.reverse()
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets, cur, n = [], [], 0
        for p in params:
            cur.append(p)
            n += p.numel()
            if n >= cap:
                self.buckets.append(cur)
                cur, n = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.empty(sum(p.numel() for p in b), device=b[0].device, dtype=torch.float32)
                     for b in self.buckets]
        assert payload in ("fp32", "bf16"), payload
        self.payload = payload
        self.wire = ([torch.empty_like(f, dtype=torch.bfloat16) for f in self.flat] if payload == "bf16" else self.flat)
        self.pending = [len(b) for b in self.buckets]
        self.works = [None] * len(self.buckets)
        self.ready = [False] * len(self.buckets)
        self.next_bucket = 0          # lowest bucket index not launched yet (in-order launching)
        self.owner = {}
        self.handles = []
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.owner[id(p)] = bi
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def payload_bytes(self):
        """bytes each rank hands to the collective per step"""
        return sum(w.numel() * w.element_size() for w in self.wire)

    def _views(self, bi, wire=False):
        """per-parameter slices of bucket ``bi``: of the fp32 gradient buffer, or (``wire``) of the bf16 payload"""
        buf = self.wire[bi] if wire else self.flat[bi]
        out, off = [], 0
        for p in self.buckets[bi]:
            out.append(buf[off:off + p.numel()].view_as(p))
            off += p.numel()
        return out

    def _pack(self, bi, grads):
        """gradients -> collective payload.  bf16 payload: ONE converting copy straight into the wire buffer (the fp32
        bucket is only written by ``_unpack`` after the collective); fp32 payload: the bucket is the payload."""
        torch._foreach_copy_(self._views(bi, wire=self.payload == "bf16"), grads)

    def _hook(self, p):
        if not self.enabled:    # gradient accumulation micro-step (no_sync): keep local gradients
            return
        bi = self.owner[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self.ready[bi] = True
            while self.next_bucket < len(self.buckets) and self.ready[self.next_bucket]:
                self._launch(self.next_bucket)
                self.next_bucket += 1

    # ---- CUDA-graph mode: the capture contains, per bucket, the copy of its gradients into the flat buffer and an
    # EXTERNAL event record; after every replay the communication stream waits for bucket i's event and
    # all-reduces it while the rest of the replayed backward is still running (NCCL itself stays out of the graph).
    def begin_capture(self):
        self.graph_mode, self.enabled = True, True
        self.events = [torch.cuda.Event(external=True) for _ in self.buckets]
        self.order = []
        self.pending = [len(b) for b in self.buckets]
        self.ready = [False] * len(self.buckets)
        self.next_bucket = 0
        if self.comm_stream is None and self.flat[0].is_cuda:
            self.comm_stream = torch.cuda.Stream()

    def end_capture(self):
        """still inside the capture: flush buckets whose parameters produced no gradient"""
        for bi in range(len(self.buckets)):
            if bi not in self.order:
                self._launch_captured(bi)
        self.order = sorted(self.order)     # collectives go out in index order on every rank
        self.graph_mode, self.enabled = False, False
        self.ready, self.next_bucket = [False] * len(self.buckets), 0
        self.pending = [len(b) for b in self.buckets]

    def _launch_captured(self, bi):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.buckets[bi]]
        self._pack(bi, grads)
        for p, v in zip(self.buckets[bi], self._views(bi)):
            p.grad = v          # valid after the collective + _unpack, i.e. when the optimizer reads it
        self.events[bi].record()
        self.order.append(bi)

    def reduce_after_replay(self):
        """call right after graph.replay(): per-bucket all-reduce on the communication stream, gated by the events
        the replay records; returns once the main stream is ordered after every collective."""
        main = torch.cuda.current_stream()
        op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
        works = []
        with torch.cuda.stream(self.comm_stream):
            for bi in self.order:
                self.comm_stream.wait_event(self.events[bi])
                works.append(dist.all_reduce(self.wire[bi], op=op, group=self.group, async_op=True))
            for w, bi in zip(works, self.order):     # bucket i is unpacked while the later collectives are in flight
                w.wait()
                self._unpack(bi)
        main.wait_stream(self.comm_stream)

    def _unpack(self, bi):
        """after the collective: wire -> fp32 gradient slices (bf16 payload), mean for backends without AVG"""
        if self.payload == "bf16":
            self.flat[bi].copy_(self.wire[bi])
        if not self.avg_in_collective:
            self.flat[bi].mul_(1.0 / self.world)

    def _launch(self, bi):
        if self.graph_mode:
            return self._launch_captured(bi)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.buckets[bi]]
        self._pack(bi, grads)
        for p, v in zip(self.buckets[bi], self._views(bi)):
            p.grad = v          # gradients live in the bucket from here on: no copy back after the collective
        op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
        self.works[bi] = dist.all_reduce(self.wire[bi], op=op, group=self.group, async_op=True)

    def finish(self):
        """Call after ``loss.backward()``: flush buckets whose parameters received no gradient this
        step, wait for the collectives and write the averaged gradients back."""
        for bi in range(len(self.buckets)):      # in index order: the same sequence of collectives on every rank
            if self.works[bi] is None:
                self._launch(bi)
        for bi, b in enumerate(self.buckets):
            self.works[bi].wait()
            self._unpack(bi)
            self.works[bi] = None
            self.pending[bi] = len(b)
            self.ready[bi] = False
        self.next_bucket = 0

    def reduce_inplace(self):
        """Average the gradients that already sit in ``param.grad`` WITHOUT re-pointing them (CUDA-graph mode: a
        captured forward+backward writes into static gradient tensors): bucket copy-in, one all-reduce per bucket
        (all in flight together), copy-out.  Not overlapped with backward -- the replayed graph is one launch."""
        works = []
        for bi, b in enumerate(self.buckets):
            grads = [p.grad for p in b]
            assert all(g is not None for g in grads), "reduce_inplace: every bucketed parameter needs a gradient"
            self._pack(bi, grads)
            op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
            works.append(dist.all_reduce(self.wire[bi], op=op, group=self.group, async_op=True))
        for bi, b in enumerate(self.buckets):
            works[bi].wait()
            self._unpack(bi)
            torch._foreach_copy_([p.grad for p in b], self._views(bi))

    def no_sync(self):
        """Context manager for gradient-accumulation micro-steps (ref:train.py:295 gradient_accumulation_steps):
        gradients accumulate locally; the first backward outside the context reduces the accumulated sum."""
        red = self

        class _Ctx:
            def __enter__(self):
                red.enabled = False

            def __exit__(self, *exc):
                red.enabled = True
                return False
        return _Ctx()

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


def shard_batch(global_batch, rank, world):
    """Contiguous per-rank slice of a global batch dimension (even split required)."""
    assert global_batch % world == 0, "global batch must divide evenly across ranks"
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def init_like_reference(model, seed=0):
    """Random initialisation with the statistics of the transformers initialisers
    (hf:...wav2vec2.py _init_weights, hf:...bart.py _init_weights) -- used for synthetic
    benchmarks where no checkpoint exists; parity tests load reference weights instead."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.02)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.02)
            elif isinstance(m, (nn.LayerNorm, nn.GroupNorm)):
                m.weight.fill_(1.0)
                m.bias.zero_()
            elif isinstance(m, nn.Conv1d) and not hasattr(m, "parametrizations"):
                fan_in = m.in_channels // m.groups * m.kernel_size[0]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.Conv1d):
                p = m.parametrizations.weight
                v = torch.randn(p.original1.shape, generator=g) * (2 * math.sqrt(1.0 / (m.kernel_size[0] * m.in_channels)))
                p.original1.copy_(v)
                p.original0.copy_(v.norm(dim=(0, 1), keepdim=True))
                m.bias.zero_()
    return model
