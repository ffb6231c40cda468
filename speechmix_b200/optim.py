"""Fused Adafactor: the optimizer of the reference recipe (ref:train.py:298 ``optim="adafactor"``; the HF Trainer builds
``transformers.optimization.Adafactor(lr=..., scale_parameter=False, relative_step=False)``), as SIX kernel launches
per step over all parameters (four over a tile table, two block-per-slice launches for the small factored slices) (``smx_adafactor_step``, csrc/adafactor.cu) instead of ~15 small launches per parameter.

State layout and names follow the transformers implementation (``step``, ``exp_avg_sq_row``, ``exp_avg_sq_col``,
``exp_avg_sq``, ``RMS``), so optimizer checkpoints are interchangeable.  Only the configuration the reference uses is
implemented: explicit ``lr``, no relative step, no parameter scaling, no first moment.  CUDA fp32 parameters only --
anything else raises (there is no CPU path)."""
import ctypes
import math

import numpy as np
import torch

from . import _lib

TILE_R, TILE_C = 64, 256
SMALL_ELEMS, SMALL_RC = 16384, 4096     # csrc/adafactor.cu: block-per-slice path for small factored slices


def is_small_slice(rows, cols):
    return rows * cols <= SMALL_ELEMS and rows + cols <= SMALL_RC


def factored_dims(shape):
    """(factored, batch, rows, cols) the way transformers' Adafactor splits a parameter: len(shape) >= 2 -> factored
    over the last two dims, leading dims are independent slices (hf:optimization.py ``_get_options`` / ``step``)."""
    if len(shape) >= 2:
        batch = 1
        for d in shape[:-2]:
            batch *= int(d)
        return True, batch, int(shape[-2]), int(shape[-1])
    n = 1
    for d in shape:
        n *= int(d)
    return False, 1, 1, n


def tile_table(shapes):
    """int32 [n_tiles, 4] (tensor, b, r0, c0) covering every element of every LARGE tensor exactly once with 64 x 256
    tiles (vectors are viewed as [ceil(n / 256), 256]); int32 [n_slices, 2] (tensor, b) for the large factored tensors;
    int32 [n_small, 2] (tensor, b) for the factored slices that take the block-per-slice kernels instead of tiles
    (``is_small_slice``: a [512][3] convolution-weight slice would otherwise be eight near-empty tiles); and the
    shared-memory floats the largest of those needs."""
    tiles, slices, small, small_floats = [], [], [], 0
    for i, shape in enumerate(shapes):
        factored, batch, rows, cols = factored_dims(shape)
        b = np.arange(batch, dtype=np.int32)
        if factored and is_small_slice(rows, cols):
            small.append(np.stack([np.full(batch, i, np.int32), b], 1))
            small_floats = max(small_floats, rows * cols + rows + cols)
            continue
        if not factored:
            rows, cols = (cols + TILE_C - 1) // TILE_C, TILE_C
        r0 = np.arange(0, max(rows, 1), TILE_R, dtype=np.int32)
        c0 = np.arange(0, max(cols, 1), TILE_C, dtype=np.int32)
        bb, rr, cc = np.meshgrid(b, r0, c0, indexing="ij")
        t = np.stack([np.full(bb.size, i, np.int32), bb.ravel(), rr.ravel(), cc.ravel()], 1)
        tiles.append(t)
        if factored:
            slices.append(np.stack([np.full(batch, i, np.int32), b], 1))
    tiles = np.concatenate(tiles, 0) if tiles else np.zeros((0, 4), np.int32)
    slices = np.concatenate(slices, 0) if slices else np.zeros((0, 2), np.int32)
    small = np.concatenate(small, 0) if small else np.zeros((0, 2), np.int32)
    return np.ascontiguousarray(tiles), np.ascontiguousarray(slices), np.ascontiguousarray(small), int(small_floats)


_TENSOR_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("row", "<u8"), ("col", "<u8"), ("row_acc", "<u8"),
                          ("col_acc", "<u8"), ("rmean", "<u8"), ("sumsq", "<u8"), ("batch", "<i8"), ("rows", "<i8"),
                          ("cols", "<i8"), ("numel", "<i8"), ("factored", "<i4"), ("pad_", "<i4")])
assert _TENSOR_DTYPE.itemsize == ctypes.sizeof(_lib.SmxAdafactorTensor)


class _Plan:
    """Static part of one launch: tile / slice tables on the device, the scratch block, and the tensor table in pinned
    host memory with everything but the gradient pointers filled in once (per step only the ``g`` column changes)."""

    def __init__(self, params, states):
        dev = params[0].device
        self.shapes = [tuple(p.shape) for p in params]
        tiles, slices, small, self.small_floats = tile_table(self.shapes)
        self.n_tiles, self.n_slices, self.n_small = int(tiles.shape[0]), int(slices.shape[0]), int(small.shape[0])
        pad = torch.zeros(1, 4, dtype=torch.int32, device=dev)
        self.tiles = torch.from_numpy(tiles).to(dev) if self.n_tiles else pad
        self.slices = torch.from_numpy(slices).to(dev) if self.n_slices else pad
        self.small = torch.from_numpy(small).to(dev) if self.n_small else pad
        # scratch (floats): per tensor row_acc [batch*rows] | col_acc [batch*cols] | sumsq [1]; rmean separately
        dims = [factored_dims(shape) for shape in self.shapes]
        n = sum((b * r + b * c if f else 0) + 1 for f, b, r, c in dims)
        m = sum(b for _, b, _, _ in dims)
        self.scratch = torch.zeros(n, dtype=torch.float32, device=dev)
        self.rmean = torch.zeros(max(m, 1), dtype=torch.float32, device=dev)
        self.pinned = torch.zeros(len(params) * _TENSOR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.host = self.pinned.numpy().view(_TENSOR_DTYPE)     # structured view of the pinned block
        self.table = torch.empty_like(self.pinned, device=dev)
        self.uploaded = None   # event after the last host -> device copy of the table (the pinned block is reused)
        # (pinned snapshot, device table) used by a CUDA-graph capture of step(): allocated here, outside any capture
        self.captured = (torch.zeros_like(self.pinned).pin_memory(), torch.empty_like(self.table))
        sc, rm, o, om = self.scratch.data_ptr(), self.rmean.data_ptr(), 0, 0
        for i, (p, st, (f, b, r, c)) in enumerate(zip(params, states, dims)):
            e = self.host[i]
            e["p"] = p.data_ptr()
            e["row"] = (st["exp_avg_sq_row"] if f else st["exp_avg_sq"]).data_ptr()
            e["col"] = st["exp_avg_sq_col"].data_ptr() if f else 0
            ra, ca = (b * r, b * c) if f else (0, 0)
            e["row_acc"], e["col_acc"], e["sumsq"], e["rmean"] = sc + 4 * o, sc + 4 * (o + ra), sc + 4 * (o + ra + ca), rm + 4 * om
            e["batch"], e["rows"], e["cols"], e["numel"], e["factored"] = b, r, c, p.numel(), 1 if f else 0
            o += ra + ca + 1
            om += b
        self.state_ptrs = [int(self.host[i]["row"]) for i in range(len(params))]   # to notice re-allocated state


class FusedAdafactor(torch.optim.Optimizer):
    """Drop-in for ``transformers.optimization.Adafactor(params, lr=lr, scale_parameter=False, relative_step=False)``."""

    def __init__(self, params, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                 weight_decay=0.0, scale_parameter=False, relative_step=False, warmup_init=False, capturable=False):
        if lr is None or relative_step or warmup_init or scale_parameter or beta1 is not None:
            raise NotImplementedError("FusedAdafactor implements the reference recipe only: explicit lr, "
                                      "relative_step=False, scale_parameter=False, warmup_init=False, beta1=None")
        defaults = dict(lr=lr, eps=eps, clip_threshold=clip_threshold, decay_rate=decay_rate, beta1=beta1,
                        weight_decay=weight_decay, scale_parameter=scale_parameter, relative_step=relative_step,
                        warmup_init=warmup_init, capturable=bool(capturable))
        super().__init__(params, defaults)
        self._plans = {}
        # capturable=True (whole-step CUDA graph, graph.GraphedTrainStep): ONE step counter on the device drives
        # beta2(t) inside the kernels and is bumped by the launch itself, so replays advance the schedule; all
        # parameters must then share one step count, learning rate and gradient addresses (true for a captured step).
        self._dev_step = None

    def _init_state(self, p):
        st = self.state[p]
        factored, batch, rows, cols = factored_dims(p.shape)
        st["step"] = 0
        if factored:
            st["exp_avg_sq_row"] = torch.zeros(p.shape[:-1], dtype=torch.float32, device=p.device)
            st["exp_avg_sq_col"] = torch.zeros(p.shape[:-2] + p.shape[-1:], dtype=torch.float32, device=p.device)
        else:
            st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
        st["RMS"] = 0   # only read by scale_parameter / relative_step, kept for checkpoint compatibility
        return st

    def _plan_for(self, params):
        key = tuple(id(p) for p in params)
        plan = self._plans.get(key)
        states = [self.state[p] for p in params]
        if plan is not None:   # load_state_dict / .to() may have replaced parameter or state storage
            rows = [(st["exp_avg_sq_row"] if "exp_avg_sq_row" in st else st["exp_avg_sq"]).data_ptr() for st in states]
            if rows != plan.state_ptrs or int(plan.host[0]["p"]) != params[0].data_ptr():
                plan = None
        if plan is None:
            if len(self._plans) >= 8:   # the trainable set changed repeatedly (gradual unfreezing): start over
                self._plans.clear()
            plan = self._plans[key] = _Plan(params, states)
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adafactor does not support sparse gradients.")
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdafactor needs contiguous fp32 CUDA parameters (no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    st = self._init_state(p)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append(p)
            for step_no, params in by_step.items():   # one launch per distinct step count (normally one)
                plan = self._plan_for(params)
                grads = [p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous())
                         else p.grad.float().contiguous() for p in params]
                capturing = torch.cuda.is_current_stream_capturing()
                if plan.uploaded is not None and not capturing:
                    plan.uploaded.synchronize()
                plan.host["g"] = [g.data_ptr() for g in grads]      # the only per-step column of the table
                table = plan.table
                if capturing:
                    # a captured step replays this host -> device copy: give the graph its OWN pinned snapshot and device
                    # table, so that eager steps taken later (other gradient addresses) cannot change what it uploads
                    plan.captured[0].copy_(plan.pinned)              # host-side copy (buffers pre-allocated in _Plan)
                    table = plan.captured[1]
                    table.copy_(plan.captured[0], non_blocking=True)
                else:
                    table.copy_(plan.pinned, non_blocking=True)
                    plan.uploaded = torch.cuda.Event()
                    plan.uploaded.record()
                beta2t = 1.0 - math.pow(step_no, group["decay_rate"])
                step_dev = 0
                if group["capturable"]:
                    if len(by_step) != 1:
                        raise RuntimeError("FusedAdafactor(capturable=True): all parameters must share one step count")
                    if self._dev_step is None:
                        self._dev_step = torch.full((1,), step_no - 1, dtype=torch.int64, device=params[0].device)
                    step_dev = self._dev_step.data_ptr()
                rc = lib.smx_adafactor_step(table.data_ptr(), len(params), plan.tiles.data_ptr(), plan.n_tiles,
                                            plan.slices.data_ptr(), plan.n_slices, plan.small.data_ptr(), plan.n_small,
                                            plan.small_floats, plan.scratch.data_ptr(),
                                            plan.scratch.numel() * 4, beta2t, group["eps"][0], group["lr"],
                                            group["clip_threshold"], group["weight_decay"], step_dev,
                                            group["decay_rate"], torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "smx_adafactor_step")
                del grads
        return loss

    def steps_done(self):
        """capturable mode: the device-side step count (graph replays do not pass through ``step()``)."""
        return int(self._dev_step.item()) if self._dev_step is not None else None
