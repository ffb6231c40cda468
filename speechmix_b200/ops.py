"""Autograd functions of the hot path.  Every forward AND backward here is a
sequence of calls into libspeechmix_sm100.so (``kernels.py``); torch is used for
tensor allocation, views and the autograd graph only.

Parameters stay fp32 ``nn.Parameter``s (state_dict-compatible with the
reference); bf16 / packed copies are cached per parameter version
(``WeightCache``) and refreshed after an optimizer step.  Activations are bf16.
Weight gradients come out of the TN GEMM in fp32.
"""
import math
import os
import weakref

import torch

from . import kernels as K
from .kernels import ACT_DGELU, ACT_DRELU, ACT_GELU, ACT_GELU_G, ACT_MULAUX, ACT_NONE, ACT_RELU, BF16


class WeightCache:
    """bf16 / packed working copies of the fp32 parameters.

    An entry is valid while (a) the cache epoch has not moved, (b) the source tensors are the same
    objects at the same address with the same ``_version``.  (b) alone is NOT enough: fused optimizers
    (``torch.optim.AdamW(fused=True)``, ``_fused_adamw_``) update parameters without bumping
    ``_version``, so every training forward starts a new epoch (``new_step``) and refreshes all plain
    casts in ONE multi-tensor launch (``smx_multi_cast``); packed layouts (conv tap-major, pos-conv
    core-matrix order) are rebuilt lazily on first use in the epoch.  Entries hold weak references,
    so a recycled ``id()`` can never alias a dead tensor."""

    def __init__(self):
        self._store = {}
        self.epoch = 0
        self.dirty = False  # an optimizer step ran since the last refresh (set by the hook below / graph replays)
        self._tables = {}   # device -> (signature, device table tensor, n_entries, total_chunks)

    def get(self, key_tensors, kind, build, simple=None):
        if not isinstance(key_tensors, (tuple, list)):
            key_tensors = (key_tensors,)
        key = (kind,) + tuple(id(t) for t in key_tensors)
        ver = tuple((t._version, t.data_ptr()) for t in key_tensors)
        hit = self._store.get(key)
        if hit is not None and hit[3] == self.epoch and hit[0] == ver and \
                all(r() is t for r, t in zip(hit[2], key_tensors)):
            return hit[1]
        with torch.no_grad():
            val = build(*key_tensors)
        if len(self._store) > 4096:
            self._store = {k: v for k, v in self._store.items() if all(r() is not None for r in v[2])}
        self._store[key] = [ver, val, tuple(weakref.ref(t) for t in key_tensors), self.epoch,
                            simple(val, key_tensors) if simple is not None else None]
        return val

    def invalidate(self):
        """Forget every copy (lazy rebuild); called on train()/eval() switches."""
        self.epoch += 1

    def new_step(self):
        """Start a new epoch and refresh every plain cast from its fp32 master in one launch."""
        self.epoch += 1
        self.dirty = False
        per_dev = {}
        for key, ent in list(self._store.items()):
            spec = ent[4]
            if spec is None:
                continue
            srcs = [r() for r in ent[2]]
            if any(t is None for t in srcs):
                del self._store[key]
                continue
            ok = all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
                     for t, (_, n, _) in zip(srcs, spec))
            if not ok:
                continue
            per_dev.setdefault(srcs[0].device, []).append((ent, srcs, spec))
        for dev, items in per_dev.items():
            rows = []
            for ent, srcs, spec in items:
                for t, (dst, n, f32) in zip(srcs, spec):
                    if n:
                        rows.append((t.data_ptr(), dst.data_ptr(), n, f32))
            sig = tuple(rows)
            tab = self._tables.get(dev)
            if tab is None or tab[0] != sig:
                tab = (sig,) + K.build_cast_table(rows, dev)
                self._tables[dev] = tab
            with torch.cuda.device(dev):
                K.multi_cast(tab[1], tab[2], tab[3])
            for ent, srcs, _ in items:
                ent[0] = tuple((t._version, t.data_ptr()) for t in srcs)
                ent[3] = self.epoch

    def clear(self):
        self._store.clear()
        self._tables.clear()


CACHE = WeightCache()


def _mark_weights_moved(optimizer, args, kwargs):
    """global optimizer-step post-hook: ANY optimizer (fused ones do not bump tensor versions) may have moved the fp32
    masters, so copies made before this point are stale for every later pass -- also a no_grad one."""
    CACHE.dirty = True
    CACHE.epoch += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_register_step_hook(_mark_weights_moved)


def _rows_spec(val, srcs, f32):
    out, r = [], 0
    for t in srcs:
        n = t.shape[0]
        out.append((val[r:r + n], t.numel(), f32))
        r += n
    return out


class fp32_verification:
    """Context manager: run the forward graph with fp32 activations and fp32 CUDA-core arithmetic (csrc/fp32.cu)
    -- inference only; used to check greedy-decoded ids bit for bit against the reference's fp32 run."""

    def __enter__(self):
        if torch.is_grad_enabled():
            raise RuntimeError("fp32 verification mode is inference-only: wrap the call in torch.no_grad()")
        self._prev = K.FP32_MODE
        K.FP32_MODE = True
        return self

    def __exit__(self, *exc):
        K.FP32_MODE = self._prev
        return False


class DropoutState:
    """Device-resident {seed, step} pair behind every dropout mask (csrc/sm100_prims.cuh ``drop_key``) plus the host-side
    call counter that numbers the dropout sites of one forward pass in execution order (the same order in which the HF
    modules call ``F.dropout``, which is what lets the parity tests feed OUR masks to the CPU reference run).

    ``begin_step()`` -- called by the model at the start of every training forward -- advances ``step`` ON THE DEVICE, so
    a replayed CUDA graph (which re-executes that increment) draws fresh masks on every replay; the backward pass of the
    same step regenerates the masks from the unchanged pair."""

    def __init__(self):
        self._state = {}
        self.calls = 0
        self.seed = 0x5eed
        self.trace = None     # tests: list of (call, p, kind, shape) of the sites of a forward pass, in call order

    def state(self, device):
        device = torch.device(device)
        st = self._state.get(device)
        if st is None:
            st = torch.tensor([self.seed, 0], dtype=torch.int64, device=device)
            self._state[device] = st
        return st

    def manual_seed(self, seed):
        self.seed = int(seed)
        for st in self._state.values():
            st.copy_(torch.tensor([self.seed, 0], dtype=torch.int64))

    def begin_step(self, device):
        self.state(device)[1:2].add_(1)
        self.calls = 0

    def next_call(self, p=None, kind=None, shape=None):
        self.calls += 1
        if self.trace is not None:
            self.trace.append((self.calls - 1, p, kind, tuple(shape) if shape is not None else None))
        return self.calls - 1


DROPOUT = DropoutState()


class DropoutFn(torch.autograd.Function):
    """y = keep ? x / (1 - p) : 0 at an elementwise dropout site (embedding / encoder-level dropout of the HF modules)."""

    @staticmethod
    def forward(ctx, x, p):
        ctx.call, ctx.p, ctx.state = DROPOUT.next_call(p, "elementwise", x.shape), p, DROPOUT.state(x.device)
        return K.dropout(x, ctx.state, ctx.call, p)

    @staticmethod
    def backward(ctx, dy):
        return K.dropout(dy, ctx.state, ctx.call, ctx.p), None


def dropout(x, p, training):
    if not training or p <= 0.0 or K.FP32_MODE:
        return x
    return DropoutFn.apply(x, float(p))


def _drop_site(p, device, kind="elementwise", shape=None):
    """(state, call, p) of a dropout site inside a fused block, or None when it is off"""
    if p is None or p <= 0.0 or K.FP32_MODE:
        return None
    return (DROPOUT.state(device), DROPOUT.next_call(float(p), kind, shape), float(p))


def w16(p):
    if K.FP32_MODE:
        return p.detach()
    return CACHE.get(p, "bf16", lambda t: K.to_bf16(t), simple=lambda v, ts: [(v, ts[0].numel(), 0)])


def cat16(ps):
    if K.FP32_MODE:
        return cat32(ps)
    return CACHE.get(tuple(ps), "cat16", lambda *ts: K.to_bf16(torch.cat([t.detach() for t in ts], 0)),
                     simple=lambda v, ts: _rows_spec(v, ts, 0))


def cat32(ps):
    return CACHE.get(tuple(ps), "cat32", lambda *ts: torch.cat([t.detach().float() for t in ts], 0).contiguous(),
                     simple=lambda v, ts: _rows_spec(v, ts, 1))


def conv_packed16(p):
    return CACHE.get(p, "convpack32" if K.FP32_MODE else "convpack", lambda t: K.pack_conv_weight(t))


_DROPOUT_TAIL = os.environ.get("SMX_DROPOUT_TAIL", "1") != "0"   # A/B switch (development): "0" = the separate launches


def _fused_dropout_tail(drop_h, pre_ln, rms):
    """post-LN block with output dropout: dropout + residual add + LayerNorm run as one launch each way
    (smx_layernorm_dropout_fwd / _bwd) instead of three / four"""
    return drop_h is not None and not pre_ln and not rms and not K.FP32_MODE and _DROPOUT_TAIL


def _act_codes(name):
    """(forward epilogue code, backward epilogue code).  For GELU the forward GEMM stores gelu'(pre) as its
    auxiliary output (ACT_GELU_G) so the data-gradient GEMM's epilogue is one multiply (ACT_MULAUX) -- the
    derivative costs ~22 issue slots + 2 MUFU per element, which made those epilogues slower than their
    K = 768 main loops."""
    if name in ("gelu", "gelu_new"):
        return (ACT_GELU if K.FP32_MODE else ACT_GELU_G), ACT_MULAUX
    if name == "relu":
        return ACT_RELU, ACT_DRELU
    raise ValueError("unsupported activation %r" % (name,))


def _need(ctx, i):
    return ctx.needs_input_grad[i]


# ---------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, rms_only):
        x = x.contiguous()
        y, _, mean, rstd = K.layernorm_fwd(x, gamma.detach(), None if beta is None else beta.detach(), eps,
                                           rms_only=rms_only)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.rms_only = rms_only
        ctx.has_beta = beta is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx, dg, db = K.layernorm_bwd(dy.contiguous(), x, gamma.detach(), mean, rstd, rms_only=ctx.rms_only,
                                     want_dbeta=ctx.has_beta)
        return dx, dg, db, None, None


def layer_norm(x, gamma, beta, eps=1e-5, rms_only=False):
    return LayerNormFn.apply(x, gamma, beta, eps, rms_only)


# ---------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ residual)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        wb = w16(weight)
        r2 = residual.reshape(-1, weight.shape[0]) if residual is not None else None
        y = K.linear_fwd(x2, wb, None if bias is None else bias.detach(), residual=r2)
        ctx.save_for_backward(x2, wb)
        ctx.shp = shp
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = K.linear_dgrad(dy2, wb).view(ctx.shp) if _need(ctx, 0) else None
        dw = K.linear_wgrad(dy2, x2) if _need(ctx, 1) else None
        db = K.colsum(dy2) if (ctx.has_bias and _need(ctx, 2)) else None
        dres = dy if ctx.has_res else None
        return dx, dw, db, dres


def linear(x, weight, bias=None, residual=None):
    return LinearFn.apply(x, weight, bias, residual)


# ---------------------------------------------------------------------------
class MaskRowsFn(torch.autograd.Function):
    """y[b, t, :] = x[b, t, :] for t < lens[b], else 0 ("make sure padded tokens output 0",
    hf:models/wav2vec2/modeling_wav2vec2.py:672-675); the gradient of the zeroed rows is zero."""

    @staticmethod
    def forward(ctx, x, lens):
        ctx.lens = lens
        return K.mask_rows(x.clone(memory_format=torch.contiguous_format), lens)

    @staticmethod
    def backward(ctx, dy):
        return K.mask_rows(dy.clone(memory_format=torch.contiguous_format), ctx.lens), None


def mask_rows(x, lens):
    return MaskRowsFn.apply(x, lens)


class SpecAugmentFn(torch.autograd.Function):
    """hf:models/wav2vec2/modeling_wav2vec2.py:1280-1324: time-masked frames are replaced by ``masked_spec_embed``,
    feature-masked channels are cleared.  The masks (uint8, on the device) are drawn on the host by the caller."""

    @staticmethod
    def forward(ctx, x, embed, time_mask, feat_mask):
        ctx.masks = (time_mask, feat_mask)
        return K.spec_augment_fwd(x.contiguous(), time_mask, feat_mask, embed.detach().float().contiguous())

    @staticmethod
    def backward(ctx, dy):
        time_mask, feat_mask = ctx.masks
        dx, dembed = K.spec_augment_bwd(dy.contiguous(), time_mask, feat_mask, want_dembed=ctx.needs_input_grad[1])
        return dx, dembed, None, None


def spec_augment(x, embed, time_mask, feat_mask):
    return SpecAugmentFn.apply(x, embed, time_mask, feat_mask)


# ---------------------------------------------------------------------------
class AttnBlockFn(torch.autograd.Function):
    """(self- or cross-) attention sub-block with its residual and LayerNorm.

      post-LN:  y = LN(x + Wo.attn(q(x), kv(src)) + bo)          (wav2vec2-base, BART)
      pre-LN :  y = x + Wo.attn(q(LN(x)), kv(LN(x) | src)) + bo   (stable-LN wav2vec2/HuBERT, mBART)

    hf:models/wav2vec2/modeling_wav2vec2.py:466-549,576-655 ; hf:models/bart/modeling_bart.py:143-391 ;
    hf:models/mbart/modeling_mbart.py:274-430.
    inputs: x, src (None for self-attention), cfg, q_w,q_b,k_w,k_b,v_w,v_b,o_w,o_b, ln_w, ln_b
    cfg["kv_len"] (optional int32 [B] on the device): key-padding mask as per-sample key counts -- the keys / values
    of the padded frames are zeroed in the projection output and ignored by the attention kernels, their
    gradients are zero (hf:...wav2vec2.py:1026-1044 + create_bidirectional_mask).
    """

    @staticmethod
    def forward(ctx, x, src, cfg, q_w, q_b, k_w, k_b, v_w, v_b, o_w, o_b, ln_w, ln_b, pos_bias=None):
        heads, causal, pre_ln, eps = cfg["heads"], cfg["causal"], cfg["pre_ln"], cfg["eps"]
        rms = bool(cfg.get("rms", False))          # T5: RMSNorm (no mean, no beta)
        scale = cfg.get("scale", 1.0 / math.sqrt(64))
        kv_len = cfg.get("kv_len")
        B, T, H = x.shape
        Hi = q_w.shape[0]                          # heads * 64 (== H except for some T5 sizes)
        x2 = x.reshape(B * T, H)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ln_b_d = None if ln_b is None else ln_b.detach()
        if pre_ln:
            n, _, mean, rstd = K.layernorm_fwd(x2, ln_w.detach(), ln_b_d, eps, rms_only=rms)
            a_in = n
        else:
            a_in = x2
        if src is None:
            wqkv = cat16((q_w, k_w, v_w))
            bqkv = cat32((q_b, k_b, v_b)) if q_b is not None else None
            qkv = K.linear_fwd(a_in, wqkv, bqkv).view(B, T, 3 * Hi)
            if kv_len is not None:
                K.mask_rows(qkv, kv_len, col_begin=Hi, col_count=2 * Hi)
            q, k, v = qkv[..., :Hi], qkv[..., Hi:2 * Hi], qkv[..., 2 * Hi:]
            kv_src2 = None
            Ts = T
        else:
            Bs, Ts, Hs = src.shape
            src2 = src.reshape(Bs * Ts, Hs)
            if not src2.is_contiguous():
                src2 = src2.contiguous()
            q = K.linear_fwd(a_in, w16(q_w), None if q_b is None else q_b.detach()).view(B, T, Hi)
            wkv = cat16((k_w, v_w))
            bkv = cat32((k_b, v_b)) if k_b is not None else None
            kv = K.linear_fwd(src2, wkv, bkv).view(Bs, Ts, 2 * Hi)
            if kv_len is not None:
                K.mask_rows(kv, kv_len)
            k, v = kv[..., :Hi], kv[..., Hi:]
            qkv = (q, kv)
            kv_src2 = src2
        pb = None if pos_bias is None else pos_bias.detach().float().contiguous()
        # train-mode dropout of the HF blocks, in their call order: attention probabilities, then the block output
        # before the residual add (hf:...wav2vec2.py:466-549,590-596 ; hf:...bart.py:143-258,280-290)
        drop_a = _drop_site(cfg.get("p_attn"), x.device, "attention", (B, heads, T, Ts))
        drop_h = _drop_site(cfg.get("p_hidden"), x.device, "elementwise", (B, T, H))
        o, lse = K.attn_fwd(q, k, v, heads, causal=causal, scale=scale, bias=pb, kv_len=kv_len, dropout=drop_a)
        fused_tail = _fused_dropout_tail(drop_h, pre_ln, rms)
        if drop_h is None:
            s = K.linear_fwd(o.view(B * T, Hi), w16(o_w), None if o_b is None else o_b.detach(), residual=x2)
        else:
            s0 = K.linear_fwd(o.view(B * T, Hi), w16(o_w), None if o_b is None else o_b.detach())
            if fused_tail:       # dropout + residual add + LayerNorm of a post-LN block: one launch
                y, s, mean, rstd = K.layernorm_dropout_fwd(s0, x2, ln_w.detach(), ln_b_d, eps, *drop_h)
            else:
                s = K.dropout(s0, *drop_h, residual=x2)
        if pre_ln:
            y = s
            ctx.save_for_backward(x2, n, mean, rstd, o, lse, kv_src2, ln_w, pb, *(qkv if src is not None else (qkv,)))
        else:
            if not fused_tail:
                y, _, mean, rstd = K.layernorm_fwd(s, ln_w.detach(), ln_b_d, eps, rms_only=rms)
            ctx.save_for_backward(x2, s, mean, rstd, o, lse, kv_src2, ln_w, pb, *(qkv if src is not None else (qkv,)))
        ctx.cfg = dict(cfg, scale=scale, rms=rms)
        ctx.drop_a, ctx.drop_h = drop_a, drop_h
        ctx.kv_len = kv_len
        ctx.dims = (B, T, H, Ts, Hi)
        ctx.cross = src is not None
        ctx.wrefs = (q_w, k_w, v_w, o_w)
        ctx.has_bias = q_b is not None
        ctx.has_obias = o_b is not None
        ctx.has_lnb = ln_b is not None
        return y.view(B, T, H)

    @staticmethod
    def backward(ctx, dy):
        cfg = ctx.cfg
        heads, causal, pre_ln, scale, rms = cfg["heads"], cfg["causal"], cfg["pre_ln"], cfg["scale"], cfg["rms"]
        B, T, H, Ts, Hi = ctx.dims
        q_w, k_w, v_w, o_w = ctx.wrefs
        sv = ctx.saved_tensors
        x2, a, mean, rstd, o, lse, kv_src2, ln_w, pb = sv[:9]
        dy2 = dy.reshape(B * T, H).contiguous()
        fused_tail = _fused_dropout_tail(ctx.drop_h, pre_ln, rms)
        if pre_ln:
            ds = dy2
            a_in = a          # normalised input
        elif fused_tail:      # LayerNorm backward + the dropout mask on its result + the out-proj bias gradient: one launch
            ds_res, ds, dlnw, dlnb, d_ob = K.layernorm_dropout_bwd(dy2, a, ln_w.detach(), mean, rstd, *ctx.drop_h,
                                                                   want_dbeta=ctx.has_lnb, want_colsum=ctx.has_obias)
            a_in = x2
        else:
            ds, dlnw, dlnb, d_ob = K.layernorm_bwd(dy2, a, ln_w.detach(), mean, rstd, rms_only=rms, want_dbeta=ctx.has_lnb,
                                                   want_colsum=True)   # colsum(ds) = out-proj bias gradient, fused
            a_in = x2
        o2 = o.view(B * T, Hi)
        if not fused_tail:
            ds_res = ds                      # gradient of the residual branch (not dropped)
            if ctx.drop_h is not None:       # gradient through the block-output dropout: the same mask on ds
                ds = K.dropout(ds, *ctx.drop_h)
            if pre_ln or not ctx.has_obias or ctx.drop_h is not None:
                d_ob = K.colsum(ds) if ctx.has_obias else None
        d_ow = K.linear_wgrad(ds, o2) if _need(ctx, 9) else None
        do = K.linear_dgrad(ds, w16(o_w)).view(B, T, Hi)
        ds = ds_res
        dsrc = None
        dpb = K.zeros_f32(*pb.shape, device=pb.device) if (pb is not None and _need(ctx, 13)) else None
        if not ctx.cross:
            qkv = sv[9]
            q, k, v = qkv[..., :Hi], qkv[..., Hi:2 * Hi], qkv[..., 2 * Hi:]
            dqkv = torch.empty_like(qkv)
            K.attn_bwd(do, q, k, v, o, lse, heads, causal=causal, scale=scale, bias=pb, dq=dqkv[..., :Hi],
                       dk=dqkv[..., Hi:2 * Hi], dv=dqkv[..., 2 * Hi:], dbias=dpb, kv_len=ctx.kv_len, dropout=ctx.drop_a)
            dqkv2 = dqkv.view(B * T, 3 * Hi)
            need_w = _need(ctx, 3) or _need(ctx, 5) or _need(ctx, 7)
            dwqkv = K.linear_wgrad(dqkv2, a_in) if need_w else None
            dbqkv = K.colsum(dqkv2) if (ctx.has_bias and need_w) else None
            wqkv = cat16((q_w, k_w, v_w))
            d_in = K.linear_dgrad(dqkv2, wqkv, residual=None if pre_ln else ds)
            dq_w, dk_w, dv_w = (dwqkv[:Hi], dwqkv[Hi:2 * Hi], dwqkv[2 * Hi:]) if need_w else (None, None, None)
            dq_b, dk_b, dv_b = (dbqkv[:Hi], dbqkv[Hi:2 * Hi], dbqkv[2 * Hi:]) if dbqkv is not None else (None, None, None)
        else:
            q, kv = sv[9], sv[10]
            k, v = kv[..., :Hi], kv[..., Hi:]
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            K.attn_bwd(do, q, k, v, o, lse, heads, causal=causal, scale=scale, bias=pb, dq=dq, dk=dkv[..., :Hi],
                       dv=dkv[..., Hi:], dbias=dpb, kv_len=ctx.kv_len, dropout=ctx.drop_a)
            dq2 = dq.view(B * T, Hi)
            dkv2 = dkv.view(-1, 2 * Hi)
            need_q = _need(ctx, 3)
            need_kv = _need(ctx, 5) or _need(ctx, 7)
            dq_w = K.linear_wgrad(dq2, a_in) if need_q else None
            dq_b = K.colsum(dq2) if (ctx.has_bias and need_q) else None
            dwkv = K.linear_wgrad(dkv2, kv_src2) if need_kv else None
            dbkv = K.colsum(dkv2) if (ctx.has_bias and need_kv) else None
            dk_w, dv_w = (dwkv[:Hi], dwkv[Hi:]) if need_kv else (None, None)
            dk_b, dv_b = (dbkv[:Hi], dbkv[Hi:]) if dbkv is not None else (None, None)
            d_in = K.linear_dgrad(dq2, w16(q_w), residual=None if pre_ln else ds)
            if _need(ctx, 1):
                dsrc = K.linear_dgrad(dkv2, cat16((k_w, v_w))).view(-1, Ts, kv_src2.shape[1])
        if pre_ln:
            dx, dlnw, dlnb = K.layernorm_bwd(d_in, x2, ln_w.detach(), mean, rstd, dres=ds, rms_only=rms,
                                             want_dbeta=ctx.has_lnb)
        else:
            dx = d_in
        return (dx.view(B, T, H), dsrc, None, dq_w, dq_b, dk_w, dk_b, dv_w, dv_b, d_ow, d_ob, dlnw, dlnb, dpb)


class FFNBlockFn(torch.autograd.Function):
    """post-LN: y = LN(x + W2.act(W1 x + b1) + b2);  pre-LN: y = x + W2.act(W1 LN(x) + b1) + b2.
    hf:...wav2vec2.py:552-573 ; hf:...bart.py:296-309."""

    @staticmethod
    def forward(ctx, x, cfg, w1, b1, w2, b2, ln_w, ln_b):
        pre_ln, eps = cfg["pre_ln"], cfg["eps"]
        act, dact = _act_codes(cfg.get("act", "gelu"))
        shp = x.shape
        H = shp[-1]
        x2 = x.reshape(-1, H)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rms = bool(cfg.get("rms", False))
        ln_b_d = None if ln_b is None else ln_b.detach()
        if pre_ln:
            n, _, mean, rstd = K.layernorm_fwd(x2, ln_w.detach(), ln_b_d, eps, rms_only=rms)
            a_in = n
        else:
            a_in = x2
        h, pre = K.linear_fwd(a_in, w16(w1), None if b1 is None else b1.detach(), act=act, want_pre=True)
        no_res = bool(cfg.get("no_residual", False))  # adapter blocks: y = W2.act(W1 LN(x) + b1) + b2
        # train-mode dropout of the HF feed-forward blocks, in their call order: after the activation, then on the block
        # output before the residual add (hf:...wav2vec2.py:552-573 ; hf:...bart.py:296-309 ; hf:...t5.py:80-101)
        drop_act = _drop_site(cfg.get("p_act"), x.device, "elementwise", tuple(shp[:-1]) + (w1.shape[0],))
        drop_h = _drop_site(cfg.get("p_hidden"), x.device, "elementwise", shp)
        if drop_act is not None:
            # the dropped activation feeds fc2 (and its weight gradient); the stored activation-gradient operand becomes
            # the masked multiplier keep * act'(pre) / (1 - p), so the backward epilogue stays a single multiply
            h, pre = K.dropout(h, *drop_act, aux_in=pre, aux_mode=2 if dact == ACT_DRELU else 1)
            dact = ACT_MULAUX
        fused_tail = _fused_dropout_tail(drop_h, pre_ln, rms) and not no_res
        if drop_h is None:
            s = K.linear_fwd(h, w16(w2), None if b2 is None else b2.detach(), residual=None if no_res else x2)
        else:
            s0 = K.linear_fwd(h, w16(w2), None if b2 is None else b2.detach())
            if fused_tail:       # dropout + residual add + LayerNorm of a post-LN block: one launch
                y, s, mean, rstd = K.layernorm_dropout_fwd(s0, x2, ln_w.detach(), ln_b_d, eps, *drop_h)
            else:
                s = K.dropout(s0, *drop_h, residual=None if no_res else x2)
        if pre_ln:
            y = s
            ctx.save_for_backward(x2, n, mean, rstd, pre, h, ln_w)
        else:
            if not fused_tail:
                y, _, mean, rstd = K.layernorm_fwd(s, ln_w.detach(), ln_b_d, eps, rms_only=rms)
            ctx.save_for_backward(x2, s, mean, rstd, pre, h, ln_w)
        ctx.pre_ln, ctx.dact, ctx.shp = pre_ln, dact, shp
        ctx.rms, ctx.has_lnb = rms, ln_b is not None
        ctx.no_res = no_res
        ctx.drop_h = drop_h
        ctx.fused_tail = fused_tail
        ctx.wrefs = (w1, w2)
        ctx.has_bias = b1 is not None
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, a, mean, rstd, pre, h, ln_w = ctx.saved_tensors
        w1, w2 = ctx.wrefs
        H = ctx.shp[-1]
        dy2 = dy.reshape(-1, H).contiguous()
        if ctx.pre_ln:
            ds, a_in = dy2, a
        elif ctx.fused_tail:  # LayerNorm backward + the dropout mask on its result + the fc2 bias gradient: one launch
            ds_res, ds, dlnw, dlnb, db2 = K.layernorm_dropout_bwd(dy2, a, ln_w.detach(), mean, rstd, *ctx.drop_h,
                                                                  want_dbeta=ctx.has_lnb, want_colsum=ctx.has_bias)
            a_in = x2
        else:
            ds, dlnw, dlnb, db2 = K.layernorm_bwd(dy2, a, ln_w.detach(), mean, rstd, rms_only=ctx.rms, want_dbeta=ctx.has_lnb,
                                                  want_colsum=True)    # colsum(ds) = fc2 bias gradient, fused
            a_in = x2
        if not ctx.fused_tail:
            ds_res = ds
            if ctx.drop_h is not None:       # gradient through the block-output dropout
                ds = K.dropout(ds, *ctx.drop_h)
            if ctx.pre_ln or not ctx.has_bias or ctx.drop_h is not None:
                db2 = K.colsum(ds) if ctx.has_bias else None
        dw2 = K.linear_wgrad(ds, h) if _need(ctx, 4) else None
        dpre = K.linear_dgrad(ds, w16(w2), act=ctx.dact, aux_in=pre)
        ds = ds_res
        db1 = K.colsum(dpre) if ctx.has_bias else None
        dw1 = K.linear_wgrad(dpre, a_in) if _need(ctx, 2) else None
        d_in = K.linear_dgrad(dpre, w16(w1), residual=None if (ctx.pre_ln or ctx.no_res) else ds)
        if ctx.pre_ln:
            dx, dlnw, dlnb = K.layernorm_bwd(d_in, x2, ln_w.detach(), mean, rstd, dres=None if ctx.no_res else ds,
                                             rms_only=ctx.rms, want_dbeta=ctx.has_lnb)
        else:
            dx = d_in
        return dx.view(ctx.shp), None, dw1, db1, dw2, db2, dlnw, dlnb


# ---------------------------------------------------------------------------
class FeatureEncoderGroupFn(torch.autograd.Function):
    """wav2vec2 / HuBERT conv feature encoder, feat_extract_norm="group", conv_bias=False:
    conv0 + GroupNorm + GELU (fused, CUDA cores), then k=3/2 stride-2 convs + GELU as
    implicit tcgen05 GEMMs.  hf:...wav2vec2.py:254-323, 382-419.
    inputs: audio, kernel sizes, w0, gn_w, gn_b, w1..w6   ->  [B, T, C] bf16 (channels-last)"""

    @staticmethod
    def forward(ctx, audio, ks, w0, gn_w, gn_b, *ws):
        audio = audio.contiguous().float()
        need_bwd = bool(any(ctx.needs_input_grad))   # training: the forward also stores gelu'(z) for a streaming backward
        res = K.conv0_fwd(audio, w0.detach().contiguous(), gn_w.detach(), gn_b.detach(), want_gprime=need_bwd)
        y, stats, moments, gp0 = res if need_bwd else (*res, None)
        acts, pres = [y], []
        for w, k in zip(ws, ks):
            # `pre` holds gelu'(pre-activation) (ACT_GELU_G), consumed by ACT_MULAUX in backward
            y, pre = K.conv_s2_fwd(y, conv_packed16(w), k, act=ACT_GELU if K.FP32_MODE else ACT_GELU_G, want_pre=True)
            acts.append(y)
            pres.append(pre)
        ctx.save_for_backward(audio, w0, gn_w, gn_b, stats, moments, gp0, *acts[:-1], *pres)
        ctx.ks, ctx.ws, ctx.n = ks, ws, len(ws)
        return y

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        audio, w0, gn_w, gn_b, stats, moments, gp0 = sv[:7]
        n = ctx.n
        acts, pres = sv[7:7 + n], sv[7 + n:7 + 2 * n]
        dpre = K.dact(dy.contiguous(), pres[n - 1], ACT_MULAUX)
        dws = [None] * n
        for i in range(n - 1, -1, -1):
            w, k = ctx.ws[i], ctx.ks[i]
            x_in = acts[i]
            if ctx.needs_input_grad[5 + i]:
                dws[i] = K.unpack_conv_wgrad(K.conv_s2_wgrad(dpre, x_in, k), x_in.shape[2], k)
            if i > 0:
                dpre = K.conv_s2_dgrad(dpre, conv_packed16(w), k, x_in.shape[1], act=ACT_MULAUX, aux_in=pres[i - 1])
            else:
                dpre = K.conv_s2_dgrad(dpre, conv_packed16(w), k, x_in.shape[1])
        dw0, dg, db = K.conv0_bwd(audio, w0.detach().contiguous(), gn_w.detach(), gn_b.detach(), stats, moments, dpre, gp0)
        return (None, None, dw0, dg, db, *dws)


class FeatureEncoderLayerNormFn(torch.autograd.Function):
    """Conv feature encoder with feat_extract_norm="layer", conv_bias=True (HuBERT-large,
    wav2vec2-large-lv60): every layer is Conv1d + bias -> LayerNorm over channels -> GELU
    (hf:...wav2vec2.py:275-299).  Layer 0 is one fused CUDA-core kernel; layers 1.. are implicit
    tcgen05 GEMMs with the bias in the epilogue followed by a fused LayerNorm+GELU pass.
    inputs: audio, kernel sizes, (w, b, ln_w, ln_b) x n_layers  ->  [B, T, C] bf16"""

    @staticmethod
    def forward(ctx, audio, ks, *params):
        audio = audio.contiguous().float()
        n = len(params) // 4
        w0, b0, g0, be0 = params[:4]
        y = K.conv0_ln_fwd(audio, w0.detach().contiguous(), b0.detach(), g0.detach(), be0.detach())
        saved = [audio, y]
        for i in range(1, n):
            w, b, g, be = params[4 * i:4 * i + 4]
            z = K.conv_s2_fwd(y, conv_packed16(w), ks[i - 1], bias=b.detach())
            B, T, C = z.shape
            out = K.alloc_act(B, T, C, z.device)
            y, _, mean, rstd = K.layernorm_fwd(z.view(B * T, C), g.detach(), be.detach(), 1e-5, act=ACT_GELU,
                                               out=out.view(B * T, C))
            y = out
            saved += [z, mean, rstd, y]
        ctx.save_for_backward(*saved[:-1])  # the last activation is the output
        ctx.params, ctx.ks, ctx.n = params, ks, n
        return y

    @staticmethod
    def backward(ctx, dy):
        sv = ctx.saved_tensors
        audio, n, params = sv[0], ctx.n, ctx.params
        # layout of sv: audio, y0, (z1, mean1, rstd1, y1), ..., (z_{n-1}, mean, rstd)  [y_{n-1} not saved]
        grads = [None] * (4 * n)
        d = dy.contiguous()
        for i in range(n - 1, 0, -1):
            w, b, g, be = params[4 * i:4 * i + 4]
            base = 2 + 4 * (i - 1)
            z, mean, rstd = sv[base], sv[base + 1], sv[base + 2]
            x_in = sv[1] if i == 1 else sv[2 + 4 * (i - 2) + 3]
            B, T, C = z.shape
            dz, dg, dbe = K.layernorm_bwd(d.view(B * T, C), z.view(B * T, C), g.detach(), mean, rstd, act=ACT_GELU,
                                          beta=be.detach())
            dz3 = dz.view(B, T, C)
            k = ctx.ks[i - 1]
            grads[4 * i + 2], grads[4 * i + 3] = dg, dbe
            if ctx.needs_input_grad[2 + 4 * i + 1]:
                grads[4 * i + 1] = K.colsum(dz)
            if ctx.needs_input_grad[2 + 4 * i]:
                grads[4 * i] = K.unpack_conv_wgrad(K.conv_s2_wgrad(dz3, x_in, k), x_in.shape[2], k)
            d = K.conv_s2_dgrad(dz3, conv_packed16(w), k, x_in.shape[1])
        w0, b0, g0, be0 = params[:4]
        dw0, db0, dg0, dbe0 = K.conv0_ln_bwd(audio, w0.detach().contiguous(), b0.detach(), g0.detach(), be0.detach(), d)
        grads[0], grads[1], grads[2], grads[3] = dw0, db0, dg0, dbe0
        return (None, None, *grads)


class ConvS2Fn(torch.autograd.Function):
    """Conv1d(C -> N, k, stride 2, bias) on channels-last input, no activation
    (the down_scale length adapters, ref:speechmix/hf_model.py:253-266, 426-427)."""

    @staticmethod
    def forward(ctx, x, weight, bias, k):
        B, T, C = x.shape
        xa = K.alloc_act(B, T, C, x.device)
        xa.copy_(x)
        wp = conv_packed16(weight)
        y = K.conv_s2_fwd(xa, wp, k, bias=None if bias is None else bias.detach())
        ctx.save_for_backward(xa, wp)
        ctx.k = k
        return y

    @staticmethod
    def backward(ctx, dy):
        xa, wp = ctx.saved_tensors
        k = ctx.k
        B, T_out, N = dy.shape
        dya = K.alloc_act(B, T_out, N, dy.device)
        dya.copy_(dy)
        dx = K.conv_s2_dgrad(dya, wp, k, xa.shape[1]) if _need(ctx, 0) else None
        dw = K.unpack_conv_wgrad(K.conv_s2_wgrad(dya, xa, k), xa.shape[2], k) if _need(ctx, 1) else None
        db = K.colsum(dya.view(B * T_out, N)) if _need(ctx, 2) else None
        return dx, dw, db, None


class GatedFFNBlockFn(torch.autograd.Function):
    """T5 v1.1 / mT5 / flan-T5 feed-forward (hf:models/t5/modeling_t5.py T5DenseGatedActDense + T5LayerFF, pre-RMSNorm):
        y = x + Wo ( act(W0 n) * (W1 n) ),   n = RMSNorm(x)
    Kernels: the W0 GEMM's epilogue evaluates act and act' (stored, so the backward never recomputes it), the gate
    product and its two gradients are one elementwise kernel each, the residual rides in the Wo GEMM's epilogue; in
    the backward the second data-gradient GEMM adds the first one's result in its epilogue."""

    @staticmethod
    def forward(ctx, x, cfg, w0, w1, wo, ln_w):
        eps = cfg["eps"]
        act, _ = _act_codes(cfg.get("act", "gelu_new"))
        shp = x.shape
        H = shp[-1]
        x2 = x.reshape(-1, H)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        n, _, mean, rstd = K.layernorm_fwd(x2, ln_w.detach(), None, eps, rms_only=True)
        need_bwd = not K.FP32_MODE       # the fp32 verification mode is inference-only
        if not need_bwd:
            g = K.linear_fwd(n, w16(w0), None, act=ACT_GELU if act == ACT_GELU_G else act)
            gp = None
        elif act == ACT_RELU:
            g, pre0 = K.linear_fwd(n, w16(w0), None, act=ACT_RELU, want_pre=True)
            gp = K.dact(torch.ones_like(pre0), pre0, ACT_RELU)          # relu'(pre)
        else:
            g, gp = K.linear_fwd(n, w16(w0), None, act=act, want_pre=True)   # GELU_G: aux output = act'(pre)
        u = K.linear_fwd(n, w16(w1), None)
        h = K.mul_bf16(g, u)
        drop_act = _drop_site(cfg.get("p_act"), x.device, "elementwise", tuple(shp[:-1]) + (w0.shape[0],))
        drop_h = _drop_site(cfg.get("p_hidden"), x.device, "elementwise", shp)
        if drop_act is not None:
            h = K.dropout(h, *drop_act)
        if drop_h is None:
            y = K.linear_fwd(h, w16(wo), None, residual=x2)
        else:
            y = K.dropout(K.linear_fwd(h, w16(wo), None), *drop_h, residual=x2)
        m = K.mul_bf16(u, gp) if need_bwd else None                      # d h / d pre0 (before dropout)
        ctx.save_for_backward(x2, n, mean, rstd, g, m, h, ln_w)
        ctx.shp, ctx.drop_act, ctx.drop_h, ctx.wrefs = shp, drop_act, drop_h, (w0, w1, wo)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, n, mean, rstd, g, m, h, ln_w = ctx.saved_tensors
        w0, w1, wo = ctx.wrefs
        ds_res = dy.reshape(-1, ctx.shp[-1]).contiguous()
        ds = K.dropout(ds_res, *ctx.drop_h) if ctx.drop_h is not None else ds_res
        dwo = K.linear_wgrad(ds, h) if _need(ctx, 4) else None
        dh = K.linear_dgrad(ds, w16(wo))
        if ctx.drop_act is not None:
            dh = K.dropout(dh, *ctx.drop_act)
        dpre0 = K.mul_bf16(dh, m)
        du = K.mul_bf16(dh, g)
        dw0 = K.linear_wgrad(dpre0, n) if _need(ctx, 2) else None
        dw1 = K.linear_wgrad(du, n) if _need(ctx, 3) else None
        dn = K.linear_dgrad(du, w16(w1), residual=K.linear_dgrad(dpre0, w16(w0)))
        dx, dlnw, _ = K.layernorm_bwd(dn, x2, ln_w.detach(), mean, rstd, dres=ds_res, rms_only=True, want_dbeta=False)
        return dx.view(ctx.shp), None, dw0, dw1, dwo, dlnw


@torch.no_grad()
def decode_gated_ffn_step(x2, cfg, w0, w1, wo, ln_w):
    act, _ = _act_codes(cfg.get("act", "gelu_new"))
    if act == ACT_GELU_G:
        act = ACT_GELU
    a_in = _ln_maybe(x2, ln_w, None, cfg["eps"], True)
    h = K.mul_bf16(K.linear_fwd(a_in, w16(w0), None, act=act), K.linear_fwd(a_in, w16(w1), None))
    return K.linear_fwd(h, w16(wo), None, residual=x2)


class ConvK2Fn(torch.autograd.Function):
    """Conv1d(C -> N, kernel 2, stride 2, bias) on channels-last input through the copy-free frame-group view
    (K.conv_ks_*): the non-final down_scale length adapters (ref:speechmix/hf_model.py:253-266, 426-427)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        wp = conv_packed16(weight)
        y = K.conv_ks_fwd(x, wp, 2, bias=None if bias is None else bias.detach())
        ctx.save_for_backward(x, wp)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp = ctx.saved_tensors
        dy = dy.contiguous()
        dx = K.conv_ks_dgrad(dy, wp, 2, x.shape[1]) if _need(ctx, 0) else None
        dw = K.unpack_conv_wgrad(K.conv_ks_wgrad(dy, x, 2), x.shape[2], 2) if _need(ctx, 1) else None
        db = K.colsum(dy.view(-1, dy.shape[-1])) if _need(ctx, 2) else None
        return dx, dw, db


class BridgeProjFn(torch.autograd.Function):
    """The last down_scale length adapter and the encoder-to-decoder projector as ONE kernel launch.

    ref:speechmix/hf_model.py:426-430 runs ``enc_to_dec_proj(length_adapters(x))`` with NO non-linearity in between
    (ref :253-272: bare Conv1d(k 2, s 2) layers, then nn.Linear), so the pair is a single linear map of a frame pair:

        y[t] = Wp (W[:, :, 0] x[2t] + W[:, :, 1] x[2t+1] + b) + bp  =  W_eff [x[2t] ; x[2t+1]] + b_eff
        W_eff = Wp . pack(W)   [D, 2C]          b_eff = bp + Wp b

    W_eff is a weight-sized product (one small GEMM per step, D.C.2C MACs); the activations then go through ONE implicit
    GEMM over the copy-free frame-pair view instead of conv GEMM -> HBM round trip of the [B, T/2, C] intermediate ->
    projector GEMM: a third fewer MACs on the activations and no intermediate tensor (forward) / no intermediate
    gradient (backward).  Backward: dx, G = dW_eff (fp32) and db_eff come from the usual NN / TN / column-sum kernels;
    the two weight gradients follow from G with two more weight-sized GEMMs (dWp = G pack(W)^T, dpack(W) = Wp^T G).
    The caller (model.bridge) takes this path when it saves work: rows_out >= 4 C (ops.bridge_fusion_pays)."""

    @staticmethod
    def forward(ctx, x, cw, cb, pw, pb):
        x = x.contiguous()
        P = conv_packed16(cw)                       # [C, 2C] bf16, tap-major
        pw16 = w16(pw)                              # [D, C]
        w_eff = K.linear_dgrad(pw16, P)             # [D, 2C] = pw16 @ P   (NN GEMM, fp32 accumulation, bf16 out)
        b_eff = torch.addmv(pb.detach(), pw.detach(), cb.detach())
        y = K.conv_ks_fwd(x, w_eff, 2, bias=b_eff)
        ctx.save_for_backward(x, w_eff, P, pw16, cb.detach(), pw.detach())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_eff, P, pw16, cb, pw = ctx.saved_tensors
        dy = dy.contiguous()
        C, D = x.shape[2], dy.shape[2]
        dx = K.conv_ks_dgrad(dy, w_eff, 2, x.shape[1]) if _need(ctx, 0) else None
        dcw = dcb = dpw = dpb = None
        if any(_need(ctx, i) for i in (1, 2, 3, 4)):
            G16 = K.to_bf16(K.conv_ks_wgrad(dy, x, 2))            # dW_eff [D, 2C]
            db_eff = K.colsum(dy.view(-1, D))
            if _need(ctx, 3):
                dpw = K.linear_fwd(G16, P, out_f32=True)          # G P^T  [D, C]
                dpw.addr_(db_eff, cb)                             # b_eff = bp + Wp b
            if _need(ctx, 1):
                dcw = K.unpack_conv_wgrad(K.linear_wgrad(pw16, G16), C, 2)   # Wp^T G  [C, 2C] -> [C, C, 2]
            if _need(ctx, 2):
                dcb = torch.mv(pw.t(), db_eff)
            if _need(ctx, 4):
                dpb = db_eff
        return dx, dcw, dcb, dpw, dpb


def bridge_fusion_pays(rows_out, channels):
    """composing W_eff costs D.C.2C MACs forward (3x that with the two backward products); folding the projector into
    the conv saves rows_out.D.C MACs forward (and as much again in each backward product)."""
    if K.FP32_MODE or BRIDGE_FUSION == "never":
        return False
    return BRIDGE_FUSION == "always" or rows_out >= 4 * channels


BRIDGE_FUSION = "auto"     # "auto" | "always" | "never"  (the GPU tests pin both paths on the small fixtures)


class WeightNormFn(torch.autograd.Function):
    """w = g * v / ||v||  with one norm per kernel tap (torch weight_norm(dim=2) on the positional conv,
    hf:...wav2vec2.py:341-355); g = parametrizations.weight.original0 [1,1,k], v = original1 [H, H/groups, k]."""

    @staticmethod
    def forward(ctx, g, v):
        w, sq = K.weightnorm_fwd(v, g)
        ctx.save_for_backward(g, v, sq)
        return w

    @staticmethod
    def backward(ctx, dw):
        g, v, sq = ctx.saved_tensors
        dv, dg = K.weightnorm_bwd(v, g, sq, dw.float())
        return dg.view(g.shape), dv


class PosConvFn(torch.autograd.Function):
    """y = x + GELU(grouped_conv(x) + b)[drop last frame]   hf:...wav2vec2.py:326-379, 690-693."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, key_params):
        x = x.contiguous()
        ksize = weight.shape[2]
        # `weight` is usually recomputed every step from its weight-norm factors; the packed bf16
        # copies are cached on the leaf parameters it derives from.
        if K.FP32_MODE:
            return K.posconv_fwd(x, weight.detach(), bias.detach(), groups, ksize, add_input=True)[0]
        wf, wd = CACHE.get(tuple(key_params), "posconv%d" % groups, lambda *_: K.posconv_pack(weight, groups))
        y, pre = K.posconv_fwd(x, wf, bias.detach(), groups, ksize, add_input=True)
        ctx.save_for_backward(x, pre, wd)
        ctx.groups, ctx.ksize = groups, ksize
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, wd = ctx.saved_tensors
        dy = dy.contiguous()
        dpre = K.dact(dy, pre, ACT_GELU)
        dx = K.posconv_dgrad(dpre, wd, ctx.groups, ctx.ksize, residual=dy)
        dw = K.posconv_wgrad(dpre, x, ctx.groups, ctx.ksize) if _need(ctx, 1) else None
        db = K.colsum(dpre.view(-1, dpre.shape[-1])) if _need(ctx, 2) else None
        return dx, dw, db, None, None


# ---------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """out = tok[ids]*scale (ids given) + x_in (given) + pos[t + offset]   (bf16)
    hf:...bart.py:74-111, 508-530, 620-640."""

    @staticmethod
    def forward(ctx, ids, x_in, tok, pos, scale, pos_offset, t_start):
        if ids is not None:
            B, T = ids.shape
            dev = ids.device
        else:
            B, T = x_in.shape[:2]
            dev = x_in.device
        D = pos.shape[1] if pos is not None else (tok.shape[1] if tok is not None else x_in.shape[2])
        out = K.embed_fwd(ids, None if tok is None else tok.detach(), None if pos is None else pos.detach(),
                          None if x_in is None else x_in.contiguous(), B, T, D, scale, pos_offset, t_start, device=dev)
        ctx.save_for_backward(ids)
        ctx.meta = (scale, pos_offset + t_start, None if tok is None else tok.shape, None if pos is None else pos.shape,
                    x_in is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        scale, off, tok_shape, pos_shape, has_x = ctx.meta
        dout = dout.contiguous()
        d_tok = torch.zeros(tok_shape, device=dout.device, dtype=torch.float32) if (tok_shape and _need(ctx, 2)) else None
        d_pos = torch.zeros(pos_shape, device=dout.device, dtype=torch.float32) if (pos_shape and _need(ctx, 3)) else None
        if d_tok is not None or d_pos is not None:
            K.embed_bwd(ids, dout, d_tok, d_pos, scale, off)
        return None, (dout if has_x else None), d_tok, d_pos, None, None, None


# ---------------------------------------------------------------------------
LM_CHUNK = 8192


class LMHeadCEFn(torch.autograd.Function):
    """(loss, argmax ids) = CE(h E^T * s + b, labels) with the [rows, vocab] logits never
    materialised.  hf:...bart.py:940-947 ; ref:speechmix/hf_model.py:446.
    Returns loss (mean over labels != -100; 0-dim fp32) and argmax ids [rows] int64."""

    @staticmethod
    def forward(ctx, h, emb, bias, labels, logit_scale):
        shp = h.shape
        h2 = h.reshape(-1, shp[-1]).contiguous()
        e16 = w16(emb)
        lab = labels.reshape(-1).contiguous()
        b = None if bias is None else bias.detach().reshape(-1).float().contiguous()
        lse, argmax, row_loss, acc = K.lmhead_ce_fwd(h2, e16, b, lab, logit_scale)
        loss = acc[0] / acc[1]
        ctx.save_for_backward(h2, e16, lab, lse, acc)
        ctx.bias = b
        ctx.scale = logit_scale
        ctx.shp = shp
        ctx.emb_shape = emb.shape
        ctx.mark_non_differentiable(argmax)
        return loss, argmax.view(shp[:-1])

    @staticmethod
    def backward(ctx, dloss, _unused):
        h2, e16, lab, lse, acc = ctx.saved_tensors
        M, D = h2.shape
        V = e16.shape[0]
        coef = ((lab != -100).float() * (dloss.float() / acc[1])).contiguous()
        need_h, need_e = _need(ctx, 0), _need(ctx, 1)
        dh = torch.zeros(M, D, device=h2.device, dtype=torch.float32) if need_h else None
        dE = torch.empty(V, D, device=h2.device, dtype=torch.float32) if need_e else None
        buf = torch.empty(M, LM_CHUNK, device=h2.device, dtype=BF16)
        for v0 in range(0, V, LM_CHUNK):
            vn = min(LM_CHUNK, V - v0)
            K.lmhead_dlogits(h2, e16, ctx.bias, lab, lse, coef, buf, v0, vn, ctx.scale)
            if need_h:
                K.gemm_nn_acc_f32(buf, vn, e16[v0:v0 + vn], dh, alpha=ctx.scale, accumulate=True)
            if need_e:
                K.gemm_tn_into(buf, vn, h2, dE[v0:v0 + vn], alpha=ctx.scale)
        dh16 = K.to_bf16(dh).view(ctx.shp) if need_h else None
        return dh16, dE, None, None, None


# ---------------------------------------------------------------------------
class WeightedSumFn(torch.autograd.Function):
    """out = sum_l softmax(w)[l] * x_l   (ref:speechmix/hf_model.py:411-423); takes normalised weights."""

    @staticmethod
    def forward(ctx, norm_w, *xs):
        xs = [x.contiguous() for x in xs]
        out = K.weighted_sum_fwd(xs, norm_w.detach().float().contiguous())
        ctx.save_for_backward(norm_w, *xs)
        return out

    @staticmethod
    def backward(ctx, dout):
        norm_w, *xs = ctx.saved_tensors
        dout = dout.contiguous()
        dw = K.weighted_sum_bwd_w(xs, dout)
        wl = norm_w.detach().float()
        dxs = [K.weighted_sum_fwd([dout], wl[l:l + 1].contiguous()) for l in range(len(xs))]
        return (dw, *dxs)


# ---------------------------------------------------------------------------
_BUCKET_TABLES = {}


def t5_bucket_table(tq, tk, bidirectional, num_buckets, max_distance, device, q_offset=0):
    """int32 table: index (j - (i + q_offset)) + (tq + q_offset - 1) -> relative-attention bucket.  Built on the
    host with the reference's own arithmetic (hf:models/t5/modeling_t5.py:188-233, fp32 log on CPU) so that
    bucket boundaries agree bit for bit with the CPU reference."""
    key = (tq, tk, bool(bidirectional), num_buckets, max_distance, str(device), q_offset)
    tab = _BUCKET_TABLES.get(key)
    if tab is not None:
        return tab
    rel = torch.arange(-(tq + q_offset - 1), tk, dtype=torch.long)
    nb = num_buckets
    buckets = torch.zeros_like(rel)
    if bidirectional:
        nb //= 2
        buckets += (rel > 0).to(torch.long) * nb
        rel = torch.abs(rel)
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    buckets += torch.where(is_small, rel, large)
    tab = buckets.to(torch.int32).to(device)
    _BUCKET_TABLES[key] = tab
    return tab


class RelPosBiasFn(torch.autograd.Function):
    """bias[h, i, j] = weight[bucket(j - i), h]   (hf:models/t5/modeling_t5.py:235-247, compute_bias)."""

    @staticmethod
    def forward(ctx, weight, table, tq, tk, q_offset):
        nb, heads = weight.shape
        ctx.save_for_backward(table)
        ctx.meta = (heads, tq, tk, q_offset, nb)
        return K.relpos_bias_fwd(weight.detach().float().contiguous(), table, heads, tq, tk, q_offset)

    @staticmethod
    def backward(ctx, dbias):
        (table,) = ctx.saved_tensors
        heads, tq, tk, q_offset, nb = ctx.meta
        return K.relpos_bias_bwd(dbias.contiguous(), table, heads, tq, tk, nb, q_offset), None, None, None, None


# ---------------------------------------------------------------------------
KL_CHUNK = 4096


class SelfDistillHeadFn(torch.autograd.Function):
    """SpeechMixSelf output losses in one vocabulary sweep (ref:speechmix/hf_model.py:551-581):
       ce  = CrossEntropy(student logits, labels)                       (ignore_index -100, mean)
       kld = KLDiv(log_softmax(student logits), softmax(teacher logits), reduction="batchmean")
    plus the student's argmax ids.  Student / teacher logits (h E^T * s + b) only ever exist as fp32
    [rows, 4096] chunks that stay L2-resident between the GEMM and the reduction kernels."""

    @staticmethod
    def forward(ctx, h_s, h_t, emb, bias, labels, logit_scale, batch):
        shp = h_s.shape
        hs2 = h_s.reshape(-1, shp[-1]).contiguous()
        ht2 = h_t.detach().reshape(-1, shp[-1]).contiguous()
        e16 = w16(emb)
        lab = labels.reshape(-1).contiguous()
        b = None if bias is None else bias.detach().reshape(-1).float().contiguous()
        M, V = hs2.shape[0], e16.shape[0]
        lse_s, argmax, _, acc = K.lmhead_ce_fwd(hs2, e16, b, lab, logit_scale)
        lse_t, _, _, _ = K.lmhead_ce_fwd(ht2, e16, b, torch.full_like(lab, -100), logit_scale)
        cross = torch.zeros(M, device=hs2.device, dtype=torch.float32)
        bs = torch.empty(M, KL_CHUNK, device=hs2.device, dtype=torch.float32)
        bt = torch.empty(M, KL_CHUNK, device=hs2.device, dtype=torch.float32)
        for v0 in range(0, V, KL_CHUNK):
            vn = min(KL_CHUNK, V - v0)
            bb = None if b is None else b[v0:v0 + vn]
            K.logits_chunk_f32(hs2, e16[v0:v0 + vn], bb, logit_scale, bs[:, :vn])
            K.logits_chunk_f32(ht2, e16[v0:v0 + vn], bb, logit_scale, bt[:, :vn])
            K.kl_chunk_fwd(bs, bt, vn, lse_t, cross)
        kld = K.kl_finalize(cross, lse_s, lse_t, 1.0 / batch)[0]
        ce = acc[0] / acc[1]
        ctx.save_for_backward(hs2, ht2, e16, lab, lse_s, lse_t, acc)
        ctx.bias, ctx.scale, ctx.shp, ctx.batch = b, logit_scale, shp, batch
        ctx.mark_non_differentiable(argmax)
        return ce, kld, argmax.view(shp[:-1])

    @staticmethod
    def backward(ctx, d_ce, d_kl, _unused):
        hs2, ht2, e16, lab, lse_s, lse_t, acc = ctx.saved_tensors
        M, D = hs2.shape
        V = e16.shape[0]
        b = ctx.bias
        coef_ce = ((lab != -100).float() * (d_ce.float() / acc[1])).contiguous()
        coef_kl = (d_kl.float() / ctx.batch).reshape(1).contiguous()
        need_h, need_e = _need(ctx, 0), _need(ctx, 2)
        dh = torch.zeros(M, D, device=hs2.device, dtype=torch.float32) if need_h else None
        dE = torch.empty(V, D, device=hs2.device, dtype=torch.float32) if need_e else None
        bs = torch.empty(M, KL_CHUNK, device=hs2.device, dtype=torch.float32)
        bt = torch.empty(M, KL_CHUNK, device=hs2.device, dtype=torch.float32)
        buf = torch.empty(M, KL_CHUNK, device=hs2.device, dtype=BF16)
        for v0 in range(0, V, KL_CHUNK):
            vn = min(KL_CHUNK, V - v0)
            bb = None if b is None else b[v0:v0 + vn]
            K.logits_chunk_f32(hs2, e16[v0:v0 + vn], bb, ctx.scale, bs[:, :vn])
            K.logits_chunk_f32(ht2, e16[v0:v0 + vn], bb, ctx.scale, bt[:, :vn])
            K.kl_chunk_bwd(bs, bt, vn, v0, lab, lse_s, lse_t, coef_ce, coef_kl, buf)
            if need_h:
                K.gemm_nn_acc_f32(buf, vn, e16[v0:v0 + vn], dh, alpha=ctx.scale, accumulate=True)
            if need_e:
                K.gemm_tn_into(buf, vn, hs2, dE[v0:v0 + vn], alpha=ctx.scale)
        dh16 = K.to_bf16(dh).view(ctx.shp) if need_h else None
        return dh16, None, dE, None, None, None, None


class GramLogitFn(torch.autograd.Function):
    """SpeechMixGAN discriminator logit (ref:speechmix/hf_model.py:637-686):
    ``Linear(D*D, 1)(flatten(bmm(X.view(B, D, T), X.view(B, T, D))))`` without the D x D Gram matrix: with
    W = weight.view(D, D) the logit is <W, G_b> = sum_{t,i} Xflat_b[i T + t] Z_b[t, i] for Z = X W^T -- one GEMM on the
    tensor cores plus a streaming contraction (``smx_gram_dot_*``); the D*D features per sample (590k for bart-base)
    never reach HBM.  x [B, T, D] activations, weight [1, D*D] fp32 parameter, bias [1] -> logits [B] fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if K.FP32_MODE:
            raise NotImplementedError("the GAN discriminator has no fp32 verification path (training-only term)")
        B, T, D = x.shape
        xc = x.contiguous()
        wb = w16(weight).view(D, D)
        z = K.linear_fwd(xc.view(B * T, D), wb, None).view(B, T, D)
        out = K.gram_dot_fwd(xc, z)
        ctx.save_for_backward(xc, z, wb)
        return out + bias.detach().float()

    @staticmethod
    def backward(ctx, g):
        xc, z, wb = ctx.saved_tensors
        B, T, D = xc.shape
        g = g.float().contiguous()
        dx, dz = K.gram_dot_bwd(xc, z, g)
        dz2 = dz.view(B * T, D)
        # both gradient paths of X: the reinterpreted factor (dx) and the GEMM operand (dZ W), summed in the GEMM epilogue
        dx = K.linear_dgrad(dz2, wb, residual=dx.view(B * T, D)).view(B, T, D) if _need(ctx, 0) else None
        dw = K.linear_wgrad(dz2, xc.view(B * T, D)).reshape(1, D * D) if _need(ctx, 1) else None
        db = g.sum().reshape(1) if _need(ctx, 2) else None
        return dx, dw, db


class SelfMSEFn(torch.autograd.Function):
    """mse(softmax(T . view(S, [D, Ts]) / sqrt(D)) . S, T)   (ref:speechmix/hf_model.py:561-570);
    T = teacher text-encoder states (no gradient), S = text-encoder states of the speech path."""

    @staticmethod
    def forward(ctx, text_h, speech_h):
        t = text_h.detach().contiguous()
        s = speech_h.contiguous()
        loss, attn, diff = K.self_mse_fwd(t, s)
        ctx.save_for_backward(t, s, attn, diff)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        t, s, attn, diff = ctx.saved_tensors
        gscale = (g.float() * (2.0 / diff.numel())).reshape(1).contiguous()
        return None, K.self_mse_bwd(t, s, attn, diff, gscale)


# ---------------------------------------------------------------------------
# Incremental decoding (no autograd): one new token per call, self-attention keys / values appended to a
# preallocated cache, cross-attention keys / values projected once per utterance.
# hf:models/bart/modeling_bart.py:143-258 (past_key_values branch), ref:speechmix/hf_model.py:314-338.
# ---------------------------------------------------------------------------
def _ln_maybe(x2, w, b, eps, rms):
    return K.layernorm_fwd(x2, w.detach(), None if b is None else b.detach(), eps, rms_only=rms)[0]


@torch.no_grad()
def decode_self_attn_step(x2, cfg, q_w, q_b, k_w, k_b, v_w, v_b, o_w, o_b, ln_w, ln_b, cache, t, pos_bias=None):
    """x2 [B, H] bf16 (the new token), cache [B, Tmax, 2*Hi] bf16 (k | v).  Returns the sub-block output [B, H]."""
    heads, pre_ln, eps, rms = cfg["heads"], cfg["pre_ln"], cfg["eps"], bool(cfg.get("rms", False))
    scale = cfg.get("scale", 1.0 / math.sqrt(64))
    B, H = x2.shape
    Hi = q_w.shape[0]
    a_in = _ln_maybe(x2, ln_w, ln_b, eps, rms) if pre_ln else x2
    q = K.linear_fwd(a_in, w16(q_w), None if q_b is None else q_b.detach())
    K.linear_fwd(a_in, cat16((k_w, v_w)), cat32((k_b, v_b)) if k_b is not None else None, out=cache[:, t, :])
    kv = cache[:, :t + 1]
    o, _ = K.attn_fwd(q.view(B, 1, Hi), kv[..., :Hi], kv[..., Hi:], heads, causal=False, scale=scale, bias=pos_bias)
    s = K.linear_fwd(o.view(B, Hi), w16(o_w), None if o_b is None else o_b.detach(), residual=x2)
    return s if pre_ln else _ln_maybe(s, ln_w, ln_b, eps, rms)


@torch.no_grad()
def cross_kv(src, k_w, k_b, v_w, v_b):
    """keys | values of the encoder states for one decoder layer: [B, Ts, 2*Hi] bf16 (computed once)."""
    Bs, Ts, Hs = src.shape
    kv = K.linear_fwd(src.reshape(Bs * Ts, Hs).contiguous(), cat16((k_w, v_w)),
                      cat32((k_b, v_b)) if k_b is not None else None)
    return kv.view(Bs, Ts, -1)


@torch.no_grad()
def decode_cross_attn_step(x2, cfg, q_w, q_b, o_w, o_b, ln_w, ln_b, kv):
    heads, pre_ln, eps, rms = cfg["heads"], cfg["pre_ln"], cfg["eps"], bool(cfg.get("rms", False))
    scale = cfg.get("scale", 1.0 / math.sqrt(64))
    B, H = x2.shape
    Hi = q_w.shape[0]
    a_in = _ln_maybe(x2, ln_w, ln_b, eps, rms) if pre_ln else x2
    q = K.linear_fwd(a_in, w16(q_w), None if q_b is None else q_b.detach())
    o, _ = K.attn_fwd(q.view(B, 1, Hi), kv[..., :Hi], kv[..., Hi:], heads, causal=False, scale=scale)
    s = K.linear_fwd(o.view(B, Hi), w16(o_w), None if o_b is None else o_b.detach(), residual=x2)
    return s if pre_ln else _ln_maybe(s, ln_w, ln_b, eps, rms)


@torch.no_grad()
def decode_ffn_step(x2, cfg, w1, b1, w2, b2, ln_w, ln_b):
    pre_ln, eps, rms = cfg["pre_ln"], cfg["eps"], bool(cfg.get("rms", False))
    act, _ = _act_codes(cfg.get("act", "gelu"))
    a_in = _ln_maybe(x2, ln_w, ln_b, eps, rms) if pre_ln else x2
    h = K.linear_fwd(a_in, w16(w1), None if b1 is None else b1.detach(), act=act)
    s = K.linear_fwd(h, w16(w2), None if b2 is None else b2.detach(), residual=x2)
    return s if pre_ln else _ln_maybe(s, ln_w, ln_b, eps, rms)


@torch.no_grad()
def lm_head_argmax(h2, emb, bias, logit_scale):
    """argmax_v (h E^T * s + b) per row, lowest index wins ties (torch.argmax); no logits in HBM."""
    b = None if bias is None else bias.detach().reshape(-1).float().contiguous()
    lab = torch.full((h2.shape[0],), -100, device=h2.device, dtype=torch.long)
    return K.lmhead_ce_fwd(h2.contiguous(), w16(emb), b, lab, logit_scale)[1]
