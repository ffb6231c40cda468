"""Tensor-level wrappers over the C ABI (no autograd here; see ``ops.py``).

Every function takes CUDA tensors, enqueues one or more kernels of
``libspeechmix_sm100.so`` on the current torch stream and returns the output
tensors.  Activations are bf16, channels-last.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (ACT_DGELU, ACT_DRELU, ACT_GELU, ACT_GELU_G, ACT_MULAUX, ACT_NONE, ACT_RELU, GEMM_NN, GEMM_NT, GEMM_TN, OUT_BF16,
                   OUT_F32, SmxAttn, SmxGemm, SmxView3)

BF16 = torch.bfloat16

# fp32 verification mode (ops.fp32_verification()): activations fp32, every call below is routed to fp32path.py
FP32_MODE = False


def act_dtype():
    return torch.float32 if FP32_MODE else BF16

# bookkeeping for bench.py: number of kernels of this library launched so far, and an optional
# list collecting (tag, flops, start_event, end_event) around tagged GEMM launches.
LAUNCHES = _lib.LAUNCHES
TIMING = None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, offset_elems=0):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _view(t, inner, rows, batches, row_stride, batch_stride, offset=0):
    return SmxView3(_ptr(t, offset), inner, rows, batches, row_stride, batch_stride)


def alloc_act(batch, t, c, device, dtype=None, slack=None):
    """[batch, t, c] activation with ``slack`` (default c) zeroed elements after
    the end, so frame-pair TMA views of an odd-length signal stay in bounds."""
    slack = c if slack is None else slack
    dtype = act_dtype() if dtype is None else dtype
    n = batch * t * c
    flat = torch.empty(n + slack, device=device, dtype=dtype)
    flat[n:].zero_()
    return flat[:n].view(batch, t, c)


class _ZeroArena:
    """fp32 zero-initialised scratch for accumulate-by-atomics outputs (split-K weight gradients, bias /
    affine gradients): one memset per 128 MiB chunk instead of one fill launch per tensor (~360 per step).
    Views keep their chunk alive; a chunk is dropped here once it is exhausted."""
    CHUNK = 32 * 1024 * 1024  # fp32 elements

    def __init__(self):
        self.buf, self.off = {}, {}

    def reset(self):
        """Drop the current chunks (around a CUDA-graph capture: captured chunks belong to the graph's pool)."""
        self.buf, self.off = {}, {}

    def zeros(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        if n * 2 > self.CHUNK:
            return torch.zeros(shape, device=device, dtype=torch.float32)
        dev = torch.device(device)
        buf = self.buf.get(dev)
        off = self.off.get(dev, 0)
        n_al = (n + 63) // 64 * 64          # keep every carve-out 256-byte aligned
        if buf is None or off + n_al > self.CHUNK:
            buf = torch.zeros(self.CHUNK, device=dev, dtype=torch.float32)
            self.buf[dev], off = buf, 0
        self.off[dev] = off + n_al
        return buf[off:off + n].view(shape)


_ARENA = _ZeroArena()


def zeros_f32(*shape, device):
    return _ARENA.zeros(shape, device)


def _run_gemm(g, tag=None, flops=0.0):
    lib = _lib.load()
    LAUNCHES[0] += 1
    if TIMING is not None and tag is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.smx_gemm(ctypes.byref(g), _stream()), "smx_gemm")
        e1.record()
        TIMING.append((tag, flops, e0, e1))
        return
    _lib.check(lib.smx_gemm(ctypes.byref(g), _stream()), "smx_gemm")


def _set_seg(g, nseg, seg_len, a_row=None, a_col=None, b_row=None, b_col=None):
    g.nseg, g.seg_len = nseg, seg_len
    for i in range(nseg):
        g.a_row_off[i] = a_row[i] if a_row else 0
        g.a_col_off[i] = a_col[i] if a_col else 0
        g.b_row_off[i] = b_row[i] if b_row else 0
        g.b_col_off[i] = b_col[i] if b_col else 0


def _epilogue(g, c, c_row_stride, c_batch_stride, bias, act, residual, res_strides, aux_out, aux_in, alpha,
              c_offset=0):
    g.c = _ptr(c, c_offset)
    g.c_row_stride, g.c_batch_stride = c_row_stride, c_batch_stride
    g.act = act
    g.alpha = alpha
    g.bias = _ptr(bias)
    if residual is not None:
        g.residual = _ptr(residual)
        g.res_row_stride, g.res_batch_stride = res_strides
    g.aux_out = _ptr(aux_out, c_offset) if aux_out is not None else None
    g.aux_in = _ptr(aux_in, c_offset) if aux_in is not None else None


# ---------------------------------------------------------------------------
# linear layers  (x: [M, K] bf16 row-major, w: [N, K] bf16 = torch Linear layout)
# ---------------------------------------------------------------------------
def linear_fwd(x, w, bias=None, act=ACT_NONE, residual=None, want_pre=False, out_f32=False, alpha=1.0, out=None):
    if FP32_MODE:
        from . import fp32path
        return fp32path.linear_fwd(x, w, bias, act, residual, want_pre, out_f32, alpha, out)
    assert x.dtype == BF16 and w.dtype == BF16 and x.is_contiguous() and w.is_contiguous()
    if act == ACT_GELU_G and not want_pre:
        act = ACT_GELU      # no auxiliary output requested: plain GELU epilogue
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and K % 8 == 0
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=torch.float32 if out_f32 else BF16)
    pre = torch.empty(M, N, device=x.device, dtype=BF16) if want_pre else None
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NT, OUT_F32 if out_f32 else OUT_BF16
    g.a = _view(x, K, M, 1, K, M * K)
    g.b = _view(w, K, N, 1, K, N * K)
    g.m, g.n, g.k, g.batches = M, N, K, 1
    _set_seg(g, 1, K)
    _epilogue(g, y, y.stride(0), M * y.stride(0), bias, act, residual,
              (residual.stride(0), 0) if residual is not None else None, pre, None, alpha)
    _run_gemm(g, "nt_%d_%d_%d" % (M, N, K), 2.0 * M * N * K)
    return (y, pre) if want_pre else y


def linear_dgrad(dy, w, act=ACT_NONE, aux_in=None, residual=None, alpha=1.0):
    """dx[M, K] = dy[M, N] @ w[N, K]   (optionally * act'(aux_in), + residual)."""
    assert dy.dtype == BF16 and w.dtype == BF16 and dy.is_contiguous() and w.is_contiguous()
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N and N % 8 == 0 and K % 8 == 0
    dx = torch.empty(M, K, device=dy.device, dtype=BF16)
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NN, OUT_BF16
    g.a = _view(dy, N, M, 1, N, M * N)
    g.b = _view(w, K, N, 1, K, N * K)
    g.m, g.n, g.k, g.batches = M, K, N, 1
    _set_seg(g, 1, N)
    _epilogue(g, dx, K, M * K, None, act, residual, (K, 0) if residual is not None else None, None, aux_in, alpha)
    _run_gemm(g)
    return dx


def _pick_split(out_tiles, kblocks, pairs=74):
    """contraction splits for the weight-gradient GEMM: enough 256x256 tiles x splits to occupy the
    74 CTA pairs of a B200."""
    if out_tiles >= pairs:
        return 1
    s = max(1, pairs // out_tiles)
    return int(max(1, min(s, kblocks, 32)))


def linear_wgrad(dy, x, out=None, accumulate=False):
    """dw[N, K] (fp32) = dy[M, N]^T @ x[M, K]."""
    assert dy.dtype == BF16 and x.dtype == BF16 and dy.is_contiguous() and x.is_contiguous()
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and N % 8 == 0 and K % 8 == 0
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_TN, OUT_F32
    g.a = _view(dy, N, M, 1, N, M * N)
    g.b = _view(x, K, M, 1, K, M * K)
    g.m, g.n, g.k, g.batches = N, K, M, 1
    _set_seg(g, 1, K)
    tiles = math.ceil(N / 256) * math.ceil(K / 256)
    g.split_k = _pick_split(tiles, math.ceil(M / 64))
    g.accumulate = 1 if accumulate else 0
    if out is None:
        dw = zeros_f32(N, K, device=dy.device) if g.split_k > 1 else torch.empty(N, K, device=dy.device, dtype=torch.float32)
    else:
        dw = out
    g.c = _ptr(dw)
    g.c_row_stride, g.c_batch_stride = K, N * K
    g.alpha = 1.0
    _run_gemm(g)
    return dw


# ---------------------------------------------------------------------------
# stride-2 Conv1d (k = 2 or 3) on channels-last activations, implicit GEMM
#   x: [B, T_in, C] from alloc_act (slack!), w_packed: [N, k*C] (tap-major)
# ---------------------------------------------------------------------------
def conv_out_len(t_in, k, s):
    return (t_in - k) // s + 1


def _pair_view(x):
    B, T, C = x.shape
    return _view(x, 2 * C, (T + 1) // 2, B, 2 * C, T * C)


def conv_s2_fwd(x, w_packed, k, bias=None, act=ACT_NONE, want_pre=False):
    if FP32_MODE:
        from . import fp32path
        return fp32path.conv_s2_fwd(x, w_packed, k, bias, act, want_pre)
    B, T_in, C = x.shape
    N = w_packed.shape[0]
    assert w_packed.shape[1] == k * C and C % 64 == 0 and x.is_contiguous()
    if act == ACT_GELU_G and not want_pre:
        act = ACT_GELU
    T_out = conv_out_len(T_in, k, 2)
    y = alloc_act(B, T_out, N, x.device)
    pre = alloc_act(B, T_out, N, x.device) if want_pre else None
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NT, OUT_BF16
    g.a = _pair_view(x)
    g.b = _view(w_packed, k * C, N, 1, k * C, N * k * C)
    g.m, g.n, g.k, g.batches = T_out, N, k * C, B
    _set_seg(g, k, C, a_row=[t >> 1 for t in range(k)], a_col=[(t & 1) * C for t in range(k)],
             b_col=[t * C for t in range(k)])
    _epilogue(g, y, N, T_out * N, bias, act, None, None, pre, None, 1.0)
    _run_gemm(g)
    return (y, pre) if want_pre else y


def conv_s2_dgrad(dy, w_packed, k, t_in, act=ACT_NONE, aux_in=None):
    """dx[B, t_in, C] from dy[B, T_out, N]; optional fused * act'(aux_in) where
    aux_in has dx's shape (the producer layer's pre-activation)."""
    B, T_out, N = dy.shape
    C = w_packed.shape[1] // k
    assert N % 64 == 0 and dy.is_contiguous()
    dx = alloc_act(B, t_in, C, dy.device)
    for parity in (0, 1):
        rows = (t_in + 1 - parity) // 2
        if rows == 0:
            continue
        taps = [t for t in range(k) if (t & 1) == parity]
        g = SmxGemm()
        g.mode, g.out_dtype = GEMM_NN, OUT_BF16
        g.a = _view(dy, N, T_out, B, N, T_out * N)
        g.b = _view(w_packed, k * C, N, 1, k * C, N * k * C)
        g.m, g.n, g.k, g.batches = rows, C, len(taps) * N, B
        # dx[2j+parity] = sum_{tap} dy[j - tap//2] . W_tap
        _set_seg(g, len(taps), N, a_row=[-(t >> 1) for t in taps], b_col=[t * C for t in taps])
        _epilogue(g, dx, 2 * C, t_in * C, None, act, None, None, None, aux_in, 1.0, c_offset=parity * C)
        _run_gemm(g)
    return dx


def conv_s2_wgrad(dy, x, k):
    """dw_packed[N, k*C] fp32 = sum_{b,t} dy[b,t,:]^T x[b, 2t+tap, :]."""
    B, T_out, N = dy.shape
    C = x.shape[2]
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_TN, OUT_F32
    g.a = _view(dy, N, T_out, B, N, T_out * N)
    g.b = _pair_view(x)
    g.m, g.n, g.k, g.batches = N, k * C, T_out, B
    _set_seg(g, k, C, b_row=[t >> 1 for t in range(k)], b_col=[(t & 1) * C for t in range(k)])
    tiles = math.ceil(N / 256) * math.ceil(k * C / 256)
    g.split_k = _pick_split(tiles, B * math.ceil(T_out / 64))
    dw = zeros_f32(N, k * C, device=dy.device) if g.split_k > 1 else torch.empty(N, k * C, device=dy.device, dtype=torch.float32)
    g.c = _ptr(dw)
    g.c_row_stride, g.c_batch_stride = k * C, N * k * C
    g.alpha = 1.0
    _run_gemm(g)
    return dw


# ---------------------------------------------------------------------------
# non-overlapping Conv1d (kernel = stride = S) on channels-last activations: a frame group [S*C] is contiguous, so the
# convolution is a plain batched GEMM over the [B][T // S][S*C] view of x (batch stride T*C; the T % S tail frames are
# never touched) -- no slack frame, no staging copy.  Used for the down_scale bridge (ops.BridgeProjFn).
# ---------------------------------------------------------------------------
def conv_ks_fwd(x, w, s, bias=None):
    """y[B, T//s, N] = x-groups @ w^T + bias;  w: [N, s*C] bf16 (tap-major)."""
    B, T, C = x.shape
    N, KC = w.shape
    T_out = T // s
    assert KC == s * C and C % 8 == 0 and x.is_contiguous() and w.is_contiguous() and T_out > 0
    y = torch.empty(B, T_out, N, device=x.device, dtype=BF16)
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NT, OUT_BF16
    g.a = _view(x, KC, T_out, B, KC, T * C)
    g.b = _view(w, KC, N, 1, KC, N * KC)
    g.m, g.n, g.k, g.batches = T_out, N, KC, B
    _set_seg(g, 1, KC)
    _epilogue(g, y, N, T_out * N, bias, ACT_NONE, None, None, None, None, 1.0)
    _run_gemm(g, "bridge_nt_%d_%d_%d" % (B * T_out, N, KC), 2.0 * B * T_out * N * KC)
    return y


def conv_ks_dgrad(dy, w, s, t_in):
    """dx[B, t_in, C] = dy[B, T_out, N] @ w[N, s*C], scattered back into frame groups; tail frames get zero."""
    B, T_out, N = dy.shape
    KC = w.shape[1]
    C = KC // s
    assert dy.is_contiguous() and w.shape[0] == N and N % 8 == 0
    dx = torch.empty(B, t_in, C, device=dy.device, dtype=BF16)
    if t_in > T_out * s:
        dx[:, T_out * s:].zero_()
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NN, OUT_BF16
    g.a = _view(dy, N, T_out, B, N, T_out * N)
    g.b = _view(w, KC, N, 1, KC, N * KC)
    g.m, g.n, g.k, g.batches = T_out, KC, N, B
    _set_seg(g, 1, N)
    _epilogue(g, dx, KC, t_in * C, None, ACT_NONE, None, None, None, None, 1.0)
    _run_gemm(g)
    return dx


def conv_ks_wgrad(dy, x, s):
    """dw[N, s*C] fp32 = sum_{b,t} dy[b,t,:]^T x-group[b,t,:]."""
    B, T_out, N = dy.shape
    _, T, C = x.shape
    KC = s * C
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_TN, OUT_F32
    g.a = _view(dy, N, T_out, B, N, T_out * N)
    g.b = _view(x, KC, T_out, B, KC, T * C)
    g.m, g.n, g.k, g.batches = N, KC, T_out, B
    _set_seg(g, 1, KC)
    tiles = math.ceil(N / 256) * math.ceil(KC / 256)
    g.split_k = _pick_split(tiles, B * math.ceil(T_out / 64))
    dw = zeros_f32(N, KC, device=dy.device) if g.split_k > 1 else torch.empty(N, KC, device=dy.device, dtype=torch.float32)
    g.c = _ptr(dw)
    g.c_row_stride, g.c_batch_stride = KC, N * KC
    g.alpha = 1.0
    _run_gemm(g)
    return dw


def pack_conv_weight(w):
    """[out, in, k] fp32 -> [out, k*in] bf16 (tap-major)."""
    if FP32_MODE:
        from . import fp32path
        return fp32path.pack_conv_weight(w)
    out_c, in_c, k = w.shape
    return w.detach().permute(0, 2, 1).reshape(out_c, k * in_c).to(BF16).contiguous()


def unpack_conv_wgrad(dw_packed, in_c, k):
    out_c = dw_packed.shape[0]
    return dw_packed.view(out_c, k, in_c).permute(0, 2, 1).contiguous()


# ---------------------------------------------------------------------------
# attention (head_dim 64); q/k/v: [B, T, heads*64] views with unit inner stride
# ---------------------------------------------------------------------------
def _attn_desc(q, k, v, o, lse, heads, causal, scale, bias, kv_len=None):
    B, Tq, HD = q.shape
    Tk = k.shape[1]
    assert HD == heads * 64 and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    a = SmxAttn()
    a.q, a.k, a.v, a.o, a.lse = _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(lse)
    a.q_row_stride, a.k_row_stride, a.v_row_stride, a.o_row_stride = q.stride(1), k.stride(1), v.stride(1), o.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.o_batch_stride = (
        q.stride(0), k.stride(0), v.stride(0), o.stride(0))
    a.batch, a.heads, a.tq, a.tk, a.causal = B, heads, Tq, Tk, 1 if causal else 0
    a.scale = scale
    a.bias = _ptr(bias)
    if kv_len is not None:
        assert kv_len.dtype == torch.int32 and kv_len.is_cuda and kv_len.numel() == B and not causal
    a.kv_len = _ptr(kv_len)
    return a


def _set_dropout(a, dropout):
    """dropout = (state tensor int64 [2] on the device, call index, p) or None"""
    if dropout is not None and dropout[2] > 0.0:
        a.dropout_state, a.dropout_call, a.dropout_p = _ptr(dropout[0]), int(dropout[1]), float(dropout[2])


def attn_fwd(q, k, v, heads, causal=False, scale=None, bias=None, kv_len=None, dropout=None):
    """kv_len: optional int32 [B] key counts (key-padding mask): keys at or past kv_len[b] are ignored.
    dropout: optional (state, call, p) -- dropout on the probabilities, regenerated by attn_bwd from the same triple."""
    if FP32_MODE:
        if kv_len is not None:
            raise NotImplementedError("the fp32 verification kernels have no key-padding mask")
        from . import fp32path
        return fp32path.attn_fwd(q, k, v, heads, causal, (1.0 / math.sqrt(64)) if scale is None else scale, bias)
    B, Tq, HD = q.shape
    scale = (1.0 / math.sqrt(64)) if scale is None else scale
    o = torch.empty(B, Tq, HD, device=q.device, dtype=BF16)
    lse = torch.empty(B, heads, Tq, device=q.device, dtype=torch.float32)
    a = _attn_desc(q, k, v, o, lse, heads, causal, scale, bias, kv_len)
    _set_dropout(a, dropout)
    _lib.check(_lib.load().smx_attn_fwd(ctypes.byref(a), _stream()), "smx_attn_fwd")
    return o, lse


def attn_bwd(do, q, k, v, o, lse, heads, causal=False, scale=None, bias=None, dq=None, dk=None, dv=None, dbias=None,
             kv_len=None, dropout=None):
    """dbias: optional zero-initialised fp32 [heads, Tq, Tk]; receives sum_b dS (gradient of the additive bias).
    kv_len: as in attn_fwd; the k / v rows at or past kv_len[b] must hold zeros (mask_rows), dk / dv are zero there."""
    B, Tq, HD = q.shape
    Tk = k.shape[1]
    scale = (1.0 / math.sqrt(64)) if scale is None else scale
    assert do.stride(2) == 1
    dq = torch.empty(B, Tq, HD, device=q.device, dtype=BF16) if dq is None else dq
    dk = torch.empty(B, Tk, HD, device=q.device, dtype=BF16) if dk is None else dk
    dv = torch.empty(B, Tk, HD, device=q.device, dtype=BF16) if dv is None else dv
    delta = torch.empty(B, heads, Tq, device=q.device, dtype=torch.float32)
    a = _attn_desc(q, k, v, o, lse, heads, causal, scale, bias, kv_len)
    a.d_o, a.dq, a.dk, a.dv, a.delta = _ptr(do), _ptr(dq), _ptr(dk), _ptr(dv), _ptr(delta)
    a.dbias = _ptr(dbias)
    a.do_row_stride, a.do_batch_stride = do.stride(1), do.stride(0)
    a.dq_row_stride, a.dk_row_stride, a.dv_row_stride = dq.stride(1), dk.stride(1), dv.stride(1)
    a.dq_batch_stride, a.dk_batch_stride, a.dv_batch_stride = dq.stride(0), dk.stride(0), dv.stride(0)
    _set_dropout(a, dropout)
    _lib.check(_lib.load().smx_attn_bwd(ctypes.byref(a), _stream()), "smx_attn_bwd")
    return dq, dk, dv


# ---------------------------------------------------------------------------
# row-wise kernels
# ---------------------------------------------------------------------------
def _L():
    return _lib.load()


def layernorm_fwd(x, gamma, beta, eps=1e-5, res=None, want_sum=False, rms_only=False, act=ACT_NONE, out=None):
    if FP32_MODE:
        from . import fp32path
        return fp32path.layernorm_fwd(x, gamma, beta, eps, res, want_sum, rms_only, act, out)
    """x: [..., C] bf16 contiguous.  Returns y, (sum or x), mean, rstd."""
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x) if out is None else out
    s = torch.empty_like(x) if (want_sum and res is not None) else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(_L().smx_layernorm_fwd(_ptr(x), _ptr(res), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(s), _ptr(mean),
                                      _ptr(rstd), rows, C, eps, 1 if rms_only else 0, act, _stream()), "layernorm_fwd")
    return y, (s if s is not None else x), mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, rms_only=False, want_dbeta=True, act=ACT_NONE, beta=None,
                  want_colsum=False):
    """Returns dx, dgamma, dbeta (and, with want_colsum, the fp32 column sums of dx)."""
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    dgamma = zeros_f32(C, device=x.device)
    dbeta = zeros_f32(C, device=x.device) if want_dbeta else None
    csum = zeros_f32(C, device=x.device) if want_colsum else None
    _lib.check(_L().smx_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), _ptr(dres),
                                      _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(csum), rows, C, 1 if rms_only else 0, act,
                                      _stream()), "layernorm_bwd")
    return (dx, dgamma, dbeta, csum) if want_colsum else (dx, dgamma, dbeta)


def layernorm_dropout_fwd(x, res, gamma, beta, eps, state, call, p):
    """post-LN block tail with output dropout in one launch: s = dropout(x) + res, y = LN(s).  Returns y, s, mean, rstd
    (same mask as ``dropout(x, state, call, p, residual=res)``)."""
    C = x.shape[-1]
    rows = x.numel() // C
    assert x.dtype == BF16 and res.dtype == BF16 and x.is_contiguous() and res.is_contiguous() and res.shape == x.shape
    y, s = torch.empty_like(x), torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(_L().smx_layernorm_dropout_fwd(_ptr(x), _ptr(res), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(s), _ptr(mean),
                                              _ptr(rstd), rows, C, eps, _ptr(state), int(call), float(p), _stream()),
               "layernorm_dropout_fwd")
    return y, s, mean, rstd


def layernorm_dropout_bwd(dy, x, gamma, mean, rstd, state, call, p, want_dbeta=True, want_colsum=True):
    """backward of the same tail in one launch.  Returns dx (residual-branch gradient), dx_drop (= dropout mask applied to
    dx: the gradient of the sub-block output), dgamma, dbeta, column sums of dx_drop."""
    C = x.shape[-1]
    rows = x.numel() // C
    assert dy.dtype == BF16 and x.dtype == BF16 and dy.is_contiguous() and x.is_contiguous()
    dx, dxd = torch.empty_like(x), torch.empty_like(x)
    dgamma = zeros_f32(C, device=x.device)
    dbeta = zeros_f32(C, device=x.device) if want_dbeta else None
    csum = zeros_f32(C, device=x.device) if want_colsum else None
    _lib.check(_L().smx_layernorm_dropout_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dxd),
                                              _ptr(dgamma), _ptr(dbeta), _ptr(csum), rows, C, _ptr(state), int(call),
                                              float(p), _stream()), "layernorm_dropout_bwd")
    return dx, dxd, dgamma, dbeta, csum


def mask_rows(x, lens, col_begin=0, col_count=None):
    """In place: x[b, t, col_begin:col_begin+col_count] = 0 for t >= lens[b]  (x: bf16 [B, T, C], lens: int32 [B])."""
    B, T, C = x.shape
    assert x.dtype == BF16 and x.stride(2) == 1 and lens.dtype == torch.int32 and lens.numel() == B
    col_count = C - col_begin if col_count is None else col_count
    _lib.check(_L().smx_mask_rows(_ptr(x), _ptr(lens), B, T, x.stride(1), x.stride(0), col_begin, col_count, _stream()),
               "mask_rows")
    return x


def dropout(x, state, call, p, residual=None, aux_in=None, aux_mode=0):
    """out = residual + keep ? x / (1 - p) : 0 (bf16, any shape, numel % 8 == 0); with aux_mode 1 / 2 also returns the
    masked multiplier tensor for the activation-gradient epilogue (see smx_dropout).  state: int64 [2] on the device."""
    x = x.contiguous()
    assert x.dtype == BF16 and state.dtype == torch.int64 and state.numel() == 2
    out = torch.empty_like(x)
    aux_out = torch.empty_like(x) if aux_mode else None
    res = None if residual is None else residual.contiguous()
    aux = None if aux_in is None else aux_in.contiguous()
    _lib.check(_L().smx_dropout(_ptr(x), _ptr(res), _ptr(out), _ptr(aux), _ptr(aux_out), aux_mode, x.numel(), _ptr(state),
                                int(call), float(p), _stream()), "dropout")
    return (out, aux_out) if aux_mode else out


def dropout_mask(shape, state, call, p, attention=False):
    """keep decisions (uint8, `shape`) of dropout call `call` -- tests feed them to the CPU reference run.  attention=True:
    shape [B, H, Tq, Tk], numbered the way the attention kernels number their probabilities."""
    mask = torch.empty(shape, dtype=torch.uint8, device=state.device)
    _lib.check(_L().smx_dropout_mask(_ptr(mask), mask.numel(), shape[-1] if attention else 0, 1 if attention else 0,
                                     _ptr(state), int(call), float(p), _stream()), "dropout_mask")
    return mask


def spec_augment_fwd(x, time_mask, feat_mask, embed):
    """x: bf16 [B, T, H] contiguous; time_mask: uint8 [B, T] or None; feat_mask: uint8 [B, H] or None; embed: fp32 [H]."""
    B, T, H = x.shape
    y = torch.empty_like(x)
    _lib.check(_L().smx_spec_augment_fwd(_ptr(x), _ptr(y), _ptr(time_mask), _ptr(feat_mask), _ptr(embed), B, T, H,
                                         _stream()), "spec_augment_fwd")
    return y


def spec_augment_bwd(dy, time_mask, feat_mask, want_dembed=True):
    B, T, H = dy.shape
    dx = torch.empty_like(dy)
    dembed = zeros_f32(H, device=dy.device) if (want_dembed and time_mask is not None) else None
    _lib.check(_L().smx_spec_augment_bwd(_ptr(dy), _ptr(dx), _ptr(dembed), _ptr(time_mask), _ptr(feat_mask), B, T, H,
                                         _stream()), "spec_augment_bwd")
    return dx, dembed


def colsum(x2d):
    """fp32 column sums of a [rows, cols] bf16 matrix (row stride may exceed cols)."""
    rows, cols = x2d.shape
    out = zeros_f32(cols, device=x2d.device)
    assert x2d.stride(1) == 1
    _lib.check(_L().smx_colsum(_ptr(x2d), _ptr(out), rows, cols, x2d.stride(0), _stream()), "colsum")
    return out


def to_bf16(src):
    """fp32 -> bf16 copy through the library (weights are cached as bf16 between optimizer steps)."""
    src = src.detach().contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=BF16)
    if src.numel() % 8 == 0 and src.data_ptr() % 16 == 0:
        _lib.check(_L().smx_cast_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "cast")
    else:
        dst.copy_(src)
    return dst


def build_cast_table(rows, device):
    """rows: [(src_ptr, dst_ptr, n, dst_is_f32)] -> (device table of SmxCastEntry, n_entries, total_chunks)."""
    import numpy as np
    dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("n", "<i8"), ("f32", "<i4"), ("first", "<i4")])
    tab = np.zeros(len(rows), dtype=dt)
    chunk, first = 4096, 0
    for i, (s, d, n, f) in enumerate(rows):
        tab[i] = (s, d, n, f, first)
        first += (n + chunk - 1) // chunk
    host = torch.from_numpy(tab.view(np.uint8).copy()).pin_memory()   # pinned: the copy is legal inside a graph capture
    t = host.to(device, non_blocking=True)
    t._smx_host = host                                                 # keep the staging buffer alive with the table
    return t, len(rows), first


def multi_cast(table, n_entries, total_chunks):
    _lib.check(_L().smx_multi_cast(_ptr(table), n_entries, total_chunks, _stream()), "multi_cast")


def add_bf16(a, b):
    out = torch.empty_like(a)
    _lib.check(_L().smx_add_bf16(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "add")
    return out


def mul_bf16(a, b):
    """a * b elementwise (bf16; fp32 tensors in the fp32 verification mode)"""
    if FP32_MODE or a.dtype != BF16:
        return a * b
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a)
    _lib.check(_L().smx_mul_bf16(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "mul")
    return out


def dact(dy, pre, act=ACT_GELU):
    out = torch.empty_like(dy)
    _lib.check(_L().smx_dact_bf16(_ptr(dy), _ptr(pre), _ptr(out), dy.numel(), act, _stream()), "dact")
    return out


# ---------------------------------------------------------------------------
# conv0 + GroupNorm + GELU
# ---------------------------------------------------------------------------
def conv0_fwd(audio, w, gamma, beta, eps=1e-5, want_gprime=False):
    """audio [B, n] fp32, w [C, 1, 10] fp32.  Returns y [B, T, C] bf16 (slack-padded), stats, moments and, with
    want_gprime (training), gelu'(z) [B, T, C] bf16 for the streaming backward."""
    if FP32_MODE:
        from . import fp32path
        return fp32path.conv0_fwd(audio, w, gamma, beta, eps=eps) + ((None,) if want_gprime else ())
    B, n = audio.shape
    C, _, k = w.shape
    s = 5
    T = (n - k) // s + 1
    moments = torch.empty(B, 110, device=audio.device, dtype=torch.float32)
    stats = torch.empty(B, C, 2, device=audio.device, dtype=torch.float32)
    L = _L()
    part = torch.empty(B, 32, 65, device=audio.device, dtype=torch.float32)    # SMX_CONV0_MOMENT_BLOCKS x partials
    _lib.check(L.smx_conv0_stats(_ptr(audio), _ptr(w), _ptr(moments), _ptr(stats), _ptr(part), B, n, T, C, k, s, eps,
                                 _stream()), "conv0_stats")
    y = alloc_act(B, T, C, audio.device)
    gp = torch.empty(B, T, C, device=audio.device, dtype=BF16) if want_gprime else None
    _lib.check(L.smx_conv0_gn_gelu_fwd(_ptr(audio), _ptr(w), _ptr(gamma), _ptr(beta), _ptr(stats), _ptr(y), _ptr(gp), B, n, T,
                                       C, k, s, _stream()), "conv0_fwd")
    return (y, stats, moments, gp) if want_gprime else (y, stats, moments)


def conv0_bwd(audio, w, gamma, beta, stats, moments, dy, gprime):
    B, n = audio.shape
    C, _, k = w.shape
    T = dy.shape[1]
    partial = torch.empty(B, C, k + 2, device=audio.device, dtype=torch.float32)
    dw = torch.empty_like(w)
    dgamma = torch.empty(C, device=audio.device, dtype=torch.float32)
    dbeta = torch.empty(C, device=audio.device, dtype=torch.float32)
    _lib.check(_L().smx_conv0_gn_gelu_bwd(_ptr(audio), _ptr(w), _ptr(gamma), _ptr(beta), _ptr(stats), _ptr(moments),
                                          _ptr(dy), _ptr(gprime), _ptr(partial), _ptr(dw), _ptr(dgamma), _ptr(dbeta), B, n,
                                          T, C, k, 5, _stream()), "conv0_bwd")
    return dw, dgamma, dbeta


def conv0_ln_fwd(audio, w, conv_bias, gamma, beta, eps=1e-5):
    if FP32_MODE:
        from . import fp32path
        return fp32path.conv0_ln_fwd(audio, w, conv_bias, gamma, beta, eps)
    B, n = audio.shape
    C, _, k = w.shape
    T = (n - k) // 5 + 1
    y = alloc_act(B, T, C, audio.device)
    _lib.check(_L().smx_conv0_ln_gelu_fwd(_ptr(audio), _ptr(w), _ptr(conv_bias), _ptr(gamma), _ptr(beta), _ptr(y), B, n,
                                          T, C, k, 5, eps, _stream()), "conv0_ln_fwd")
    return y


def conv0_ln_bwd(audio, w, conv_bias, gamma, beta, dy, eps=1e-5):
    B, n = audio.shape
    C, _, k = w.shape
    T = dy.shape[1]
    dconv = torch.empty(B, T, C, device=audio.device, dtype=BF16)
    dgamma = torch.zeros(C, device=audio.device, dtype=torch.float32)
    dbeta = torch.zeros(C, device=audio.device, dtype=torch.float32)
    L = _L()
    _lib.check(L.smx_conv0_ln_gelu_bwd(_ptr(audio), _ptr(w), _ptr(conv_bias), _ptr(gamma), _ptr(beta), _ptr(dy),
                                       _ptr(dconv), _ptr(dgamma), _ptr(dbeta), B, n, T, C, k, 5, eps, _stream()),
               "conv0_ln_bwd")
    dw = torch.zeros_like(w)
    dcb = torch.zeros(C, device=audio.device, dtype=torch.float32) if conv_bias is not None else None
    _lib.check(L.smx_conv0_wgrad(_ptr(audio), _ptr(dconv), _ptr(dw), _ptr(dcb), B, n, T, C, k, 5, _stream()),
               "conv0_wgrad")
    return dw, dcb, dgamma, dbeta


# ---------------------------------------------------------------------------
# positional conv embedding
# ---------------------------------------------------------------------------
def posconv_pack(weight, groups):
    """weight [H, cg, k] fp32 (Conv1d layout, out-major) ->
    (fwd pack, dgrad pack), each bf16 [G][k][K/8][N][8] with
      fwd:   Wp[g][tap][c][o]   = w[g*cg+o, c, tap]
      dgrad: Wp[g][tap''][o][c] = w[g*cg+o, c, k-1-tap'']."""
    H, cg, k = weight.shape
    w = weight.detach().view(groups, cg, cg, k)              # [g][o][c][tap]
    fwd = w.permute(0, 3, 2, 1)                              # [g][tap][c(K)][o(N)]
    dgr = w.flip(3).permute(0, 3, 1, 2)                      # [g][tap''][o(K)][c(N)]

    def pack(m):  # [g][tap][K][N] -> [g][tap][K/8][N][8]
        return m.reshape(groups, k, cg // 8, 8, cg).permute(0, 1, 2, 4, 3).to(BF16).contiguous()

    return pack(fwd), pack(dgr)


def weightnorm_fwd(v, g):
    """v [H, cg, k] fp32, g [1, 1, k] -> (w [H, cg, k], sq [k])"""
    v = v.detach().contiguous()
    k = v.shape[-1]
    rows = v.numel() // k
    w = torch.empty_like(v)
    sq = torch.empty(k, device=v.device, dtype=torch.float32)
    _lib.check(_L().smx_weightnorm_fwd(_ptr(v), _ptr(g.detach().reshape(-1).contiguous()), _ptr(sq), _ptr(w), rows, k,
                                       _stream()), "weightnorm_fwd")
    return w, sq


def weightnorm_bwd(v, g, sq, dw):
    v = v.detach().contiguous()
    k = v.shape[-1]
    rows = v.numel() // k
    dv = torch.empty_like(v)
    dg = torch.empty(k, device=v.device, dtype=torch.float32)
    dot = torch.empty(k, device=v.device, dtype=torch.float32)
    _lib.check(_L().smx_weightnorm_bwd(_ptr(v), _ptr(g.detach().reshape(-1).contiguous()), _ptr(sq), _ptr(dw.contiguous()),
                                       _ptr(dot), _ptr(dv), _ptr(dg), rows, k, _stream()), "weightnorm_bwd")
    return dv, dg


def posconv_fwd(x, w_fwd, bias, groups, ksize, add_input=True):
    if FP32_MODE:
        from . import fp32path
        return fp32path.posconv_fwd(x, w_fwd, bias, groups, ksize, add_input)
    B, T, H = x.shape
    y = torch.empty_like(x)
    pre = torch.empty_like(x)
    _lib.check(_L().smx_posconv_fwd(_ptr(x), _ptr(w_fwd), _ptr(bias), _ptr(y), _ptr(pre), B, T, H, groups, ksize,
                                    1 if add_input else 0, _stream()), "posconv_fwd")
    return y, pre


def posconv_dgrad(dpre, w_dgrad, groups, ksize, residual=None):
    B, T, H = dpre.shape
    dx = torch.empty_like(dpre)
    _lib.check(_L().smx_posconv_dgrad(_ptr(dpre), _ptr(w_dgrad), _ptr(residual), _ptr(dx), B, T, H, groups, ksize,
                                      _stream()), "posconv_dgrad")
    return dx


def posconv_wgrad(dpre, x, groups, ksize):
    """Returns dweight in Conv1d layout [H, cg, k] fp32."""
    B, T, H = x.shape
    cg = H // groups
    dw = zeros_f32(groups, ksize, cg, cg, device=x.device)  # [g][tap][o][c]
    _lib.check(_L().smx_posconv_wgrad(_ptr(dpre), _ptr(x), _ptr(dw), B, T, H, groups, ksize, _stream()),
               "posconv_wgrad")
    return dw.permute(0, 2, 3, 1).reshape(H, cg, ksize).contiguous()


# ---------------------------------------------------------------------------
# embeddings
# ---------------------------------------------------------------------------
def embed_fwd(ids, tok_emb, pos_emb, x_in, batch, t, dim, scale=1.0, pos_offset=0, t_start=0, device=None):
    if FP32_MODE:
        from . import fp32path
        return fp32path.embed_fwd(ids, tok_emb, pos_emb, x_in, batch, t, dim, scale, pos_offset, t_start, device)
    out = torch.empty(batch, t, dim, device=device, dtype=BF16)
    _lib.check(_L().smx_embed_fwd(_ptr(ids), _ptr(tok_emb), _ptr(pos_emb), _ptr(x_in), _ptr(out), batch, t, dim, scale,
                                  pos_offset, t_start, _stream()), "embed_fwd")
    return out


def embed_bwd(ids, dout, d_tok, d_pos, scale=1.0, pos_offset=0):
    B, T, D = dout.shape
    _lib.check(_L().smx_embed_bwd(_ptr(ids), _ptr(dout), _ptr(d_tok), _ptr(d_pos), B, T, D, scale, pos_offset,
                                  _stream()), "embed_bwd")


# ---------------------------------------------------------------------------
# LM head + cross entropy
# ---------------------------------------------------------------------------
def lmhead_ce_fwd(h, emb16, bias, labels, logit_scale=1.0, ignore_index=-100):
    if FP32_MODE:
        from . import fp32path
        return fp32path.lmhead_ce_fwd(h, emb16, bias, labels, logit_scale, ignore_index)
    """h [M, D] bf16, emb16 [V, D] bf16, labels [M] int64 -> lse, argmax, row_loss, loss_sum, count"""
    M, D = h.shape
    V = emb16.shape[0]
    L = _L()
    ws = torch.empty(L.smx_lmhead_ws_bytes(M, V) // 4 + 4, device=h.device, dtype=torch.float32)
    lse = torch.empty(M, device=h.device, dtype=torch.float32)
    argmax = torch.empty(M, device=h.device, dtype=torch.int64)
    row_loss = torch.empty(M, device=h.device, dtype=torch.float32)
    acc = torch.zeros(2, device=h.device, dtype=torch.float32)
    _lib.check(L.smx_lmhead_ce_fwd(_ptr(h), _ptr(emb16), _ptr(bias), _ptr(labels), _ptr(lse), _ptr(argmax),
                                   _ptr(row_loss), _ptr(acc), _ptr(acc, 1), _ptr(ws), M, D, V, logit_scale,
                                   ignore_index, _stream()), "lmhead_ce_fwd")
    return lse, argmax, row_loss, acc


def lmhead_dlogits(h, emb16, bias, labels, lse, coef, buf, v0, vn, logit_scale=1.0):
    M, D = h.shape
    V = emb16.shape[0]
    _lib.check(_L().smx_lmhead_dlogits(_ptr(h), _ptr(emb16), _ptr(bias), _ptr(labels), _ptr(lse), _ptr(coef),
                                       _ptr(buf), buf.stride(0), M, D, V, v0, vn, logit_scale, _stream()),
               "lmhead_dlogits")


def gemm_nn_acc_f32(a, a_cols, w_rows_view, out_f32, alpha=1.0, accumulate=True):
    """out_f32[M, K] (+)= a[M, :a_cols] @ w_rows_view[a_cols, K]  (a may have a larger row stride)."""
    M = a.shape[0]
    K = w_rows_view.shape[1]
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_NN, OUT_F32
    g.a = _view(a, a_cols, M, 1, a.stride(0), M * a.stride(0))
    g.b = _view(w_rows_view, K, a_cols, 1, w_rows_view.stride(0), a_cols * w_rows_view.stride(0))
    g.m, g.n, g.k, g.batches = M, K, a_cols, 1
    _set_seg(g, 1, a_cols)
    g.accumulate = 1 if accumulate else 0
    if accumulate:      # few output tiles, long contraction: split it (fp32 reductions into the accumulating output)
        tiles = math.ceil(M / 256) * math.ceil(K / 256)
        g.split_k = _pick_split(tiles, math.ceil(a_cols / 64))
    _epilogue(g, out_f32, K, M * K, None, ACT_NONE, None, None, None, None, alpha)
    _run_gemm(g)


def gemm_tn_into(a, a_cols, x, out_rows_view, alpha=1.0):
    """out_rows_view[a_cols, K] (fp32, plain store) = a[M, :a_cols]^T @ x[M, K]."""
    M = a.shape[0]
    K = x.shape[1]
    g = SmxGemm()
    g.mode, g.out_dtype = GEMM_TN, OUT_F32
    g.a = _view(a, a_cols, M, 1, a.stride(0), M * a.stride(0))
    g.b = _view(x, K, M, 1, x.stride(0), M * x.stride(0))
    g.m, g.n, g.k, g.batches = a_cols, K, M, 1
    _set_seg(g, 1, K)
    g.split_k = 1
    g.c = _ptr(out_rows_view)
    g.c_row_stride, g.c_batch_stride = out_rows_view.stride(0), a_cols * out_rows_view.stride(0)
    g.alpha = alpha
    _run_gemm(g)


# ---------------------------------------------------------------------------
# weighted layer sum
# ---------------------------------------------------------------------------
def weighted_sum_fwd(xs, w):
    if FP32_MODE:
        from . import fp32path
        return fp32path.weighted_sum_fwd(xs, w)
    n = xs[0].numel()
    out = torch.empty_like(xs[0])
    arr = (ctypes.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    _lib.check(_L().smx_weighted_sum_fwd(arr, _ptr(w), _ptr(out), len(xs), n, _stream()), "wsum_fwd")
    return out


def weighted_sum_bwd_w(xs, dout):
    n = xs[0].numel()
    dw = torch.zeros(len(xs), device=dout.device, dtype=torch.float32)
    arr = (ctypes.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    _lib.check(_L().smx_weighted_sum_bwd_w(arr, _ptr(dout), _ptr(dw), len(xs), n, _stream()), "wsum_bwd")
    return dw


# ---------------------------------------------------------------------------
# SpeechMixSelf losses, T5 relative position bias
# ---------------------------------------------------------------------------
def logits_chunk_f32(h, emb16_rows, bias_rows, alpha, out):
    """out[M, vn] fp32 = alpha * h @ emb16_rows^T + bias_rows  (one vocabulary chunk of the LM head)."""
    return linear_fwd(h, emb16_rows, bias=bias_rows, out_f32=True, alpha=alpha, out=out)


def kl_chunk_fwd(s, t, vn, lse_t, cross):
    _lib.check(_L().smx_kl_chunk_fwd(_ptr(s), _ptr(t), s.stride(0), s.shape[0], vn, _ptr(lse_t), _ptr(cross), _stream()),
               "kl_chunk_fwd")


def kl_finalize(cross, lse_s, lse_t, inv_batch):
    out = torch.empty(1, device=cross.device, dtype=torch.float32)
    _lib.check(_L().smx_kl_finalize(_ptr(cross), _ptr(lse_s), _ptr(lse_t), cross.numel(), inv_batch, _ptr(out), _stream()),
               "kl_finalize")
    return out


def kl_chunk_bwd(s, t, vn, v0, labels, lse_s, lse_t, coef_ce, coef_kl, dlogits):
    _lib.check(_L().smx_kl_chunk_bwd(_ptr(s), _ptr(t), s.stride(0), s.shape[0], vn, v0, _ptr(labels), _ptr(lse_s),
                                     _ptr(lse_t), _ptr(coef_ce), _ptr(coef_kl), _ptr(dlogits), dlogits.stride(0),
                                     _stream()), "kl_chunk_bwd")


def self_mse_fwd(text_h, speech_h):
    """text_h [B,Tt,D], speech_h [B,Ts,D] bf16 contiguous -> (loss[1] fp32, attn, diff)"""
    B, Tt, D = text_h.shape
    Ts = speech_h.shape[1]
    attn = torch.empty(B, Tt, Ts, device=text_h.device, dtype=torch.float32)
    diff = torch.empty(B, Tt, D, device=text_h.device, dtype=torch.float32)
    loss = torch.zeros(1, device=text_h.device, dtype=torch.float32)
    _lib.check(_L().smx_self_mse_fwd(_ptr(text_h), _ptr(speech_h), _ptr(attn), _ptr(diff), _ptr(loss), B, Tt, Ts, D,
                                     _stream()), "self_mse_fwd")
    return loss, attn, diff


def self_mse_bwd(text_h, speech_h, attn, diff, gscale):
    B, Tt, D = text_h.shape
    Ts = speech_h.shape[1]
    dsc = torch.empty(B, Tt, Ts, device=text_h.device, dtype=torch.float32)
    ds = torch.empty_like(speech_h)
    _lib.check(_L().smx_self_mse_bwd(_ptr(text_h), _ptr(speech_h), _ptr(attn), _ptr(diff), _ptr(dsc), _ptr(gscale),
                                     _ptr(ds), B, Tt, Ts, D, _stream()), "self_mse_bwd")
    return ds


def gram_dot_fwd(x, z):
    """x, z [B,T,D] bf16 contiguous -> out[B] fp32 = sum_{t,i} xflat[b][i T + t] z[b][t][i]  (SpeechMixGAN features)"""
    B, T, D = x.shape
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    _lib.check(_L().smx_gram_dot_fwd(_ptr(x), _ptr(z), _ptr(out), B, T, D, _stream()), "gram_dot_fwd")
    return out


def gram_dot_bwd(x, z, g):
    """g [B] fp32 -> (dx, dz) bf16 [B,T,D]: gradients of gram_dot_fwd w.r.t. its two arguments"""
    B, T, D = x.shape
    dx, dz = torch.empty_like(x), torch.empty_like(z)
    _lib.check(_L().smx_gram_dot_bwd(_ptr(x), _ptr(z), _ptr(g), _ptr(dx), _ptr(dz), B, T, D, _stream()), "gram_dot_bwd")
    return dx, dz


def relpos_bias_fwd(weight, table, heads, tq, tk, q_offset=0):
    bias = torch.empty(heads, tq, tk, device=weight.device, dtype=torch.float32)
    _lib.check(_L().smx_relpos_bias_fwd(_ptr(weight), _ptr(table), _ptr(bias), heads, tq, tk, q_offset, _stream()),
               "relpos_fwd")
    return bias


def relpos_bias_bwd(dbias, table, heads, tq, tk, n_buckets, q_offset=0):
    dw = torch.zeros(n_buckets, heads, device=dbias.device, dtype=torch.float32)
    _lib.check(_L().smx_relpos_bias_bwd(_ptr(dbias), _ptr(table), _ptr(dw), heads, tq, tk, q_offset, n_buckets,
                                        _stream()), "relpos_bwd")
    return dw
