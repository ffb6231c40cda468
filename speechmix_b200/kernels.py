"""Tensor-level wrappers over the C ABI (no autograd here; see ``ops.py``).

Every function takes CUDA tensors, enqueues one or more kernels of
``libspeechmix_sm100.so`` on the current torch stream and returns the output
tensors.  Activations are bf16, channels-last.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (ACT_DGELU, ACT_DRELU, ACT_GELU, ACT_NONE, ACT_RELU, GEMM_NN, GEMM_NT, GEMM_TN, OUT_BF16,
                   OUT_F32, SmxAttn, SmxGemm, SmxView3)

BF16 = torch.bfloat16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, offset_elems=0):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _view(t, inner, rows, batches, row_stride, batch_stride, offset=0):
    return SmxView3(_ptr(t, offset), inner, rows, batches, row_stride, batch_stride)


def alloc_act(batch, t, c, device, dtype=BF16, slack=None):
    """[batch, t, c] activation with ``slack`` (default c) zeroed elements after
    the end, so frame-pair TMA views of an odd-length signal stay in bounds."""
    slack = c if slack is None else slack
    n = batch * t * c
    flat = torch.empty(n + slack, device=device, dtype=dtype)
    flat[n:].zero_()
    return flat[:n].view(batch, t, c)


def _run_gemm(g):
    lib = _lib.load()
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
    assert x.dtype == BF16 and w.dtype == BF16 and x.is_contiguous() and w.is_contiguous()
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
    _run_gemm(g)
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


def _pick_split(out_tiles, kblocks, sms=148):
    if out_tiles >= sms:
        return 1
    s = max(1, sms // out_tiles)
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
    tiles = math.ceil(N / 128) * math.ceil(K / 256)
    g.split_k = _pick_split(tiles, math.ceil(M / 64))
    g.accumulate = 1 if accumulate else 0
    if out is None:
        dw = (torch.zeros if g.split_k > 1 else torch.empty)(N, K, device=dy.device, dtype=torch.float32)
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
    B, T_in, C = x.shape
    N = w_packed.shape[0]
    assert w_packed.shape[1] == k * C and C % 64 == 0 and x.is_contiguous()
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
    tiles = math.ceil(N / 128) * math.ceil(k * C / 256)
    g.split_k = _pick_split(tiles, B * math.ceil(T_out / 64))
    dw = (torch.zeros if g.split_k > 1 else torch.empty)(N, k * C, device=dy.device, dtype=torch.float32)
    g.c = _ptr(dw)
    g.c_row_stride, g.c_batch_stride = k * C, N * k * C
    g.alpha = 1.0
    _run_gemm(g)
    return dw


def pack_conv_weight(w):
    """[out, in, k] fp32 -> [out, k*in] bf16 (tap-major)."""
    out_c, in_c, k = w.shape
    return w.detach().permute(0, 2, 1).reshape(out_c, k * in_c).to(BF16).contiguous()


def unpack_conv_wgrad(dw_packed, in_c, k):
    out_c = dw_packed.shape[0]
    return dw_packed.view(out_c, k, in_c).permute(0, 2, 1).contiguous()


# ---------------------------------------------------------------------------
# attention (head_dim 64); q/k/v: [B, T, heads*64] views with unit inner stride
# ---------------------------------------------------------------------------
def _attn_desc(q, k, v, o, lse, heads, causal, scale, bias):
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
    return a


def attn_fwd(q, k, v, heads, causal=False, scale=None, bias=None):
    B, Tq, HD = q.shape
    scale = (1.0 / math.sqrt(64)) if scale is None else scale
    o = torch.empty(B, Tq, HD, device=q.device, dtype=BF16)
    lse = torch.empty(B, heads, Tq, device=q.device, dtype=torch.float32)
    a = _attn_desc(q, k, v, o, lse, heads, causal, scale, bias)
    _lib.check(_lib.load().smx_attn_fwd(ctypes.byref(a), _stream()), "smx_attn_fwd")
    return o, lse


def attn_bwd(do, q, k, v, o, lse, heads, causal=False, scale=None, bias=None, dq=None, dk=None, dv=None):
    B, Tq, HD = q.shape
    Tk = k.shape[1]
    scale = (1.0 / math.sqrt(64)) if scale is None else scale
    assert do.stride(2) == 1
    dq = torch.empty(B, Tq, HD, device=q.device, dtype=BF16) if dq is None else dq
    dk = torch.empty(B, Tk, HD, device=q.device, dtype=BF16) if dk is None else dk
    dv = torch.empty(B, Tk, HD, device=q.device, dtype=BF16) if dv is None else dv
    delta = torch.empty(B, heads, Tq, device=q.device, dtype=torch.float32)
    a = _attn_desc(q, k, v, o, lse, heads, causal, scale, bias)
    a.d_o, a.dq, a.dk, a.dv, a.delta = _ptr(do), _ptr(dq), _ptr(dk), _ptr(dv), _ptr(delta)
    a.do_row_stride, a.do_batch_stride = do.stride(1), do.stride(0)
    a.dq_row_stride, a.dk_row_stride, a.dv_row_stride = dq.stride(1), dk.stride(1), dv.stride(1)
    a.dq_batch_stride, a.dk_batch_stride, a.dv_batch_stride = dq.stride(0), dk.stride(0), dv.stride(0)
    _lib.check(_lib.load().smx_attn_bwd(ctypes.byref(a), _stream()), "smx_attn_bwd")
    return dq, dk, dv
