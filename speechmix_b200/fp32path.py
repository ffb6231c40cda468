"""fp32 verification path: tensor-level wrappers over the ``smx_f32_*`` entry points (csrc/fp32.cu).

Inference only.  ``kernels.py`` dispatches here when verification mode is on (``ops.fp32_verification()``):
activations are fp32, parameters are used as their fp32 masters, every contraction is an fp32 CUDA-core GEMM.
The point is bit-exact greedy ids against the reference's fp32 run (BASELINE.json north_star), not speed.
"""
import ctypes

import torch

from . import _lib

F32 = torch.float32
CHUNK = 4096


def _L():
    return _lib.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, off=0):
    return None if t is None else ctypes.c_void_p(t.data_ptr() + off * t.element_size())


def _f(t):
    return None if t is None else t.detach().to(F32).contiguous()


def gemm_nt(a, lda, a_bs, w, bias, residual, ldr, r_bs, c, ldc, c_bs, m, n, k, batches, act=0, alpha=1.0):
    _lib.check(_L().smx_f32_gemm_nt(_p(a), lda, a_bs, _p(w), _p(bias), _p(residual), ldr, r_bs, _p(c), ldc, c_bs, m, n, k,
                                    batches, act, alpha, _stream()), "f32_gemm")


def linear_fwd(x, w, bias=None, act=0, residual=None, want_pre=False, out_f32=False, alpha=1.0, out=None):
    assert x.dtype == F32 and x.stride(1) == 1
    M, K = x.shape
    w = _f(w)
    N = w.shape[0]
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=F32)
    r = residual
    gemm_nt(x, x.stride(0), 0, w, _f(bias), r, 0 if r is None else r.stride(0), 0, y, y.stride(0), 0, M, N, K, 1, act, alpha)
    return (y, None) if want_pre else y


def layernorm_fwd(x, gamma, beta, eps=1e-5, res=None, want_sum=False, rms_only=False, act=0, out=None):
    assert res is None, "fused residual input is not used on the inference path"
    C = x.shape[-1]
    x = x.contiguous()
    y = torch.empty_like(x) if out is None else out
    _lib.check(_L().smx_f32_layernorm(_p(x), _p(_f(gamma)), _p(_f(beta)), _p(y), x.numel() // C, C, eps,
                                      1 if rms_only else 0, act, _stream()), "f32_layernorm")
    return y, x, None, None


def attn_fwd(q, k, v, heads, causal=False, scale=0.125, bias=None):
    B, Tq, HD = q.shape
    Tk = k.shape[1]
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and HD == heads * 64
    o = torch.empty(B, Tq, HD, device=q.device, dtype=F32)
    _lib.check(_L().smx_f32_attn(_p(q), _p(k), _p(v), _p(o), q.stride(1), q.stride(0), k.stride(1), k.stride(0), v.stride(1),
                                 v.stride(0), o.stride(1), o.stride(0), B, heads, Tq, Tk, 1 if causal else 0, scale,
                                 _p(_f(bias)), _stream()), "f32_attn")
    return o, None


def conv_frames(x, n_in, lda, w2d, bias, t_out, act=0):
    """x: [B, *] fp32 whose row t of batch b starts at x[b].flatten()[t * lda] and spans w2d.shape[1] floats."""
    B = x.shape[0]
    N, K = w2d.shape
    y = torch.empty(B, t_out, N, device=x.device, dtype=F32)
    gemm_nt(x, lda, n_in, w2d, _f(bias), None, 0, 0, y, N, t_out * N, t_out, N, K, B, act, 1.0)
    return y


def conv0_fwd(audio, w, gamma, beta, k=10, s=5, eps=1e-5):
    """Conv1d(1 -> C, k, s, no bias) + GroupNorm(C groups) + GELU   hf:models/wav2vec2/modeling_wav2vec2.py:302-323"""
    audio = audio.contiguous().to(F32)
    B, n = audio.shape
    C = w.shape[0]
    T = (n - k) // s + 1
    y = conv_frames(audio, n, s, _f(w).view(C, k), None, T)
    ws = torch.empty(B * C * 2, device=audio.device, dtype=torch.float64)
    _lib.check(_L().smx_f32_groupnorm_gelu(_p(y), _p(ws), _p(_f(gamma)), _p(_f(beta)), B, T, C, eps, _stream()), "f32_gn")
    return y, None, None


def conv0_ln_fwd(audio, w, conv_bias, gamma, beta, eps=1e-5, k=10, s=5):
    audio = audio.contiguous().to(F32)
    B, n = audio.shape
    C = w.shape[0]
    T = (n - k) // s + 1
    z = conv_frames(audio, n, s, _f(w).view(C, k), conv_bias, T)
    return layernorm_fwd(z.view(B * T, C), gamma, beta, eps, act=1)[0].view(B, T, C)


def pack_conv_weight(w):
    """[out, in, k] -> [out, k*in] (tap-major), fp32"""
    return w.detach().to(F32).permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()


def conv_s2_fwd(x, w_packed, k, bias=None, act=0, want_pre=False):
    B, T_in, C = x.shape
    x = x.contiguous()
    T_out = (T_in - k) // 2 + 1
    y = conv_frames(x, T_in * C, 2 * C, w_packed, bias, T_out, act)
    return (y, None) if want_pre else y


def posconv_fwd(x, weight, bias, groups, ksize, add_input=True):
    B, T, H = x.shape
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(_L().smx_f32_posconv(_p(x), _p(_f(weight)), _p(_f(bias)), _p(y), B, T, H, groups, ksize,
                                    1 if add_input else 0, _stream()), "f32_posconv")
    return y, None


def embed_fwd(ids, tok, pos, x_in, batch, t, dim, scale, pos_offset, t_start, device):
    out = torch.empty(batch, t, dim, device=device, dtype=F32)
    xi = None if x_in is None else x_in.to(F32).contiguous()
    _lib.check(_L().smx_f32_embed(_p(ids), _p(_f(tok)), _p(_f(pos)), _p(xi), _p(out), batch, t, dim, scale,
                                  pos_offset + t_start, _stream()), "f32_embed")
    return out


def lmhead_ce_fwd(h, emb, bias, labels, logit_scale=1.0, ignore_index=-100):
    """fp32 logits one [M, 4096] chunk at a time: running argmax (lowest index wins) and the CE loss."""
    M, D = h.shape
    emb = _f(emb)
    V = emb.shape[0]
    best = torch.empty(M, device=h.device, dtype=F32)
    idx = torch.empty(M, device=h.device, dtype=torch.int64)
    buf = torch.empty(M, CHUNK, device=h.device, dtype=F32)
    run_max = torch.full((M,), float("-inf"), device=h.device, dtype=F32)
    run_sum = torch.zeros(M, device=h.device, dtype=F32)
    lab_logit = torch.zeros(M, device=h.device, dtype=F32)
    b = _f(bias)
    for v0 in range(0, V, CHUNK):
        vn = min(CHUNK, V - v0)
        linear_fwd(h, emb[v0:v0 + vn], None if b is None else b[v0:v0 + vn], alpha=logit_scale, out=buf[:, :vn])
        _lib.check(_L().smx_f32_argmax_chunk(_p(buf), CHUNK, M, vn, v0, _p(best), _p(idx), _stream()), "f32_argmax")
        # loss bookkeeping on [M]-sized vectors (verification only)
        chunk = buf[:, :vn]
        cmax = torch.maximum(run_max, chunk.max(dim=1).values)
        run_sum = run_sum * torch.exp(run_max - cmax) + torch.exp(chunk - cmax[:, None]).sum(dim=1)
        run_max = cmax
        inside = (labels >= v0) & (labels < v0 + vn)
        if bool(inside.any()):
            rows = inside.nonzero().squeeze(1)
            lab_logit[rows] = chunk[rows, labels[rows] - v0]
    lse = run_max + torch.log(run_sum)
    valid = labels != ignore_index
    row_loss = torch.where(valid, lse - lab_logit, torch.zeros_like(lse))
    acc = torch.stack([row_loss.sum(), valid.sum().to(F32)])
    return lse, idx, row_loss, acc


def weighted_sum_fwd(xs, w):
    out = torch.empty_like(xs[0])
    for i, x in enumerate(xs):
        _lib.check(_L().smx_f32_axpy(_p(x.contiguous()), _p(w), i, _p(out), out.numel(), 1 if i == 0 else 0, _stream()),
                   "f32_axpy")
    return out
