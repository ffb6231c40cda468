"""Small driver for ncu captures: runs each hot kernel a few times at the bench shapes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from speechmix_b200 import kernels as K

which = sys.argv[1:] or ["attn", "gemm"]
g = torch.Generator(device="cuda").manual_seed(0)
B, T, H = 32, 749, 12
def rounds():
    """one warm-up pass outside the capture range, one pass inside (ncu --profile-from-start off)"""
    yield 0
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    yield 1
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if "attn" in which:
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    do = torch.randn(B, T, H * 64, device="cuda", generator=g).to(torch.bfloat16)
    for _ in rounds():
        o, lse = K.attn_fwd(q, k, v, H)
        K.attn_bwd(do, q, k, v, o, lse, H)
if "gemm" in which:
    M, N, Kd = B * T, 3072, 768
    x = torch.randn(M, Kd, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, Kd, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    dy = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    for _ in rounds():
        y, pre = K.linear_fwd(x, w, bias=b, act=K.ACT_GELU_G, want_pre=True) # NT + GELU + gelu' (the step's dominant kernel)
        K.linear_fwd(x, w)                                                    # NT plain
        K.linear_wgrad(dy, x)                                                 # TN
        K.linear_dgrad(dy, w)                                                 # NN
torch.cuda.synchronize()
if "rowwise" in which:
    # HBM-bound kernels at the bench shapes: LayerNorm fwd/bwd over [B*749, 768], bias-gradient column sums,
    # weighted layer sum over 13 hidden states, conv0 + GroupNorm + GELU fwd/bwd over [B, 47999, 512]
    M, C = B * T, 768
    x = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    gamma = torch.ones(C, device="cuda")
    beta = torch.zeros(C, device="cuda")
    xs = [torch.randn(B, T, C, device="cuda", generator=g).to(torch.bfloat16) for _ in range(13)]
    w13 = torch.softmax(torch.zeros(13, device="cuda"), 0)
    audio = torch.randn(B, 240000, device="cuda", generator=g)
    w0 = torch.randn(512, 1, 10, device="cuda", generator=g) * 0.3
    g0, b0 = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda")
    for _ in rounds():
        y, _, mean, rstd = K.layernorm_fwd(x, gamma, beta)
        K.layernorm_bwd(dy, x, gamma, mean, rstd, want_colsum=True)
        K.colsum(dy)
        K.weighted_sum_fwd(xs, w13)
        y0, stats, mom, gp0 = K.conv0_fwd(audio, w0, g0, b0, want_gprime=True)
        K.conv0_bwd(audio, w0, g0, b0, stats, mom, y0, gp0)
torch.cuda.synchronize()
