"""GPU probe for the tcgen05 GEMM family: numerical check of every mode against
torch fp32 matmul on the same bf16 inputs, with per-panel error maps to localise
descriptor / swizzle mistakes, plus a throughput sweep.  Development tool (run
under gpurun); the pytest version lives in tests/test_gemm_gpu.py."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _report(name, got, ref, extra=None):
    import torch

    got = got.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    rec = {"case": name, "max_err": err.max().item(), "ref_max": scale, "rel": err.max().item() / scale,
           "nan": bool(torch.isnan(got).any().item())}
    if rec["rel"] > 2e-2 or rec["nan"]:
        # error map over 32-row x 64-col blocks (first 8x8 blocks)
        R, C = err.shape[-2], err.shape[-1]
        e2 = err.reshape(-1, R, C)[0]
        blocks = []
        for i in range(0, min(R, 256), 32):
            blocks.append([round(e2[i:i + 32, j:j + 64].max().item() / scale, 3) for j in range(0, min(C, 512), 64)])
        rec["blockmap_32x64"] = blocks
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)
    return rec["rel"] <= 2e-2 and not rec["nan"]


def _mk(shape, scale=1.0, seed=0):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


@case
def nt_basic():
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (M, N, Kd) in [(128, 256, 64), (256, 256, 128), (384, 512, 768), (1000, 768, 512), (2048, 3072, 768), (130, 72, 200)]:
        x, w = _mk((M, Kd), seed=1), _mk((N, Kd), 0.05, seed=2)
        y = K.linear_fwd(x, w)
        torch.cuda.synchronize()
        ok &= _report(f"nt {M}x{N}x{Kd}", y, x.float() @ w.float().t())
    return ok


@case
def nt_epilogue():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    M, N, Kd = 520, 768, 512
    x, w = _mk((M, Kd), seed=1), _mk((N, Kd), 0.05, seed=2)
    b = torch.randn(N, device="cuda")
    r = _mk((M, N), seed=3)
    ok = True
    y, pre = K.linear_fwd(x, w, bias=b, act=K.ACT_GELU, residual=r, want_pre=True)
    ref_pre = x.float() @ w.float().t() + b
    ok &= _report("nt pre", pre, ref_pre)
    ok &= _report("nt gelu+res", y, F.gelu(ref_pre) + r.float())
    y = K.linear_fwd(x, w, bias=b, act=K.ACT_RELU, out_f32=True, alpha=0.5)
    ok &= _report("nt relu f32 alpha", y, F.relu(0.5 * (x.float() @ w.float().t()) + b))
    # GELU with the derivative as auxiliary output (ACT_GELU_G), TMA-store path (M, N large) and ragged path
    for (M2, N2) in [(520, 768), (130, 72)]:
        x2, w2 = _mk((M2, Kd), seed=5), _mk((N2, Kd), 0.05, seed=6)
        b2 = torch.randn(N2, device="cuda")
        y, gp = K.linear_fwd(x2, w2, bias=b2, act=K.ACT_GELU_G, want_pre=True)
        pr = (x2.float() @ w2.float().t() + b2).requires_grad_(True)
        yr = F.gelu(pr)
        yr.sum().backward()
        ok &= _report(f"nt gelu_g y {M2}x{N2}", y, yr.detach())
        ok &= _report(f"nt gelu_g dgelu {M2}x{N2}", gp, pr.grad)
    return ok


@case
def nn_basic():
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (M, N, Kd) in [(128, 64, 256), (256, 128, 256), (1000, 768, 512), (2048, 768, 3072), (130, 72, 200)]:
        dy, w = _mk((M, N), seed=1), _mk((N, Kd), 0.05, seed=2)
        dx = K.linear_dgrad(dy, w)
        torch.cuda.synchronize()
        ok &= _report(f"nn {M}x{N}->{Kd}", dx, dy.float() @ w.float())
    # fused dgelu
    M, N, Kd = 300, 256, 512
    dy, w, pre = _mk((M, N), seed=1), _mk((N, Kd), 0.05, seed=2), _mk((M, Kd), seed=4)
    dx = K.linear_dgrad(dy, w, act=K.ACT_DGELU, aux_in=pre)
    p = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(p).backward(dy.float() @ w.float())
    ok &= _report("nn dgelu", dx, p.grad)
    for (M2, N2, K2) in [(300, 256, 512), (130, 72, 200)]:
        dy, w, aux = _mk((M2, N2), seed=1), _mk((N2, K2), 0.05, seed=2), _mk((M2, K2), seed=4)
        dx = K.linear_dgrad(dy, w, act=K.ACT_MULAUX, aux_in=aux)
        ok &= _report(f"nn mulaux {M2}x{N2}->{K2}", dx, (dy.float() @ w.float()) * aux.float())
    return ok


@case
def tn_basic():
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (M, N, Kd) in [(64, 128, 256), (256, 128, 256), (1000, 768, 512), (4096, 768, 3072), (130, 72, 200)]:
        dy, x = _mk((M, N), seed=1), _mk((M, Kd), seed=2)
        dw = K.linear_wgrad(dy, x)
        torch.cuda.synchronize()
        ok &= _report(f"tn {M}: {N}x{Kd}", dw, dy.float().t() @ x.float())
    return ok


@case
def conv_all():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    ok = True
    for (B, T, C, N, k) in [(2, 399, 128, 128, 3), (3, 200, 128, 256, 2), (2, 1001, 512, 512, 3), (2, 499, 512, 512, 2)]:
        x = K.alloc_act(B, T, C, "cuda")
        x.copy_(_mk((B, T, C), seed=1))
        w = torch.randn(N, C, k, device="cuda") * 0.03
        wp = K.pack_conv_weight(w)
        wq = wp.float().view(N, k, C).permute(0, 2, 1).contiguous()  # bf16-rounded weights in torch layout
        xr = x.float().transpose(1, 2).requires_grad_(True)
        wr = wq.clone().requires_grad_(True)
        pre_ref = F.conv1d(xr, wr, stride=2)
        y_ref = F.gelu(pre_ref)
        y, pre = K.conv_s2_fwd(x, wp, k, act=K.ACT_GELU, want_pre=True)
        torch.cuda.synchronize()
        ok &= _report(f"conv fwd pre B{B} T{T} C{C} N{N} k{k}", pre, pre_ref.transpose(1, 2))
        ok &= _report(f"conv fwd gelu", y, y_ref.transpose(1, 2))
        dy = K.alloc_act(B, y.shape[1], N, "cuda")
        dy.copy_(_mk(tuple(y.shape), seed=5))
        pre_ref.backward(dy.float().transpose(1, 2))
        dx = K.conv_s2_dgrad(dy, wp, k, T)
        torch.cuda.synchronize()
        ok &= _report(f"conv dgrad", dx, xr.grad.transpose(1, 2))
        dw = K.conv_s2_wgrad(dy, x, k)
        torch.cuda.synchronize()
        ok &= _report(f"conv wgrad", K.unpack_conv_wgrad(dw, C, k).reshape(N, -1), wr.grad.reshape(N, -1))
    return ok


@case
def perf():
    import torch
    from speechmix_b200 import kernels as K

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    shapes = [("ffn1", 23968, 3072, 768), ("ffn2", 23968, 768, 3072), ("qkv", 23968, 2304, 768),
              ("proj", 23968, 768, 768), ("8k", 8192, 8192, 8192)]
    for name, M, N, Kd in shapes:
        x, w = _mk((M, Kd), seed=1), _mk((N, Kd), 0.05, seed=2)
        dy = _mk((M, N), seed=3)
        b = torch.randn(N, device="cuda")
        pre = _mk((M, Kd), seed=4)
        r = _mk((M, N), seed=5)
        fl = 2.0 * M * N * Kd
        variants = [("nt", lambda: K.linear_fwd(x, w)),
                    ("nt+bias+gelu+pre", lambda: K.linear_fwd(x, w, bias=b, act=K.ACT_GELU, want_pre=True)),
                    ("nt+bias+gelu_g (step's dominant)", lambda: K.linear_fwd(x, w, bias=b, act=K.ACT_GELU_G, want_pre=True)),
                    ("nn+mulaux", lambda: K.linear_dgrad(dy, w, act=K.ACT_MULAUX, aux_in=pre)),
                    ("nt+bias+res", lambda: K.linear_fwd(x, w, bias=b, residual=r)),
                    ("nn", lambda: K.linear_dgrad(dy, w)),
                    ("nn+dgelu", lambda: K.linear_dgrad(dy, w, act=K.ACT_DGELU, aux_in=pre)),
                    ("tn", lambda: K.linear_wgrad(dy, x))]
        for nm, fn in variants:
            ms = timeit(fn)
            print(json.dumps({"perf": name, "mode": nm, "M": M, "N": N, "K": Kd, "ms": round(ms, 4),
                              "tflops": round(fl / ms / 1e9, 1), "ew": os.environ.get("SMX_GEMM_EW", "16")}), flush=True)
    # conv1 of the feature encoder at the bench shape
    B, T, C = 32, 47999, 512
    xa = K.alloc_act(B, T, C, "cuda")
    xa.normal_()
    wp = K.pack_conv_weight(torch.randn(512, 512, 3, device="cuda") * 0.03)
    y, pre = K.conv_s2_fwd(xa, wp, 3, act=K.ACT_GELU, want_pre=True)
    dya = K.alloc_act(B, y.shape[1], 512, "cuda")
    dya.normal_()
    fl = 2.0 * B * y.shape[1] * 512 * 1536
    for nm, fn in [("conv1 fwd+gelu+pre", lambda: K.conv_s2_fwd(xa, wp, 3, act=K.ACT_GELU, want_pre=True)),
                   ("conv1 dgrad", lambda: K.conv_s2_dgrad(dya, wp, 3, T)),
                   ("conv1 wgrad", lambda: K.conv_s2_wgrad(dya, xa, 3))]:
        ms = timeit(fn, 5)
        print(json.dumps({"perf": nm, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
    return True


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        import torch  # noqa

        ok = CASES[sys.argv[2]]()
        print(json.dumps({"case_done": sys.argv[2], "ok": bool(ok)}), flush=True)
        sys.exit(0 if ok else 1)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=300)
            out, rc = r.stdout + "\n" + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() + "\nTIMEOUT", -9
        print(f"===== {n} rc={rc} ({time.time() - t:.1f}s)\n{out}", flush=True)


if __name__ == "__main__":
    main()
