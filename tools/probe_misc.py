"""GPU probe: conv0+GN+GELU, positional conv, LayerNorm, colsum, embeddings, LM-head CE, weighted sum."""
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


def _rep(name, got, ref, tol=2e-2):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    rec = {"case": name, "max_err": err, "ref_max": scale, "rel": err / scale, "nan": bool(torch.isnan(got).any())}
    print(json.dumps(rec), flush=True)
    return rec["rel"] < tol and not rec["nan"]


@case
def conv0():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    ok = True
    for (B, n, C) in [(2, 16000, 128), (3, 80000, 512)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, n, device="cuda", generator=g)
        w = (torch.randn(C, 1, 10, device="cuda", generator=g) * 0.3).requires_grad_(True)
        gamma = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        beta = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        y, stats, mom, gp = K.conv0_fwd(x, w.detach(), gamma.detach(), beta.detach(), want_gprime=True)
        ref = F.gelu(F.group_norm(F.conv1d(x[:, None], w, stride=5), C, gamma, beta, 1e-5))
        ok &= _rep(f"conv0 fwd B{B} n{n} C{C}", y, ref.transpose(1, 2))
        dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16)
        ref.backward(dy.float().transpose(1, 2))
        dw, dg, db = K.conv0_bwd(x, w.detach(), gamma.detach(), beta.detach(), stats, mom, dy, gp)
        ok &= _rep("conv0 dw", dw, w.grad)
        ok &= _rep("conv0 dgamma", dg, gamma.grad)
        ok &= _rep("conv0 dbeta", db, beta.grad)
    return ok


@case
def posconv():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    ok = True
    for (B, T, H, G) in [(2, 49, 256, 16), (2, 300, 768, 16), (1, 600, 1024, 16)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        ksz = 128
        cg = H // G
        x = torch.randn(B, T, H, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(H, cg, ksz, device="cuda", generator=g) * 0.02)
        w16 = w.to(torch.bfloat16).float().requires_grad_(True)
        bias = (0.1 * torch.randn(H, device="cuda", generator=g)).requires_grad_(True)
        wf, wd = K.posconv_pack(w, G)
        y, pre = K.posconv_fwd(x, wf, bias.detach(), G, ksz, add_input=True)
        xr = x.float().requires_grad_(True)
        pre_ref = F.conv1d(xr.transpose(1, 2), w16, bias, padding=ksz // 2, groups=G)[:, :, :-1].transpose(1, 2)
        y_ref = xr + F.gelu(pre_ref)
        ok &= _rep(f"posconv pre B{B} T{T} H{H}", pre, pre_ref)
        ok &= _rep("posconv y", y, y_ref)
        dpre = (torch.randn(B, T, H, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
        pre_ref.backward(dpre.float())
        dx = K.posconv_dgrad(dpre, wd, G, ksz)
        ok &= _rep("posconv dgrad", dx, xr.grad)
        dw = K.posconv_wgrad(dpre, x, G, ksz)
        ok &= _rep("posconv wgrad", dw, w16.grad)
    return ok


@case
def rowwise():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)
    for (R, C) in [(1000, 768), (333, 512), (100, 1024), (64, 256), (5000, 768)]:
        x = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        r = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        gamma = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        beta = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        y, s, mean, rstd = K.layernorm_fwd(x, gamma.detach(), beta.detach(), res=r, want_sum=True)
        sr = (x.float() + r.float()).to(torch.bfloat16).float().requires_grad_(True)
        y_ref = F.layer_norm(sr, (C,), gamma, beta, 1e-5)
        ok &= _rep(f"ln fwd {R}x{C}", y, y_ref)
        ok &= _rep("ln sum", s, sr)
        dy = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        y_ref.backward(dy.float())
        dx, dg, db, cs = K.layernorm_bwd(dy, s, gamma.detach(), mean, rstd, want_colsum=True)
        ok &= _rep("ln dx", dx, sr.grad)
        ok &= _rep("ln dx colsum", cs, dx.float().sum(0), 5e-3)  # kernel sums the unrounded fp32 dx
        ok &= _rep("ln dgamma", dg, gamma.grad, 5e-3)
        ok &= _rep("ln dbeta", db, beta.grad, 5e-3)
        ok &= _rep("colsum", K.colsum(dy), dy.float().sum(0), 1e-3)
        ok &= _rep("dact", K.dact(dy, x), dy.float() * torch.autograd.functional.jvp(F.gelu, x.float(), torch.ones_like(x.float()))[1])
    # LayerNorm backward with a residual-branch gradient (shared-memory-ring kernel) incl. a ragged width
    for (R, C) in [(777, 768), (300, 1024), (64, 264)]:
        x = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        gamma = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        beta = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
        dy = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        dres = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
        y, _, mean, rstd = K.layernorm_fwd(x, gamma.detach(), beta.detach())
        xr = x.float().requires_grad_(True)
        F.layer_norm(xr, (C,), gamma, beta, 1e-5).backward(dy.float())
        dx, dg, db = K.layernorm_bwd(dy, x, gamma.detach(), mean, rstd, dres=dres)
        ok &= _rep(f"ln bwd+dres dx {R}x{C}", dx, xr.grad + dres.float())
        ok &= _rep("ln bwd+dres dgamma", dg, gamma.grad, 5e-3)
        ok &= _rep("ln bwd+dres dbeta", db, beta.grad, 5e-3)
    # padded-row masking and SpecAugment replacement / gradient
    B, T, C = 3, 50, 256
    x = torch.randn(B, T, C, device="cuda", generator=g).to(torch.bfloat16)
    lens = torch.tensor([50, 17, 1], device="cuda", dtype=torch.int32)
    keep = (torch.arange(T, device="cuda")[None, :] < lens[:, None])
    ok &= _rep("mask_rows", K.mask_rows(x.clone(), lens), x.float() * keep[..., None])
    part = K.mask_rows(x.clone(), lens, col_begin=64, col_count=128).float()
    ref = x.float().clone()
    ref[:, :, 64:192] *= keep[..., None]
    ok &= _rep("mask_rows column range", part, ref)
    tm = (torch.rand(B, T, device="cuda", generator=g) < 0.3).to(torch.uint8)
    fm = (torch.rand(B, C, device="cuda", generator=g) < 0.2).to(torch.uint8)
    emb = torch.randn(C, device="cuda", generator=g)
    for (t_, f_) in [(tm, fm), (tm, None), (None, fm)]:
        y = K.spec_augment_fwd(x, t_, f_, emb)
        ref = x.float().clone()
        if t_ is not None:
            ref[t_.bool()] = emb.to(torch.bfloat16).float()
        if f_ is not None:
            ref = ref * (1 - f_.float())[:, None, :]
        ok &= _rep("spec_augment fwd", y, ref)
        dy = torch.randn(B, T, C, device="cuda", generator=g).to(torch.bfloat16)
        dx, de = K.spec_augment_bwd(dy, t_, f_)
        live = torch.ones(B, T, C, device="cuda")
        if t_ is not None:
            live = live * (1 - t_.float())[..., None]
        if f_ is not None:
            live = live * (1 - f_.float())[:, None, :]
        ok &= _rep("spec_augment dx", dx, dy.float() * live)
        if t_ is not None:
            w = t_.float()[..., None] * (1.0 if f_ is None else (1 - f_.float())[:, None, :])
            ok &= _rep("spec_augment dembed", de, (dy.float() * w).sum((0, 1)), 1e-3)
    # rms
    x = torch.randn(200, 768, device="cuda", generator=g).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(768, device="cuda", generator=g)
    y, _, _, rstd = K.layernorm_fwd(x, gamma, None, eps=1e-6, rms_only=True)
    xr = x.float()
    ok &= _rep("rmsnorm", y, xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * gamma)
    # embeddings
    V, D, B, T = 1000, 256, 3, 17
    tok = torch.randn(V, D, device="cuda", generator=g)
    pos = torch.randn(64, D, device="cuda", generator=g)
    ids = torch.randint(0, V, (B, T), device="cuda", generator=g)
    out = K.embed_fwd(ids, tok, pos, None, B, T, D, scale=2.0, pos_offset=2, device="cuda")
    ok &= _rep("embed fwd", out, tok[ids] * 2.0 + pos[2:2 + T][None])
    dout = torch.randn(B, T, D, device="cuda", generator=g).to(torch.bfloat16)
    dt, dp = torch.zeros_like(tok), torch.zeros_like(pos)
    K.embed_bwd(ids, dout, dt, dp, scale=2.0, pos_offset=2)
    dt_ref = torch.zeros_like(tok).index_add_(0, ids.view(-1), dout.float().view(-1, D) * 2.0)
    ok &= _rep("embed dtok", dt, dt_ref, 1e-3)
    dp_ref = torch.zeros_like(pos)
    dp_ref[2:2 + T] = dout.float().sum(0)
    ok &= _rep("embed dpos", dp, dp_ref, 1e-3)
    # weighted sum
    xs = [torch.randn(4, 50, 256, device="cuda", generator=g).to(torch.bfloat16) for _ in range(5)]
    w = torch.softmax(torch.randn(5, device="cuda", generator=g), 0)
    ok &= _rep("wsum fwd", K.weighted_sum_fwd(xs, w), sum(wi * x.float() for wi, x in zip(w, xs)))
    d = torch.randn(4, 50, 256, device="cuda", generator=g).to(torch.bfloat16)
    ok &= _rep("wsum dw", K.weighted_sum_bwd_w(xs, d), torch.stack([(x.float() * d.float()).sum() for x in xs]), 1e-3)
    return ok


@case
def perf():
    """Row-kernel timings at the bench shape [23968, 768] (L2 flushed between launches by reading 256 MB -- a fill would
    leave dirty lines whose write-back is then charged to the kernel under test)."""
    import torch
    from speechmix_b200 import kernels as K
    R, C = 23968, 768
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
    r = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")   # 256 MB, READ between launches: clean L2 lines
    y, s, mean, rstd = K.layernorm_fwd(x, gamma, beta, res=r, want_sum=True)
    mb = R * C * 2 / 1e6

    def timeit(fn, iters=10, cold=True):
        ts = []
        for _ in range(iters + 2):
            if cold:
                flush.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[2:])[len(ts[2:]) // 2]

    cases = [("ln_fwd x->y", 2 * mb, lambda: K.layernorm_fwd(x, gamma, beta)),
             ("ln_fwd x+res->sum,y", 4 * mb, lambda: K.layernorm_fwd(x, gamma, beta, res=r, want_sum=True)),
             ("ln_bwd dy,x->dx (+colsum)", 3 * mb, lambda: K.layernorm_bwd(dy, s, gamma, mean, rstd, want_colsum=True)),
             ("ln_bwd dy,x,dres->dx", 4 * mb, lambda: K.layernorm_bwd(dy, s, gamma, mean, rstd, dres=r)),
             ("colsum", mb, lambda: K.colsum(dy))]
    for name, mbytes, fn in cases:
        for cold in (True, False):
            ms = timeit(fn, cold=cold)
            print(json.dumps({"perf": name, "cold_l2": cold, "us": round(ms * 1e3, 2), "algorithmic_MB": round(mbytes, 1),
                              "GBps": round(mbytes / ms, 1)}), flush=True)
    return True


@case
def lmhead():
    import torch
    import torch.nn.functional as F
    from speechmix_b200 import kernels as K
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, D, V) in [(100, 256, 1000), (2048, 768, 50265)]:
        h = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
        emb = (torch.randn(V, D, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        bias = torch.randn(V, device="cuda", generator=g) * 0.1
        labels = torch.randint(0, V, (M,), device="cuda", generator=g)
        labels[::7] = -100
        lse, am, row_loss, acc = K.lmhead_ce_fwd(h, emb, bias, labels)
        logits = h.float() @ emb.float().t() + bias
        ok &= _rep(f"lmhead lse M{M} V{V}", lse, torch.logsumexp(logits, -1), 1e-4)
        ref_loss = F.cross_entropy(logits, labels, ignore_index=-100, reduction="sum")
        ok &= _rep("lmhead loss_sum", acc[0:1], ref_loss[None], 1e-4)
        ok &= _rep("lmhead count", acc[1:2], (labels != -100).sum()[None].float(), 1e-6)
        agree = (am == logits.argmax(-1)).float().mean().item()
        print(json.dumps({"case": "lmhead argmax agree", "frac": agree}), flush=True)
        ok &= agree > 0.99
        # dlogits chunk
        coef = ((labels != -100).float() / acc[1]).contiguous()
        buf = torch.empty(M, 8192, device="cuda", dtype=torch.bfloat16)
        v0 = (V - 1) // 8192 * 8192   # last (ragged) chunk
        vn = V - v0
        K.lmhead_dlogits(h, emb, bias, labels, lse, coef, buf, v0, vn)
        lr = logits.clone().requires_grad_(True)
        F.cross_entropy(lr, labels, ignore_index=-100).backward()
        ok &= _rep("lmhead dlogits", buf[:, :vn], lr.grad[:, v0:v0 + vn], 2e-2)
        # dh accumulate + dE
        dh = torch.zeros(M, D, device="cuda")
        K.gemm_nn_acc_f32(buf, vn, emb[v0:v0 + vn], dh, accumulate=True)
        ok &= _rep("lmhead dh chunk", dh, buf[:, :vn].float() @ emb[v0:v0 + vn].float(), 1e-2)
        dE = torch.zeros(V, D, device="cuda")
        K.gemm_tn_into(buf, vn, h, dE[v0:v0 + vn])
        ok &= _rep("lmhead dE chunk", dE[v0:v0 + vn], buf[:, :vn].float().t() @ h.float(), 1e-2)
    return ok


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        ok = CASES[sys.argv[2]]()
        print(json.dumps({"case_done": sys.argv[2], "ok": bool(ok)}), flush=True)
        sys.exit(0 if ok else 1)
    for n in (sys.argv[1:] or list(CASES)):
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=300)
            out, rc = r.stdout + "\n" + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() + "\nTIMEOUT", -9
        print(f"===== {n} rc={rc} ({time.time() - t:.1f}s)\n{out}", flush=True)


if __name__ == "__main__":
    main()
