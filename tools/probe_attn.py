"""GPU probe for the attention kernels (development tool; pytest version in tests/)."""
import json
import math
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


def _rep(name, got, ref):
    import torch
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    rec = {"case": name, "max_err": err, "ref_max": scale, "rel": err / scale, "nan": bool(torch.isnan(got).any())}
    print(json.dumps(rec), flush=True)
    return rec["rel"] < 2e-2 and not rec["nan"]


def _ref(q, k, v, heads, causal, scale):
    import torch
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    qh = q.float().view(B, Tq, heads, 64).transpose(1, 2)
    kh = k.float().view(B, Tk, heads, 64).transpose(1, 2)
    vh = v.float().view(B, Tk, heads, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        mask = torch.ones(Tq, Tk, device=q.device, dtype=torch.bool).tril(Tk - Tq)
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, -1)
    o = (p @ vh).transpose(1, 2).reshape(B, Tq, heads * 64)
    return o, torch.logsumexp(s, -1)


def _run(B, Tq, Tk, heads, causal, fused=False, bwd=True):
    import torch
    from speechmix_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(0)
    scale = 1 / 8.0
    if fused and Tq == Tk:
        qkv = torch.randn(B, Tq, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16)
        q, k, v = qkv[..., :heads * 64], qkv[..., heads * 64:2 * heads * 64], qkv[..., 2 * heads * 64:]
    else:
        q = torch.randn(B, Tq, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
        k = torch.randn(B, Tk, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
        v = torch.randn(B, Tk, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    name = f"B{B} Tq{Tq} Tk{Tk} H{heads} causal{int(causal)} fused{int(fused)}"
    o, lse = K.attn_fwd(q, k, v, heads, causal=causal, scale=scale)
    torch.cuda.synchronize()
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, lse_ref = _ref(qr, kr, vr, heads, causal, scale)
    ok = _rep("fwd o " + name, o, o_ref.detach())
    ok &= _rep("fwd lse " + name, lse, lse_ref.detach())
    if bwd:
        do = torch.randn(B, Tq, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
        o_ref.backward(do.float())
        dq, dk, dv = K.attn_bwd(do, q, k, v, o, lse, heads, causal=causal, scale=scale)
        torch.cuda.synchronize()
        ok &= _rep("bwd dq " + name, dq, qr.grad)
        ok &= _rep("bwd dk " + name, dk, kr.grad)
        ok &= _rep("bwd dv " + name, dv, vr.grad)
    return ok


@case
def fwd_small():
    ok = True
    for args in [(1, 128, 128, 1, False), (2, 100, 100, 2, False), (2, 256, 256, 2, False), (1, 300, 200, 4, False),
                 (2, 64, 64, 4, True), (2, 200, 200, 2, True), (2, 64, 374, 4, False)]:
        ok &= _run(*args, bwd=False)
    return ok


@case
def bwd_small():
    ok = True
    for args in [(1, 128, 128, 1, False), (2, 100, 100, 2, False), (2, 256, 256, 2, False), (1, 300, 200, 4, False),
                 (2, 64, 64, 4, True), (2, 200, 200, 2, True), (2, 64, 374, 4, False)]:
        ok &= _run(*args, bwd=True)
    return ok


@case
def bias_dbias():
    """additive [H, Tq, Tk] bias with scale 1 (T5) incl. its gradient, self (causal / not) and ragged shapes"""
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (B, Tq, Tk, H, causal) in [(2, 93, 93, 4, False), (2, 64, 64, 4, True), (3, 200, 200, 2, False), (2, 130, 130, 2, True)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        q, k, v, do = (torch.randn(B, t, H * 64, device="cuda", generator=g).mul(0.35).to(torch.bfloat16)
                       for t in (Tq, Tk, Tk, Tq))
        bias = torch.randn(H, Tq, Tk, device="cuda", generator=g)
        name = f"bias B{B} Tq{Tq} Tk{Tk} H{H} causal{int(causal)}"
        o, lse = K.attn_fwd(q, k, v, H, causal=causal, scale=1.0, bias=bias)
        dbias = torch.zeros_like(bias)
        dq, dk, dv = K.attn_bwd(do, q, k, v, o, lse, H, causal=causal, scale=1.0, bias=bias, dbias=dbias)
        qr, kr, vr, br = (t.float().clone().requires_grad_(True) for t in (q, k, v, bias))
        qh, kh, vh = (t.view(B, -1, H, 64).transpose(1, 2) for t in (qr, kr, vr))
        sc = qh @ kh.transpose(-1, -2) + br[None]
        if causal:
            mask = torch.ones(Tq, Tk, device="cuda", dtype=torch.bool).tril(Tk - Tq)
            sc = sc.masked_fill(~mask, float("-inf"))
        o_ref = (torch.softmax(sc, -1) @ vh).transpose(1, 2).reshape(B, Tq, H * 64)
        o_ref.backward(do.float())
        ok &= _rep("fwd o " + name, o, o_ref.detach())
        ok &= _rep("bwd dq " + name, dq, qr.grad)
        ok &= _rep("bwd dk " + name, dk, kr.grad)
        ok &= _rep("bwd dv " + name, dv, vr.grad)
        ok &= _rep("bwd dbias " + name, dbias, br.grad)
    return ok


@case
def kv_len_mask():
    """per-sample key counts (key-padding mask of a padded batch): fused-QKV self-attention and cross-attention,
    with and without an additive bias; k / v rows past the count are zeroed with mask_rows as the contract asks;
    reference = masked softmax in fp32; dk / dv must come back exactly zero on the masked rows."""
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (B, Tq, Tk, H, lens, with_bias) in [(3, 200, 200, 2, [200, 131, 7], False), (2, 749, 749, 4, [749, 300], False),
                                            (4, 64, 374, 4, [374, 1, 128, 129], False), (2, 130, 130, 2, [64, 130], True),
                                            (2, 256, 256, 1, [128, 256], False)]:
        g = torch.Generator(device="cuda").manual_seed(2)
        if Tq == Tk:
            qkv = torch.randn(B, Tq, 3 * H * 64, device="cuda", generator=g).mul(0.5).to(torch.bfloat16)
            q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
            kl = torch.tensor(lens, device="cuda", dtype=torch.int32)
            K.mask_rows(qkv, kl, col_begin=H * 64, col_count=2 * H * 64)
        else:
            q = torch.randn(B, Tq, H * 64, device="cuda", generator=g).mul(0.5).to(torch.bfloat16)
            kv = torch.randn(B, Tk, 2 * H * 64, device="cuda", generator=g).mul(0.5).to(torch.bfloat16)
            k, v = kv[..., :H * 64], kv[..., H * 64:]
            kl = torch.tensor(lens, device="cuda", dtype=torch.int32)
            K.mask_rows(kv, kl)
        keep = torch.arange(Tk, device="cuda")[None, :] < kl[:, None]            # [B, Tk]
        ok &= bool((k.float().abs().sum(-1)[~keep] == 0).all()) and bool((k.float().abs().sum(-1)[keep] > 0).all())
        do = torch.randn(B, Tq, H * 64, device="cuda", generator=g).to(torch.bfloat16)
        bias = torch.randn(H, Tq, Tk, device="cuda", generator=g) if with_bias else None
        scale = 1.0 if with_bias else 0.125
        name = f"kv_len B{B} Tq{Tq} Tk{Tk} H{H} lens{lens} bias{int(with_bias)}"
        o, lse = K.attn_fwd(q, k, v, H, scale=scale, bias=bias, kv_len=kl)
        dq, dk, dv = K.attn_bwd(do, q, k, v, o, lse, H, scale=scale, bias=bias, kv_len=kl)
        qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
        qh, kh, vh = (t.view(B, -1, H, 64).transpose(1, 2) for t in (qr, kr, vr))
        sc = qh @ kh.transpose(-1, -2) * scale
        if bias is not None:
            sc = sc + bias[None]
        sc = sc.masked_fill(~keep[:, None, None, :], float("-inf"))
        o_ref = (torch.softmax(sc, -1) @ vh).transpose(1, 2).reshape(B, Tq, H * 64)
        o_ref.backward(do.float())
        ok &= _rep("fwd o " + name, o, o_ref.detach())
        ok &= _rep("fwd lse " + name, lse, torch.logsumexp(sc, -1).detach())
        ok &= _rep("bwd dq " + name, dq, qr.grad)
        ok &= _rep("bwd dk " + name, dk, kr.grad)
        ok &= _rep("bwd dv " + name, dv, vr.grad)
        zero_ok = bool((dk.float().abs().sum(-1)[~keep] == 0).all()) and bool((dv.float().abs().sum(-1)[~keep] == 0).all())
        print(json.dumps({"case": "masked dk/dv rows exactly zero " + name, "ok": zero_ok}), flush=True)
        ok &= zero_ok
    return ok


@case
def fwd2_rescale():
    """two-query-tile kernel (attention_fwd2.cu): odd tile counts, ragged tails, kv_len cuts, and inputs whose row
    maxima keep growing along the keys so that the LAZY output rescale (threshold 2^8) fires on many tiles."""
    import torch
    from speechmix_b200 import kernels as K
    ok = True
    for (B, T, H, qs, ramp) in [(2, 749, 2, 1.0, 0.0), (2, 374, 3, 1.0, 0.0), (1, 129, 1, 1.0, 0.0), (2, 640, 2, 6.0, 3.0),
                                (1, 1499, 2, 8.0, 5.0), (3, 257, 1, 4.0, 8.0), (2, 1000, 2, 0.2, 0.0)]:
        g = torch.Generator(device="cuda").manual_seed(3)
        qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g)
        qkv[..., :H * 64] *= qs
        qkv[..., H * 64:2 * H * 64] *= (1.0 + ramp * torch.arange(T, device="cuda") / T)[None, :, None]
        qkv = qkv.to(torch.bfloat16)
        q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
        o, lse = K.attn_fwd(q, k, v, H, scale=0.125)
        o_ref, lse_ref = _ref(q, k, v, H, False, 0.125)
        name = f"fwd2 B{B} T{T} H{H} qscale{qs} ramp{ramp}"
        ok &= _rep("fwd o " + name, o, o_ref)
        ok &= _rep("fwd lse " + name, lse, lse_ref)
    return ok


@case
def full_size():
    ok = _run(4, 749, 749, 12, False, fused=True)
    ok &= _run(2, 1499, 1499, 16, False, fused=True)
    return ok


@case
def perf():
    import torch
    from speechmix_b200 import kernels as K
    B, T, H = 32, 749, 12
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    do = torch.randn(B, T, H * 64, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = K.attn_fwd(q, k, v, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fn, nm, mult in ((lambda: K.attn_fwd(q, k, v, H), "fwd", 4), (lambda: K.attn_bwd(do, q, k, v, o, lse, H), "bwd", 10)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"perf": nm, "ms": ms, "tflops_alg": mult * B * H * T * T * 64 / ms / 1e9,
                          "env": {k_: v_ for k_, v_ in os.environ.items() if k_.startswith("SMX_")}}), flush=True)
    # same-box library baseline: torch SDPA (bf16, [B, H, T, 64] contiguous) forward and forward+backward
    import torch.nn.functional as F
    qh, kh, vh = (t.view(B, T, H, 64).transpose(1, 2).contiguous().requires_grad_(True) for t in (q, k, v))
    doh = do.view(B, T, H, 64).transpose(1, 2).contiguous()
    for nm in ("sdpa fwd", "sdpa fwd+bwd"):
        def fn():
            out = F.scaled_dot_product_attention(qh, kh, vh, scale=0.125)
            if nm.endswith("bwd"):
                out.backward(doh)
        if nm == "sdpa fwd":
            ctx = torch.no_grad()
        else:
            ctx = torch.enable_grad()
        with ctx:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"perf": nm, "ms": ms, "tflops_alg": (4 if nm == "sdpa fwd" else 14) * B * H * T * T * 64 / ms / 1e9}), flush=True)
    return True


@case
def trace():
    """SM-clock timeline of CTA (0,0,0) of the two-query-tile forward kernel at the bench shape: per key tile, when each
    softmax group sees its scores / has them in registers / finishes the exponentials / publishes P, and when the MMA
    thread sees P_i and has issued P_i.V + the next S_i."""
    import torch
    from speechmix_b200 import _lib, kernels as K
    B, T, H = 32, 749, 12
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    for _ in range(2):
        K.attn_fwd(q, k, v, H)
    buf = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
    lib = _lib.load()
    lib.smx_debug_attn_trace(buf.data_ptr())
    K.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    lib.smx_debug_attn_trace(None)
    t = buf.cpu().view(3, 64)
    t0 = int(t[t > 0].min())
    names = ["mma", "sm0", "sm1"]
    for r in range(3):
        row = [int(x) - t0 for x in t[r].tolist() if x > 0]
        print(json.dumps({"trace": names[r], "clk": row}), flush=True)
    return True


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
