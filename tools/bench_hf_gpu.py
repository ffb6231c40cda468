"""Same-box LIBRARY baseline: what the reference's own code path
(ref:speechmix/hf_model.py:397,357 -> transformers -> cuBLAS / cuDNN / SDPA) does on the B200 this repo is measured on.

  1. whole step: the oracle restatement of HFSpeechMixEED (oracle/hf_oracle.py -- the reference glue over the same
     transformers classes) on the GPU under torch.autocast(bf16) with fused AdamW, at the bench workload
     (wav2vec2-base + bart-base, down_scale 2, 15 s audio, T_dec 64; batch as large as fits, default 32);
  2. per op: cuBLAS bf16 GEMMs (torch.matmul) at the six hot GEMM shapes of the step against smx_gemm, and
     F.scaled_dot_product_attention forward / backward at B32 x H12 x T749 x 64 against the attention kernels.

Development / reporting tool: it imports oracle/ (test infrastructure) and is NOT part of the product or of bench.py's
own arm.  Prints one JSON line per measurement."""
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def _time(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def whole_step(batch=32, steps=5):
    from oracle import hf_oracle as O
    spc, txc = O.speech_config("base"), O.text_config("bart-base")
    s, t = O.build_backbones(spc, txc, seed=0)
    with contextlib.redirect_stdout(sys.stderr):
        model = O.OracleEED(s, t, down_scale=2).cuda().train()
    x, labels = O.synthetic_batch(batch, 15.0, 64, txc.vocab_size, seed=0)
    x, labels = x.cuda(), labels.cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x, labels=labels)
        out["loss"].backward()
        opt.step()

    try:
        ms = _time(step, iters=steps, warmup=3)
    except torch.OutOfMemoryError:
        print(json.dumps({"library_step": "oom", "batch": batch}), flush=True)
        return
    print(json.dumps({"library_step": "transformers %s + torch.autocast(bf16) + SDPA + fused AdamW (eager)" % __import__("transformers").__version__,
                      "workload": "HFSpeechMixEED glue, wav2vec2-base + bart-base ds2, batch %d x 15 s, T_dec 64" % batch,
                      "ms_per_step": ms, "audio_s_per_s": batch * 15.0 / (ms * 1e-3),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


def gemm_ops():
    from speechmix_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(0)
    M = 32 * 749
    shapes = [("ffn_up NT (+bias+GELU in ours)", M, 3072, 768), ("ffn_down NT", M, 768, 3072), ("qkv NT", M, 2304, 768),
              ("out_proj NT", M, 768, 768), ("decoder NT (M=2048)", 2048, 3072, 768), ("lm_head chunk NT", 2048, 8192, 768)]
    for name, m, n, k in shapes:
        x = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        ms_lib = _time(lambda: torch.matmul(x, w.t()))
        ms_ours = _time(lambda: K.linear_fwd(x, w))
        fl = 2.0 * m * n * k
        rec = {"gemm": name, "m": m, "n": n, "k": k, "cublas_ms": ms_lib, "cublas_tflops": fl / ms_lib / 1e9,
               "ours_ms": ms_ours, "ours_tflops": fl / ms_ours / 1e9}
        if name.startswith("ffn_up"):
            b = torch.randn(n, device="cuda", generator=g)
            ms_f = _time(lambda: K.linear_fwd(x, w, bias=b, act=K.ACT_GELU_G, want_pre=True))
            ms_l = _time(lambda: F.gelu(F.linear(x, w, b.to(torch.bfloat16))))
            rec.update({"ours_fused_bias_gelu_gelugrad_ms": ms_f, "ours_fused_tflops": fl / ms_f / 1e9,
                        "cublas_plus_eager_gelu_ms": ms_l})
        print(json.dumps(rec), flush=True)
    # weight gradient (TN) and data gradient (NN) of the FFN up-projection
    m, n, k = M, 3072, 768
    x = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(m, n, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    fl = 2.0 * m * n * k
    for name, lib, ours in (("wgrad TN", lambda: torch.matmul(dy.t(), x), lambda: K.linear_wgrad(dy, x)),
                            ("dgrad NN", lambda: torch.matmul(dy, w), lambda: K.linear_dgrad(dy, w))):
        ms_lib, ms_ours = _time(lib), _time(ours)
        print(json.dumps({"gemm": name, "m": m, "n": n, "k": k, "cublas_ms": ms_lib, "cublas_tflops": fl / ms_lib / 1e9,
                          "ours_ms": ms_ours, "ours_tflops": fl / ms_ours / 1e9}), flush=True)


def attn_ops():
    from speechmix_b200 import kernels as K
    B, T, H = 32, 749, 12
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B, T, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    do = torch.randn(B, T, H * 64, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = K.attn_fwd(q, k, v, H)
    ours_f = _time(lambda: K.attn_fwd(q, k, v, H))
    ours_b = _time(lambda: K.attn_bwd(do, q, k, v, o, lse, H))
    qh, kh, vh = (t.view(B, T, H, 64).transpose(1, 2).contiguous().requires_grad_(True) for t in (q, k, v))
    doh = do.view(B, T, H, 64).transpose(1, 2).contiguous()
    with torch.no_grad():
        lib_f = _time(lambda: F.scaled_dot_product_attention(qh, kh, vh, scale=0.125))

    def fb():
        F.scaled_dot_product_attention(qh, kh, vh, scale=0.125).backward(doh)
    lib_fb = _time(fb)
    fl = B * H * T * T * 64
    print(json.dumps({"attention": "B32 H12 T749 d64 bf16", "ours_fwd_ms": ours_f, "ours_bwd_ms": ours_b,
                      "sdpa_fwd_ms": lib_f, "sdpa_fwd_bwd_ms": lib_fb, "sdpa_bwd_ms_by_difference": lib_fb - lib_f,
                      "ours_fwd_tflops": 4 * fl / ours_f / 1e9, "sdpa_fwd_tflops": 4 * fl / lib_f / 1e9,
                      "ours_bwd_tflops": 10 * fl / ours_b / 1e9, "sdpa_bwd_tflops": 10 * fl / (lib_fb - lib_f) / 1e9}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "step"]
    if "ops" in which:
        gemm_ops()
        attn_ops()
    if "step" in which:
        whole_step(batch=int(os.environ.get("LIB_BATCH", "32")))
