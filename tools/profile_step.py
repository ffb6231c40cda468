"""Per-kernel-family device time of one training step at the bench configuration (CUDA events around
every C-ABI call; eager mode).  Development tool: guides which kernel to optimise next."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
import torch  # noqa: E402

from speechmix_b200 import SpeechMixEED, _lib, parallel, presets as O  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0     # dropout probability at every site (stock checkpoints: 0.1)
    spc, txc = O.speech_config("base"), O.text_config("bart-base")
    if drop > 0:
        for cfg, keys in ((spc, ("hidden_dropout", "attention_dropout", "activation_dropout", "feat_proj_dropout")),
                          (txc, ("dropout", "attention_dropout", "activation_dropout"))):
            for k in keys:
                setattr(cfg, k, drop)
    model = SpeechMixEED(spc, txc, down_scale=2)
    parallel.init_like_reference(model)
    model = model.cuda().train()
    x = torch.randn(B, 240000, device="cuda")
    y = torch.randint(4, txc.vocab_size, (B, 64), device="cuda")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        model(x, labels=y, return_model_detail=False)["loss"].backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    _lib.PROFILE = {}
    step()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    rows = []
    for k, evs in prof.items():
        ms = sum(a.elapsed_time(b) for a, b in evs)
        rows.append((ms, len(evs), k))
    rows.sort(reverse=True)
    tot_k = sum(r[0] for r in rows)
    print(json.dumps({"step_ms_unprofiled": total, "sum_kernel_ms": tot_k, "calls": sum(r[1] for r in rows)}))
    for ms, n, k in rows[:60]:
        print("%8.3f ms  %5.1f%%  x%-4d %s" % (ms, 100 * ms / tot_k, n, k))


if __name__ == "__main__":
    main()
