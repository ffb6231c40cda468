"""Greedy-decode throughput at the bench shapes (wav2vec2-base + bart-base, batch 32 x 15 s): KV-cached decode vs
the notebook-style full-prefix recompute.  Development tool; numbers quoted in DESIGN.md."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
import torch  # noqa: E402

from speechmix_b200 import SpeechMixEED, parallel, presets  # noqa: E402


def main():
    B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 64
    m = SpeechMixEED(presets.speech_config("base"), presets.text_config("bart-base"), down_scale=2)
    parallel.init_like_reference(m)
    m = m.cuda().eval()
    x = torch.randn(B, 240000, device="cuda")
    ref_ids = None
    for name, kw in (("kv-cache + cuda graph", dict(use_cache=True, cuda_graph=True)), ("kv-cache", dict(use_cache=True)),
                     ("recompute", dict(use_cache=False))):
        for _ in range(2):
            ids = m.generate(x, max_length=T, eos_token_id=-1, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = m.generate(x, max_length=T, eos_token_id=-1, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ref_ids = ids if ref_ids is None else ref_ids
        assert torch.equal(ids, ref_ids), name
        print(json.dumps({"generate": name, "batch": B, "tokens": T, "seconds": dt,
                          "tokens_per_s": B * (T - 1) / dt, "audio_s_per_s": B * 15.0 / dt}), flush=True)


if __name__ == "__main__":
    main()
