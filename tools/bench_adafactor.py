"""Optimizer step on the bench model's parameters (wav2vec2-base + bart-base, 235.6 M fp32): the fused four-launch
Adafactor (speechmix_b200/optim.py) against transformers' eager Adafactor (the reference recipe, ref:train.py:298)
and torch's fused AdamW (what bench.py steps with).  Prints one JSON line per optimizer."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transformers.optimization import Adafactor  # noqa: E402

from speechmix_b200 import SpeechMixEED, presets  # noqa: E402
from speechmix_b200.optim import FusedAdafactor  # noqa: E402

model = SpeechMixEED(presets.speech_config("base"), presets.text_config("bart-base"), down_scale=2).cuda()
params = [p for p in model.parameters() if p.requires_grad]
g = torch.Generator(device="cuda").manual_seed(0)
for p in params:
    p.grad = torch.randn(p.shape, device="cuda", generator=g) * 1e-3
n = sum(p.numel() for p in params)
only = set(sys.argv[1:])
for name, make in [("fused_adafactor", lambda: FusedAdafactor(params, lr=5e-4)),
                   ("transformers_adafactor", lambda: Adafactor(params, lr=5e-4, scale_parameter=False, relative_step=False)),
                   ("torch_fused_adamw", lambda: torch.optim.AdamW(params, lr=5e-4, fused=True))]:
    if only and name not in only:
        continue
    opt = make()
    for _ in range(2):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"optimizer": name, "params": n, "tensors": len(params), "gpu_ms_per_step": round(e0.elapsed_time(e1) / 5, 3),
                      "wall_ms_per_step": round(wall, 3)}), flush=True)
    del opt
