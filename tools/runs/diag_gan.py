"""Diagnostic print-out of the SpeechMixGAN parity numbers (what tests/test_model_gpu.py::test_gan_matches_oracle asserts)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests._cases import build_oracle, load_fixture  # noqa: E402
from tests.test_model_gpu import _mine_from  # noqa: E402
from speechmix_b200 import SpeechMixGAN  # noqa: E402

for name in ("mini_gan", "mini_gan_mbart"):
    fx = load_fixture(name)
    ora, x, labels = build_oracle(fx)
    mine = _mine_from(ora, fx, "cuda:0", cls=SpeechMixGAN)
    ref = ora(x, labels=labels)
    out = mine(x.cuda(), labels=labels.cuda())
    print(name, "loss", float(out["loss"]), float(ref["loss"]))
    for k in ("vt_enc", "nt_enc", "vt", "nt"):
        print(" ", k, [round(v, 4) for v in out[k + "_logit"].tolist()], [round(v, 4) for v in ref[k + "_logit"].tolist()],
              float(out[k + "_loss"]), float(ref[k + "_loss"]))
    print("  ids flips", int((out["logits"].cpu() != torch.tensor(fx["argmax_ids"])).sum()))
    ref["loss"].backward()
    out["loss"].backward()
    po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
    rows = []
    for k, p in po.items():
        if pm[k].grad is None:
            print("  NO GRAD", k)
            continue
        err = float((pm[k].grad.cpu() - p.grad).norm())
        rows.append((err / (float(p.grad.norm()) + 1e-30), err, float(p.grad.norm()), k))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print("   rel %.4f err %.4g norm %.4g %s" % r)
