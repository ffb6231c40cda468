"""Per-parameter gradient error of a golden case vs the CPU reference restatement (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
import torch
from tests._cases import build_oracle, load_fixture
from tests.test_model_gpu import _mine_from

name = sys.argv[1] if len(sys.argv) > 1 else "mini_t5"
fx = load_fixture(name)
ora, x, labels = build_oracle(fx)
mine = _mine_from(ora, fx, torch.device("cuda:0"))
ref = ora(x, labels=labels, keep_full_logits=True)
out = mine(x.cuda(), labels=labels.cuda())
print("loss", float(out["loss"]), float(ref["loss"]))
ref["loss"].backward()
out["loss"].backward()
po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
for k, p in po.items():
    if p.grad is None:
        continue
    g = pm[k].grad.cpu()
    n = float(p.grad.norm())
    print("%-75s rel %.4f  |ref| %.4g" % (k, float((g - p.grad).norm()) / (n + 1e-12), n))
