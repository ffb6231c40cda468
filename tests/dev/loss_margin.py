import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); os.environ["TRANSFORMERS_OFFLINE"] = "1"
import torch
from tests._cases import build_oracle, load_fixture
from tests.test_model_gpu import _mine_from
fx = dict(load_fixture("mini_eed_ds2"), batch=8, seconds=2.0, t_dec=64, train_mode=False)
ora, x, labels = build_oracle(fx)
with torch.no_grad():
    ref = float(ora(x, labels=labels)["loss"])
mine = _mine_from(ora, fx, torch.device("cuda:0")).eval()
for i in range(6):
    with torch.no_grad():
        out = float(mine(x.cuda(), labels=labels.cuda())["loss"])
    print("scale-test diff %.2e" % abs(out - ref))
fx = load_fixture("cfg1_base")
ora, x, labels = build_oracle(fx)
mine = _mine_from(ora, fx, torch.device("cuda:0")).eval()
for i in range(4):
    with torch.no_grad():
        out = float(mine(x.cuda(), labels=labels.cuda())["loss"])
    print("cfg1 diff %.2e" % abs(out - fx["loss"]))
