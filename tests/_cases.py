"""Shared helpers: rebuild a golden case through the oracle (CPU, fp32)."""
import json
import os

import torch

from oracle import hf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def build_oracle(fx, cls=None):
    sp_cfg = O.speech_config(fx["speech"], model_type=fx["speech_type"])
    for k, v in fx.get("speech_overrides", {}).items():     # e.g. SpecAugment switched on (mini_specaug)
        setattr(sp_cfg, k, v)
    tx_cfg = O.text_config(fx["text"])
    speech, text = O.build_backbones(sp_cfg, tx_cfg, seed=0)
    model = (cls or getattr(O, "Oracle" + fx.get("cls", "EED")))(speech, text, **fx["kwargs"])
    O.reinit_glue(model, seed=1)
    model.train(fx["train_mode"])
    x, labels = O.synthetic_batch(fx["batch"], fx["seconds"], fx["t_dec"], tx_cfg.vocab_size,
                                  seed=0, ignore_tail=fx["ignore_tail"])
    return model, x, labels


def check_sample(t, rec, atol, rtol=0.0):
    flat = t.detach().reshape(-1).double().cpu()
    got = flat[torch.tensor(rec["idx"])]
    ref = torch.tensor(rec["val"], dtype=torch.float64)
    assert list(t.shape) == rec["shape"]
    err = (got - ref).abs().max().item()
    assert err <= atol + rtol * ref.abs().max().item(), (err, atol, rtol)
    return err
