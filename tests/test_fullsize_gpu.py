"""Full-size parity of BASELINE.json configs[1..4] against the CPU oracle (fp32) on identical random-init weights and
synthetic inputs: loss, full-vocabulary logits, encoder states, argmax ids and gradients.

* cfg2 (SpeechMixEED wav2vec2-base + bart-base, ds 2, 15 s) at batch 8 = 512 target tokens: the north-star bound
  |loss - reference| <= 1e-3 is asserted at the headline model size (ref:speechmix/hf_model.py:378-447).
* cfg3 (SpeechMixAdapter hubert-large + bart-large, ds 8; ref:...hf_model.py:465-502), cfg4 (SpeechMixSelf
  wav2vec2-large + t5-base, share_layer_ratio 0.5; ref:...hf_model.py:505-583) and cfg5 (SpeechMixEED hubert-large +
  mbart-large-50, V = 250 054, 30 s audio; ref:...hf_model.py:185-447) at full model size, batch cut so that the fp32
  CPU oracle finishes in tens of seconds.

Every case writes its measured errors to ``gpurun_out/parity_<case>.json`` (also when an assertion fails)."""
import json
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel_max(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _dump(case, rec):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_%s.json" % case), "w") as f:
            json.dump(rec, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _grad_report(ora, mine, names=None):
    po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
    scale = max(float(p.grad.norm()) for p in po.values() if p.grad is not None)
    rep = {}
    for k, p in po.items():
        if names is not None and k not in names:
            continue
        if p.grad is None:
            rep[k] = None if pm[k].grad is None else "unexpected gradient"
            continue
        assert pm[k].grad is not None, k
        g = pm[k].grad.float().cpu()
        rep[k] = {"rel_l2": float((g - p.grad).norm() / (p.grad.norm() + 1e-20)),
                  "abs_over_scale": float((g - p.grad).norm() / scale),
                  "norm": float(p.grad.norm())}
    return rep, scale


def _flips_not_near_ties(ref_logits, got_ids, tol_abs):
    """positions whose argmax differs from the reference AND whose reference logits separate the two candidates by
    more than the error bound: a random-init model's logits are nearly flat, so bf16 noise may pick the runner-up,
    but it may never pick a token the reference scores clearly lower."""
    ref_logits = ref_logits.detach().float().cpu()
    got_ids = got_ids.cpu()
    ref_ids = ref_logits.argmax(-1)
    diff = (got_ids != ref_ids)
    if not bool(diff.any()):
        return 0, 0
    top = ref_logits.gather(-1, ref_ids[..., None])[..., 0]
    alt = ref_logits.gather(-1, got_ids[..., None])[..., 0]
    return int(diff.sum()), int((diff & ((top - alt) > tol_abs)).sum())


def _check_grads(rep, rtol, atol_scale, sens=None, sens_factor=4.0):
    """every tensor: rel-L2 <= rtol, or (absolute error tiny against the largest gradient), or -- when a sensitivity map
    is given -- rel-L2 <= sens_factor x the deviation the REFERENCE ITSELF shows when its weights are rounded to bf16."""
    bad = {}
    for k, v in rep.items():
        if isinstance(v, str):
            bad[k] = v
        elif isinstance(v, dict) and v["rel_l2"] > rtol and v["abs_over_scale"] > atol_scale:
            if sens is not None and v["rel_l2"] <= sens_factor * sens.get(k, 0.0):
                continue
            bad[k] = dict(v, sensitivity=None if sens is None else sens.get(k))
    assert not bad, bad


def _bf16_weight_sensitivity(ora, run):
    """gradient deviation of the fp32 reference when only its WEIGHTS are rounded to bf16 (activations, accumulation
    and the loss stay fp32): the part of the bf16 error no kernel can avoid.  The Adapter variant chains 24
    replace-adapters (no residual, SURVEY 8c caveat A) and amplifies such perturbations by an order of magnitude."""
    base = {k: p.grad.clone() for k, p in ora.named_parameters() if p.grad is not None}
    saved = {k: p.detach().clone() for k, p in ora.named_parameters()}
    with torch.no_grad():
        for p in ora.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    ora.zero_grad(set_to_none=True)
    run()["loss"].backward()
    sens = {k: float((p.grad - base[k]).norm() / (base[k].norm() + 1e-20))
            for k, p in ora.named_parameters() if p.grad is not None and k in base}
    with torch.no_grad():
        for k, p in ora.named_parameters():
            p.copy_(saved[k])
    return sens


def test_cfg2_batch8_loss_within_1e3(cuda_device):
    """BASELINE.json configs[1] model at batch 8 x 15 s, T_dec 64 (512 target tokens), forward + backward."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    spc, txc = O.speech_config("base"), O.text_config("bart-base")
    s, t = O.build_backbones(spc, txc, seed=0)
    ora = O.OracleEED(s, t, down_scale=2).train()
    O.reinit_glue(ora, 1)
    mine = SpeechMixEED(spc, txc, down_scale=2)
    mine.load_state_dict(ora.state_dict())
    mine = mine.to(cuda_device).train()
    x, labels = O.synthetic_batch(8, 15.0, 64, txc.vocab_size, seed=0)
    ref = ora(x, labels=labels, keep_full_logits=True)
    out = mine(x.to(cuda_device), labels=labels.to(cuda_device))
    rec = {"loss": float(out["loss"]), "loss_ref": float(ref["loss"]),
           "dloss": abs(float(out["loss"]) - float(ref["loss"])),
           "logits_rel_max": _rel_max(mine.decoder_model.full_logits(out["decoder_last_hidden_state"]), ref["full_logits"]),
           "speech_rel_max": _rel_max(out["speech_last_hidden_state"], ref["speech_last_hidden_state"]),
           "text_enc_rel_max": _rel_max(out["encoder_last_hidden_state"], ref["encoder_last_hidden_state"]),
           "tokens": int(labels.numel())}
    rec["id_flips"], rec["id_flips_not_ties"] = _flips_not_near_ties(
        ref["full_logits"], out["logits"], 2e-2 * float(ref["full_logits"].abs().max()))
    ref["loss"].backward()
    out["loss"].backward()
    rec["grads"], _ = _grad_report(ora, mine)
    _dump("cfg2_b8", rec)
    assert rec["dloss"] < 1e-3, rec["dloss"]                 # north star: loss within 1e-3 absolute
    assert rec["logits_rel_max"] < 2e-2                      # north star: max relative error 2e-2 on logits
    assert rec["speech_rel_max"] < 4e-2 and rec["text_enc_rel_max"] < 4e-2
    # random-init logits are nearly flat (max - runner-up ~ 1e-3 of the range): bf16 may flip isolated near-ties,
    # never a pair the reference separates by more than the logits tolerance
    assert rec["id_flips"] <= 5 and rec["id_flips_not_ties"] == 0, rec["id_flips"]
    _check_grads(rec["grads"], 6e-2, 2e-4)


CASES = {
    # name: (class, speech kind, speech type, text kind, ctor kwargs, batch, seconds, t_dec)
    # 512 target tokens each: with 128 the bf16 noise of the mean NLL of these 24-layer random-init stacks is ~1e-3 by
    # itself (measured over six runs: |dloss| 1e-4 .. 1.7e-3, one 3.3e-3), which made a 1.5e-3 bound a coin toss
    "cfg3_adapter_hubert_large_bart_large": ("Adapter", "large", "hubert", "bart-large", dict(down_scale=8), 2, 15.0, 256),
    "cfg4_self_w2v2_large_t5_base": ("Self", "large", "wav2vec2", "t5-base", dict(down_scale=8, share_layer_ratio=0.5), 2, 15.0, 256),
    "cfg5_eed_hubert_large_mbart50": ("EED", "large", "hubert", "mbart-large-50", dict(down_scale=8), 1, 30.0, 512),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_baseline_config_full_size_vs_oracle(case, cuda_device):
    from oracle import hf_oracle as O
    import speechmix_b200 as S
    cls, sk, st, tk, kw, batch, secs, tdec = CASES[case]
    spc, txc = O.speech_config(sk, model_type=st), O.text_config(tk)
    if tk == "t5-base":
        txc.decoder_start_token_id = 0      # as in the released t5-base config.json (T5Config() leaves it unset)
    s, t = O.build_backbones(spc, txc, seed=0)
    ora = getattr(O, "Oracle" + cls)(s, t, **kw)
    O.reinit_glue(ora, 1)
    ora.train()
    mine = getattr(S, "SpeechMix" + cls)(spc, txc, **kw)
    mine.load_state_dict(ora.state_dict())
    mine = mine.to(cuda_device).train()
    assert mine.list_grad == ora.list_grad and mine.list_no_grad == ora.list_no_grad
    x, labels = O.synthetic_batch(batch, secs, tdec, txc.vocab_size, seed=0)
    extra_o, extra_m = {}, {}
    if cls == "Self":
        g = torch.Generator().manual_seed(5)
        tid = torch.randint(4, txc.vocab_size, (batch, 48), generator=g)
        extra_o, extra_m = {"text_input_ids": tid}, {"text_input_ids": tid.to(cuda_device)}
    ref = ora(x, labels=labels, keep_full_logits=True, **extra_o)
    out = mine(x.to(cuda_device), labels=labels.to(cuda_device), **extra_m)
    rec = {"loss": float(out["loss"]), "loss_ref": float(ref["loss"]),
           "dloss": abs(float(out["loss"]) - float(ref["loss"])), "tokens": int(labels.numel()),
           "logits_rel_max": _rel_max(mine.decoder_model.full_logits(out["decoder_last_hidden_state"]), ref["full_logits"]),
           "speech_rel_max": _rel_max(out["speech_last_hidden_state"], ref["speech_last_hidden_state"]),
           "text_enc_rel_max": _rel_max(out["encoder_last_hidden_state"], ref["encoder_last_hidden_state"])}
    rec["id_flips"], rec["id_flips_not_ties"] = _flips_not_near_ties(
        ref["full_logits"], out["logits"], 2e-2 * float(ref["full_logits"].abs().max()))
    if cls == "Self":
        for k in ("ce_loss", "kld_loss", "mse_loss"):
            rec[k] = float(out[k])
            rec[k + "_ref"] = float(ref[k])
    ref["loss"].backward()
    out["loss"].backward()
    rec["grads"], _ = _grad_report(ora, mine)
    rec["n_grad_tensors"] = sum(1 for v in rec["grads"].values() if isinstance(v, dict))
    worst = sorted(((v["rel_l2"], k) for k, v in rec["grads"].items() if isinstance(v, dict)), reverse=True)[:8]
    rec["worst_grads"] = worst
    _dump(case, rec)
    # every tensor the ORACLE has a gradient for is compared (Adapter, reference indexing: only adapters[-1] is ever
    # used, the other adapters get no gradient on either side -- checked as None == None in _grad_report)
    n_ref = sum(1 for p in ora.parameters() if p.grad is not None)
    assert rec["n_grad_tensors"] == n_ref >= 6 and n_ref <= len(mine.list_grad)
    assert rec["logits_rel_max"] < 2e-2, rec["logits_rel_max"]
    assert rec["speech_rel_max"] < 4e-2 and rec["text_enc_rel_max"] < 4e-2, rec
    # 512 target tokens: per-token bf16 noise of the NLL averages to a few 1e-4
    if cls == "Self":
        # loss = CE + KL(batchmean) + MSE (ref:speechmix/hf_model.py:551-581): the CE term is the per-token mean the
        # north-star bound speaks about; KL "batchmean" is a SUM over T_dec x V per sample (~50 here) and the MSE a
        # mean over states of O(1): both are held to 1e-3 RELATIVE
        assert abs(rec["ce_loss"] - rec["ce_loss_ref"]) < 1.5e-3, (rec["ce_loss"], rec["ce_loss_ref"])
        assert abs(rec["kld_loss"] - rec["kld_loss_ref"]) < 1e-3 * abs(rec["kld_loss_ref"]), (rec["kld_loss"], rec["kld_loss_ref"])
        assert abs(rec["mse_loss"] - rec["mse_loss_ref"]) < 1e-3 * abs(rec["mse_loss_ref"]), (rec["mse_loss"], rec["mse_loss_ref"])
        assert rec["dloss"] < 1e-3 * abs(rec["loss_ref"]), rec["dloss"]
    else:
        assert rec["dloss"] < 1.5e-3, rec["dloss"]
    # near-ties of the flat random-init logits may flip; a clearly separated pair may not.  The hard rule is
    # id_flips_not_ties == 0; the count of near-tie flips is only a sanity bound (the Adapter stack replaces every layer
    # output by adapter(output) without a residual, which flattens its logits most: 8..12 of 128 flip from run to run)
    assert rec["id_flips"] <= 0.15 * rec["tokens"] and rec["id_flips_not_ties"] == 0, (rec["id_flips"], rec["id_flips_not_ties"])
    sens = None
    if cls == "Adapter":
        sens = _bf16_weight_sensitivity(ora, lambda: ora(x, labels=labels, **extra_o))
        rec["sensitivity_worst"] = sorted(((v, k) for k, v in sens.items()), reverse=True)[:8]
        _dump(case, rec)
    _check_grads(rec["grads"], 8e-2 if "t5" in tk else 6e-2, 3e-4, sens=sens)
