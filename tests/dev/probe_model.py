"""GPU probe: SpeechMixEED (CUDA kernels) against the CPU oracle on identical weights/inputs."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
import torch  # noqa: E402

from oracle import hf_oracle as O  # noqa: E402


def rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def run(sp_kind, sp_type, tx_kind, kw, B, secs, t_dec, ignore_tail=False, backward=True, cls="eed"):
    import speechmix_b200 as P
    spc, txc = O.speech_config(sp_kind, model_type=sp_type), O.text_config(tx_kind)
    s, t = O.build_backbones(spc, txc, seed=0)
    ocls, mcls = {"eed": (O.OracleEED, P.SpeechMixEED), "adapter": (O.OracleAdapter, P.SpeechMixAdapter),
                  "fixed": (O.OracleFixed, P.SpeechMixFixed)}[cls]
    ora = ocls(s, t, **kw)
    O.reinit_glue(ora, 1)
    ora.train(backward)
    mine = mcls(spc, txc, **kw)
    mine.load_state_dict(ora.state_dict())
    mine = mine.cuda()
    mine.train(backward)
    x, labels = O.synthetic_batch(B, secs, t_dec, txc.vocab_size, seed=0, ignore_tail=ignore_tail)
    t0 = time.time()
    out_o = ora(x, labels=labels, keep_full_logits=True)
    t_cpu = time.time() - t0
    out_m = mine(x.cuda(), labels=labels.cuda())
    torch.cuda.synchronize()
    logits_m = mine.decoder_model.full_logits(out_m["decoder_last_hidden_state"])
    rec = {"case": f"{cls} {sp_kind}/{tx_kind} {kw} B{B} {secs}s",
           "loss_ref": float(out_o["loss"]), "loss": float(out_m["loss"]),
           "loss_abs_err": abs(float(out_o["loss"]) - float(out_m["loss"])),
           "speech_rel": rel(out_m["speech_last_hidden_state"], out_o["speech_last_hidden_state"]),
           "speech_l2": rel_l2(out_m["speech_last_hidden_state"], out_o["speech_last_hidden_state"]),
           "embeds_rel": rel(out_m["inputs_embeds"], out_o["inputs_embeds"]),
           "enc_rel": rel(out_m["encoder_last_hidden_state"], out_o["encoder_last_hidden_state"]),
           "logits_rel": rel(logits_m, out_o["full_logits"]),
           "logits_l2": rel_l2(logits_m, out_o["full_logits"]),
           "argmax_agree": (out_m["logits"].cpu() == out_o["logits"]).float().mean().item(),
           "cpu_s": t_cpu}
    if backward:
        out_o["loss"].backward()
        out_m["loss"].backward()
        torch.cuda.synchronize()
        po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
        worst = []
        for k, p in po.items():
            if p.grad is None:
                continue
            g = pm[k].grad
            if g is None:
                worst.append((9.9, k + " MISSING"))
                continue
            worst.append((rel_l2(g, p.grad), k))
        worst.sort(reverse=True)
        rec["grad_worst"] = [(round(a, 4), b) for a, b in worst[:8]]
        rec["grad_median_l2"] = sorted(a for a, _ in worst)[len(worst) // 2]
        rec["n_grads"] = len(worst)
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    which = sys.argv[1:] or ["mini"]
    if "mini" in which:
        run("mini", "wav2vec2", "bart-mini", dict(down_scale=2), 2, 1.0, 8, ignore_tail=True)
        run("mini", "wav2vec2", "bart-mini", dict(down_scale=8, weighted_sum=True), 2, 1.5, 8)
        run("mini", "wav2vec2", "mbart-mini", dict(down_scale=4, share_layer_ratio=0.5), 3, 1.0, 6)
    if "variants" in which:
        run("mini_large", "hubert", "mbart-mini", dict(down_scale=2), 2, 1.0, 8)
        run("mini_large", "wav2vec2", "bart-mini", dict(down_scale=8), 2, 1.5, 8)
        run("mini", "wav2vec2", "bart-mini", dict(down_scale=2), 2, 1.0, 8, cls="adapter")
        run("mini", "wav2vec2", "mbart-mini", dict(down_scale=2, adapter_indexing="per_layer"), 2, 1.0, 8, cls="adapter")
        run("mini", "wav2vec2", "bart-mini", dict(down_scale=2, fixed_speech=True, fixed_nlp=False), 2, 1.0, 8, cls="fixed")
    if "base" in which:
        run("base", "wav2vec2", "bart-base", dict(down_scale=2), 1, 5.0, 24, backward=False)
        run("base", "wav2vec2", "bart-base", dict(down_scale=2), 2, 3.0, 16, backward=True)
