"""The oracle (oracle/hf_oracle.py) against the golden fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only.

The fixtures were generated bit-exactly; here a small tolerance absorbs
different CPU kernels (MKL/oneDNN thread counts) on other hosts."""
import pytest
import torch

from tests._cases import build_oracle, check_sample, load_fixture

MINI = ["mini_eed_ds2", "mini_eed_ds8_ws", "mini_eed_share", "mini_large_mbart", "mini_t5", "mini_specaug", "mini_prompt",
        "mini_fixed", "mini_fixed_params", "mini_t5_share",
        # HFSpeechMixAdapter / HFSpeechMixSelf: the reference's own hook lambda and cal_loss body, run under the documented
        # transformers-5.x calling-convention shims of make_golden.compat_shims (no reference code edited)
        "mini_adapter", "mini_adapter_large", "mini_self", "mini_self_t5",
        # HFSpeechMixGAN: the reference's own cal_loss body under the same keyword filter as Self
        "mini_gan", "mini_gan_mbart"]


@pytest.mark.parametrize("name", MINI)
def test_oracle_matches_reference_golden(name):
    fx = load_fixture(name)
    model, x, labels = build_oracle(fx)
    assert sum(p.numel() for p in model.parameters()) == fx["n_params"]
    assert len(model.state_dict()) == fx["n_state_keys"]
    assert len(model.list_no_grad) == fx["list_no_grad"]
    assert model.speech_encoder_layer == fx["speech_encoder_layer"]
    assert model.nlp_encoder_layer == fx["nlp_encoder_layer"]
    kw = {}
    if "prompt_ids" in fx:        # ids as the reference's own tokenizer produced them from fx["prompt"]
        kw["decoder_text_prompt_ids"] = torch.tensor(fx["prompt_ids"])
    if "text_input_ids" in fx:    # SpeechMixSelf: the frozen text teacher's input
        kw["text_input_ids"] = torch.tensor(fx["text_input_ids"])
    if "np_seed" in fx:           # SpecAugment spans come from numpy's global RNG (transformers' _compute_mask_indices)
        import numpy as np
        np.random.seed(fx["np_seed"])
    out = model(x, labels=labels, keep_full_logits=True, **kw)
    assert abs(float(out["loss"]) - fx["loss"]) < 2e-5 * max(1.0, abs(fx["loss"]))
    assert out["logits"].tolist() == fx["argmax_ids"]
    for k in ("ce_loss", "kld_loss", "mse_loss", "vt_enc_loss", "nt_enc_loss", "vt_loss", "nt_loss"):
        if k in fx:
            assert abs(float(out[k]) - fx[k]) < 2e-5 * max(1.0, abs(fx[k])), k
    for k in ("vt_enc_logit", "nt_enc_logit", "vt_logit", "nt_logit"):      # SpeechMixGAN discriminator logits
        if k in fx:
            assert torch.allclose(out[k].double(), torch.tensor(fx[k], dtype=torch.float64), rtol=1e-4, atol=1e-4), k
    if fx.get("cls") == "GAN":                                              # update-phase counters after one training pass
        assert (model.update_count, model.keep_update) == (fx["update_count"], fx["keep_update"])
    check_sample(out["full_logits"], fx["logits"], atol=2e-4)
    check_sample(out["speech_last_hidden_state"], fx["speech_last_hidden_state"], atol=2e-4)
    check_sample(out["encoder_last_hidden_state"], fx["encoder_last_hidden_state"], atol=2e-4)
    if "grads" in fx:
        out["loss"].backward()
        params = dict(model.named_parameters())
        for k, rec in fx["grads"].items():
            g = params[k].grad
            assert abs(float(g.double().norm()) - rec["norm"]) <= 1e-3 * rec["norm"] + 1e-7, k
            check_sample(g, rec, atol=1e-5, rtol=1e-3)
        if "n_grads" in fx:      # freezing variants: the same parameters are trainable as in the reference
            assert sum(p.grad is not None for p in params.values()) == fx["n_grads"]
            assert len(model.list_grad) == fx["list_grad"]


def test_oracle_greedy_matches_reference_loop():
    from oracle import hf_oracle as O

    fx = load_fixture("mini_eed_share")
    model, x, _ = build_oracle(fx)
    model.eval()
    ids = O.greedy_full_recompute(model, x, max_length=8, eos_token_id=-1)
    assert ids.tolist() == fx["greedy_ids"]


def test_structural_asserts_of_reference_tests():
    """ref:test/test_hf_model.py:18-57 restated offline on mini backbones."""
    from oracle import hf_oracle as O

    for ratio, kept in [(1, 0), (0.5, 1), (0, 2)]:
        sp, tx = O.build_backbones(O.speech_config("mini"), O.text_config("bart-mini"))
        m = O.OracleEED(sp, tx, share_layer_ratio=ratio, down_scale=8)
        assert m.speech_encoder_layer == kept and m.nlp_encoder_layer == 2
        assert len(m.list_no_grad) == 0
    x, labels = O.synthetic_batch(1, 1.0, 4, 1000)
    for ds in (1, 2, 4, 8):
        sp, tx = O.build_backbones(O.speech_config("mini"), O.text_config("bart-mini"))
        m = O.OracleEED(sp, tx, share_layer_ratio=0.5, down_scale=ds, weighted_sum=True).eval()
        d = m(x, labels=labels)["detail"]
        assert round(d["shape_before_length_adapter"][1] / d["shape_before_enc_dec_projector"][1]) == ds
        assert d["weighted_sum"].shape[0] == 2  # L_kept + 1


@pytest.mark.slow
def test_oracle_cfg1_full_size():
    fx = load_fixture("cfg1_base")
    model, x, labels = build_oracle(fx)
    assert sum(p.numel() for p in model.parameters()) == fx["n_params"]
    with torch.no_grad():
        out = model(x, labels=labels, keep_full_logits=True)
    assert abs(float(out["loss"]) - fx["loss"]) < 5e-5
    assert out["logits"].tolist() == fx["argmax_ids"]
    check_sample(out["full_logits"], fx["logits"], atol=5e-4)


def test_oracle_ed_matches_reference_golden():
    """HFSpeechMixED (ref:speechmix/hf_model.py:82-182): oracle.OracleED against the fixture written by the unmodified
    reference class (tests/golden/make_golden.py::run_ed_case) -- state-dict keys, frozen feature encoder, loss, logits,
    argmax ids and sampled gradients."""
    fx = load_fixture("mini_ed")
    model, x, labels = build_oracle(fx)
    assert list(model.state_dict()) == fx["state_keys"]
    assert [k for k, p in model.named_parameters() if not p.requires_grad] == fx["frozen"]
    assert sum(p.numel() for p in model.parameters()) == fx["n_params"]
    out = model(x, labels=labels)
    assert abs(float(out["loss"]) - fx["loss"]) < 2e-5
    assert out["logits"].argmax(-1).tolist() == fx["argmax_ids"]
    check_sample(out["logits"], fx["logits"], atol=2e-4)
    out["loss"].backward()
    params = dict(model.named_parameters())
    for k, rec in fx["grads"].items():
        g = params[k].grad
        assert abs(float(g.double().norm()) - rec["norm"]) <= 1e-3 * rec["norm"] + 1e-7, k
        check_sample(g, rec, atol=1e-5, rtol=1e-3)
    assert sum(p.grad is not None for p in params.values()) == fx["n_grads"]
