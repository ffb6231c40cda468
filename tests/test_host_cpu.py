"""CPU-side checks of the host logic that surrounds the kernels (no GPU, no compute calls into the library)."""
import torch

from speechmix_b200 import ops
from speechmix_b200.model import handle_decoder_input_none, shift_tokens_right


def test_t5_bucket_table_matches_transformers():
    """ops.t5_bucket_table == hf:models/t5/modeling_t5.py:188-233 for every (query, key) offset it can see."""
    from transformers.models.t5.modeling_t5 import T5Attention
    for (tq, tk, bidir, off) in [(93, 93, True, 0), (64, 64, False, 0), (374, 374, True, 0), (1, 17, False, 16), (5, 300, True, 3)]:
        tab = ops.t5_bucket_table(tq, tk, bidir, 32, 128, "cpu", q_offset=off)
        ctx = torch.arange(tq)[:, None] + off
        mem = torch.arange(tk)[None, :]
        ref = T5Attention._relative_position_bucket(mem - ctx, bidirectional=bidir, num_buckets=32, max_distance=128)
        got = tab[(mem - ctx) + (tq + off - 1)]
        assert torch.equal(got.long(), ref), (tq, tk, bidir, off)


def test_shift_tokens_right_and_start_ids():
    """ref:speechmix/hf_model.py:20-34"""
    labels = torch.tensor([[5, 6, -100, -100], [7, 8, 9, 10]])
    out = shift_tokens_right(labels, pad_token_id=1, decoder_start_token_id=2)
    assert out.tolist() == [[2, 5, 6, 1], [2, 7, 8, 9]]

    class Cfg:
        decoder_start_token_id = 2
    assert handle_decoder_input_none(Cfg, batch=3).tolist() == [[2], [2], [2]]


def test_weight_cache_epochs_and_versions():
    """A cached working copy is reused only inside one epoch and while the master's version / address is unchanged
    (fused optimizers do not bump versions, hence the epoch)."""
    cache = ops.WeightCache()
    p = torch.nn.Parameter(torch.randn(4, 4))
    builds = []

    def build(t):
        builds.append(1)
        return t.detach().clone()
    a = cache.get(p, "k", build)
    assert cache.get(p, "k", build) is a and len(builds) == 1
    with torch.no_grad():
        p.add_(1.0)                      # version bump -> rebuilt
    b = cache.get(p, "k", build)
    assert b is not a and len(builds) == 2
    cache.invalidate()                   # epoch bump -> rebuilt even though the version did not move
    assert cache.get(p, "k", build) is not b and len(builds) == 3


def test_fp32_verification_mode_is_inference_only():
    import pytest
    with pytest.raises(RuntimeError):
        with ops.fp32_verification():
            pass
    with torch.no_grad(), ops.fp32_verification():
        from speechmix_b200 import kernels as K
        assert K.FP32_MODE and K.act_dtype() == torch.float32
    from speechmix_b200 import kernels as K
    assert not K.FP32_MODE


def test_model_classes_expose_reference_surface():
    """ctor bookkeeping of ref:speechmix/hf_model.py:222-302 and the structural asserts of ref:test/test_hf_model.py
    (layer counts for share_layer_ratio, no frozen params by default, weighted-sum length L+1)."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixAdapter, SpeechMixEED, SpeechMixFixed, SpeechMixSelf
    spc, txc = O.speech_config("mini"), O.text_config("bart-mini")
    for ratio, kept in ((0, 2), (0.5, 1), (1, 0)):
        m = SpeechMixEED(spc, txc, share_layer_ratio=ratio, down_scale=2, weighted_sum=True)
        assert m.speech_encoder_layer == kept and m.nlp_encoder_layer == 2
        assert len(m.list_no_grad) == 0
        assert m.weights_sum.shape == (kept + 1,)
    assert len(SpeechMixFixed(spc, txc, down_scale=2, fixed_speech=True, fixed_nlp=True).list_grad) == 4 + 0  # bridge only
    a = SpeechMixAdapter(spc, txc, down_scale=2)
    assert len(a.adapters) == 4 and all(n.startswith("decoder_model.model.") for n in a.list_no_grad)
    s = SpeechMixSelf(spc, O.text_config("t5-mini"), down_scale=2)
    assert all(n.startswith("decoder_model.") for n in s.list_no_grad) and len(s.list_no_grad) > 0


def test_collator_contract():
    """ref:train.py:90-133: audio padded with the value -100, label padding -> -100, shared leading bos cut
    (only for a TRUTHY bos id: the reference tests ``if self.tokenizer.bos_token_id and ...``, so bos id 0 -- BART --
    is never cut; kept as is)."""
    from speechmix_b200.training import DataCollatorWithPadding
    feats = [{"input_values": [0.1, 0.2, 0.3], "labels": [3, 5, 6, 2], "text_input_ids": [3, 9, 2]},
             {"input_values": [0.5], "labels": [3, 7, 2], "text_input_ids": [3, 8, 8, 8, 2]}]
    b = DataCollatorWithPadding(pad_token_id=1, bos_token_id=3)(feats)
    assert torch.allclose(b["input_values"], torch.tensor([[0.1, 0.2, 0.3], [0.5, -100.0, -100.0]]))
    assert b["labels"].tolist() == [[5, 6, 2], [7, 2, -100]]
    assert b["text_input_ids"].tolist() == [[3, 9, 2, 1, 1], [3, 8, 8, 8, 2]]
    b2 = DataCollatorWithPadding(pad_token_id=1, bos_token_id=3)([{"input_values": [0.0], "labels": [4, 5]},
                                                                  {"input_values": [0.0], "labels": [3, 5]}])
    assert b2["labels"].tolist() == [[4, 5], [3, 5]]            # bos is only cut when EVERY row starts with it
    b3 = DataCollatorWithPadding(pad_token_id=1, bos_token_id=0)([{"input_values": [0.0], "labels": [0, 5]}])
    assert b3["labels"].tolist() == [[0, 5]]                    # falsy bos id: reference never cuts
    # extension (off by default): zero padding + attention_mask for forward(..., attention_mask=...)
    b4 = DataCollatorWithPadding(pad_token_id=1, bos_token_id=3, return_attention_mask=True)(feats)
    assert torch.allclose(b4["input_values"], torch.tensor([[0.1, 0.2, 0.3], [0.5, 0.0, 0.0]]))
    assert b4["attention_mask"].tolist() == [[1, 1, 1], [1, 0, 0]]
    assert "attention_mask" not in b


def test_feat_extract_output_lengths_match_transformers():
    """hf:models/wav2vec2/modeling_wav2vec2.py:1005-1018 (frame counts behind the key-padding mask)"""
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    from speechmix_b200.speech import SpeechEncoderModel
    cfg = Wav2Vec2Config()
    n = torch.tensor([400, 401, 799, 16000, 80000, 240000, 123457])
    ref = Wav2Vec2Model._get_feat_extract_output_lengths(type("M", (), {"config": cfg})(), n)
    got = SpeechEncoderModel.feat_extract_output_lengths(type("M", (), {"config": cfg})(), n)
    assert got.tolist() == ref.tolist() and got[3] == 49 and got[5] == 749


def test_freezing_policy_matches_reference_callback():
    """ref:speechmix/module/utility.py:6-34"""
    from speechmix_b200.training import FreezingPolicy
    net = torch.nn.Sequential(*[torch.nn.Linear(2, 2) for _ in range(3)])     # 6 parameters
    net[0].bias.requires_grad = False
    pol = FreezingPolicy(net, freeze_epoch=3)
    names = [n for n, _ in net.named_parameters()]
    pol.on_epoch_begin(1)
    assert [n for n, p in net.named_parameters() if p.requires_grad] == names[-2:]
    pol.on_epoch_begin(2)
    assert [n for n, p in net.named_parameters() if p.requires_grad] == names[-4:]
    pol.on_epoch_begin(3)
    assert [n for n, p in net.named_parameters() if p.requires_grad] == [n for n in names if n != "0.bias"]


def test_cpu_call_fails_loudly():
    """No CPU fallback: calling the model on CPU tensors raises instead of computing something else."""
    import pytest
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    m = SpeechMixEED(O.speech_config("mini"), O.text_config("bart-mini"), down_scale=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 16000), labels=torch.randint(4, 100, (1, 4)))


def test_dropout_sites_are_discovered_from_the_configs():
    """the stock configs of the reference recipe switch dropout on; the model records the live sites (the kernels apply
    them with counter-based masks -- tests/test_dropout_gpu.py)"""
    from transformers import BartConfig, Wav2Vec2Config
    from speechmix_b200.model import _dropout_knobs
    live = _dropout_knobs(Wav2Vec2Config(), BartConfig())
    assert "speech.hidden_dropout=0.1" in live and "text.dropout=0.1" in live
    quiet = Wav2Vec2Config(hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0)
    assert _dropout_knobs(quiet, BartConfig(dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)) == []


def test_adafactor_tile_table_covers_every_element_once():
    """speechmix_b200/optim.py: the 64 x 256 tile table of the fused Adafactor step plus the block-per-slice list for
    small factored slices (host logic, no GPU): every element is owned by exactly one tile OR one small slice."""
    import numpy as np
    from speechmix_b200.optim import TILE_C, TILE_R, factored_dims, is_small_slice, tile_table
    shapes = [(768,), (1,), (1000, 768), (512, 512, 3), (512, 1, 10), (768, 48, 128), (1, 50265), (7, 3, 65, 257), (300,),
              (3, 5000, 3), (2, 128, 128), (2, 129, 128)]
    tiles, slices, small, small_floats = tile_table(shapes)
    assert tiles.dtype == np.int32 and slices.dtype == np.int32 and small.dtype == np.int32
    want_floats = 0
    for i, shape in enumerate(shapes):
        factored, batch, rows, cols = factored_dims(shape)
        n = int(np.prod(shape))
        cover = np.zeros(n, np.int64)
        for (_, b, r0, c0) in tiles[tiles[:, 0] == i]:
            if factored:
                rr = np.arange(r0, min(r0 + TILE_R, rows))
                cc = np.arange(c0, min(c0 + TILE_C, cols))
                idx = ((b * rows + rr)[:, None] * cols + cc[None, :]).ravel()
            else:
                idx = np.arange(r0 * TILE_C, min((r0 + TILE_R) * TILE_C, n))
            cover[idx] += 1
        for (_, b) in small[small[:, 0] == i]:
            cover[b * rows * cols:(b + 1) * rows * cols] += 1
        assert (cover == 1).all(), shape
        is_small = factored and is_small_slice(rows, cols)
        assert (slices[:, 0] == i).sum() == (batch if factored and not is_small else 0)
        assert (small[:, 0] == i).sum() == (batch if is_small else 0)
        if is_small:
            want_floats = max(want_floats, rows * cols + rows + cols)
    assert small_floats == want_floats
    assert is_small_slice(512, 3) and is_small_slice(48, 128) and is_small_slice(1, 10) and is_small_slice(128, 128)
    assert not is_small_slice(129, 128) and not is_small_slice(5000, 3) and not is_small_slice(768, 768)
    assert factored_dims((4, 5, 6)) == (True, 4, 5, 6) and factored_dims((9,)) == (False, 1, 1, 9)


def test_fused_adafactor_rejects_unsupported_modes():
    import pytest
    import torch
    from speechmix_b200.optim import FusedAdafactor
    w = torch.nn.Parameter(torch.zeros(4, 4))
    for kw in ({"lr": None}, {"lr": 1e-3, "relative_step": True}, {"lr": 1e-3, "scale_parameter": True},
               {"lr": 1e-3, "beta1": 0.9}):
        with pytest.raises(NotImplementedError):
            FusedAdafactor([w], **kw)
    opt = FusedAdafactor([w], lr=1e-3)
    w.grad = torch.ones(4, 4)
    with pytest.raises(RuntimeError):      # CPU parameter: no fallback
        opt.step()


def test_speechmix_alias_package_and_composite_config(tmp_path):
    """`import speechmix` call sites (ref:train.py:15, ref:eval.py) resolve to this implementation, and
    SpeechMixConfig keeps the reference's behaviour (ref:speechmix/hf_model.py:37-79): a PretrainedConfig built from
    encoder / decoder dicts, ``from_configs`` on two checkpoint directories marks the decoder as a cross-attending
    decoder, ``to_dict`` nests both sub-configs."""
    import speechmix
    import speechmix_b200
    from transformers import BartConfig, PretrainedConfig, Wav2Vec2Config
    assert speechmix.SpeechMixEED is speechmix_b200.SpeechMixEED and speechmix.HFSpeechMixSelf is speechmix_b200.SpeechMixSelf
    enc_dir, dec_dir = tmp_path / "w2v2", tmp_path / "bart"
    Wav2Vec2Config(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512).save_pretrained(enc_dir)
    BartConfig(d_model=256, encoder_layers=2, decoder_layers=2, encoder_attention_heads=4, decoder_attention_heads=4,
               encoder_ffn_dim=512, decoder_ffn_dim=512, vocab_size=1000).save_pretrained(dec_dir)
    cfg = speechmix.SpeechMixConfig.from_configs(str(enc_dir), str(dec_dir))
    assert isinstance(cfg, PretrainedConfig) and cfg.model_type == "speechmix" and cfg.is_encoder_decoder
    assert cfg.encoder.model_type == "wav2vec2" and cfg.encoder.hidden_size == 256
    assert cfg.decoder.model_type == "bart" and cfg.decoder.is_decoder and cfg.decoder.add_cross_attention
    d = cfg.to_dict()
    assert d["model_type"] == "speechmix" and d["encoder"]["hidden_size"] == 256 and d["decoder"]["d_model"] == 256
    again = speechmix.SpeechMixConfig(encoder=d["encoder"], decoder=d["decoder"])     # the reference's dict form
    assert again.decoder.vocab_size == 1000 and again.decoder_start_token_id == cfg.decoder.decoder_start_token_id


def test_checkpoint_directories_and_legacy_keys_load(tmp_path):
    """SURVEY 8f row 2 (ref:eval.py:10, ref:speechmix/hf_model.py:206-220): the constructor takes transformers
    checkpoint DIRECTORIES (safetensors or pytorch_model.bin); files from the weight_g / weight_v era of the positional
    conv and checkpoints that store the tied embedding once load too; a full-model ``pytorch_model.bin`` written by the
    reference loads with ``load_state_dict(torch.load(...))``; an unknown name fails with a clear error."""
    import torch
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    from speechmix_b200.speech import speech_from_pretrained
    spc, txc = O.speech_config("mini"), O.text_config("bart-mini")
    speech, text = O.build_backbones(spc, txc, seed=0)
    sp_dir, tx_dir = tmp_path / "wav2vec2-mini", tmp_path / "bart-mini"
    speech.save_pretrained(sp_dir)                       # model.safetensors
    text.save_pretrained(tx_dir)
    m = SpeechMixEED(str(sp_dir), str(tx_dir), down_scale=2)
    ora = O.OracleEED(speech, text, down_scale=2)
    sd_m, sd_o = m.state_dict(), ora.state_dict()
    assert list(sd_m) == list(sd_o)
    for k in sd_o:
        if not k.startswith(O.GLUE_PREFIXES):            # glue parameters are freshly initialised on both sides
            assert torch.equal(sd_m[k], sd_o[k]), k
    # legacy file: pytorch_model.bin with weight_g / weight_v and a "wav2vec2." prefix
    legacy = {}
    for k, v in speech.state_dict().items():
        k = k.replace("parametrizations.weight.original0", "weight_g").replace("parametrizations.weight.original1", "weight_v")
        legacy["wav2vec2." + k] = v
    leg_dir = tmp_path / "wav2vec2-legacy"
    leg_dir.mkdir()
    spc.save_pretrained(leg_dir)
    torch.save(legacy, leg_dir / "pytorch_model.bin")
    sp2 = speech_from_pretrained(str(leg_dir))
    for k, v in sp2.state_dict().items():
        assert torch.equal(v, sd_o["encoder_model." + k]), k
    # the reference's own full-model file (ref:eval.py:10)
    torch.save(ora.state_dict(), tmp_path / "pytorch_model.bin")
    m2 = SpeechMixEED(spc, txc, down_scale=2)
    res = m2.load_state_dict(torch.load(tmp_path / "pytorch_model.bin"))
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m2.state_dict()["enc_to_dec_proj.weight"], sd_o["enc_to_dec_proj.weight"])
    import pytest
    with pytest.raises(FileNotFoundError):
        SpeechMixEED("voidful/definitely-not-cached-wav2vec2", str(tx_dir))


def test_train_step_accumulates_and_clips_like_the_trainer():
    """training.TrainStep (ref:train.py:291-311: gradient_accumulation_steps + the Trainer's max_grad_norm): two
    micro-batches of 2 must give the update of one batch of 4, and the update is clipped to the global norm."""
    import torch
    from speechmix_b200.training import TrainStep

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.lin = torch.nn.Linear(8, 1)

        def forward(self, x=None, y=None):
            return {"loss": ((self.lin(x)[:, 0] - y) ** 2).mean()}

    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(4, 8, generator=g) * 5, torch.randn(4, generator=g) * 5
    a, b = Toy(), Toy()
    sa = TrainStep(a, torch.optim.SGD(a.parameters(), lr=0.1), grad_accum=2, max_grad_norm=0.5)
    sb = TrainStep(b, torch.optim.SGD(b.parameters(), lr=0.1), grad_accum=1, max_grad_norm=0.5)
    la, na = sa([dict(x=x[:2], y=y[:2]), dict(x=x[2:], y=y[2:])])
    lb, nb = sb([dict(x=x, y=y)])
    assert abs(float(la) - float(lb)) < 1e-5 and abs(float(na) - float(nb)) < 1e-4 and float(na) > 0.5
    assert torch.allclose(a.lin.weight, b.lin.weight, atol=1e-6)
    w0 = Toy().lin.weight
    assert abs(float((a.lin.weight - w0).norm() ** 2 + (a.lin.bias - Toy().lin.bias).norm() ** 2) ** 0.5 - 0.1 * 0.5) < 1e-4


def test_beam_state_matches_transformers_beam_search():
    """speechmix_b200/beam.py (the bookkeeping behind ``generate(num_beams=k)``, ref:speechmix/hf_model.py:304-338 ->
    GenerationMixin) against transformers' own ``generate``: same text model, same encoder states, logits from the HF
    decoder -- so only the search itself is compared.  Covers eos hypotheses finishing early (a frequent token is
    declared eos), length penalties, early_stopping False / True / "never", and BART's forced-eos processor."""
    import torch
    from transformers import GenerationConfig
    from oracle import hf_oracle as O
    from speechmix_b200.beam import BeamState
    torch.manual_seed(0)
    for kind in ("bart-mini", "t5-mini"):
        txc = O.text_config(kind)
        _, text = O.build_backbones(O.speech_config("mini"), txc, seed=0)
        text.eval()
        B, start = 3, txc.decoder_start_token_id
        emb = torch.randn(B, 9, txc.d_model) * 0.5
        feos = getattr(text.generation_config, "forced_eos_token_id", None)
        with torch.no_grad():
            enc = text.get_encoder()(inputs_embeds=emb)
            first = text(encoder_outputs=enc, decoder_input_ids=torch.full((B, 1), start)).logits[:, -1]
            frequent = int(torch.topk(first[0], 3)[1][1])
            for (k, L, eos, lp, es) in [(3, 8, txc.eos_token_id, 1.0, False), (2, 10, frequent, 1.0, False),
                                        (4, 9, frequent, 2.0, False), (3, 9, frequent, 1.0, True), (3, 9, frequent, 0.0, "never")]:
                gc = GenerationConfig(num_beams=k, max_length=L, do_sample=False, early_stopping=es, length_penalty=lp,
                                      eos_token_id=eos, pad_token_id=txc.pad_token_id, decoder_start_token_id=start,
                                      forced_bos_token_id=None, no_repeat_ngram_size=0, min_length=0, use_cache=True)
                ref = text.generate(inputs_embeds=emb, generation_config=gc)
                st = BeamState(B, k, L, [start] * B, eos_token_id=eos, pad_token_id=txc.pad_token_id, length_penalty=lp,
                               early_stopping=es, forced_eos_token_id=feos)
                enc_rep = type(enc)(last_hidden_state=enc.last_hidden_state.repeat_interleave(k, 0))
                rows_seen = []
                while not st.done:
                    dec_in = st.running[:, :, :st.cur].reshape(B * k, st.cur)
                    logits = text(encoder_outputs=enc_rep, decoder_input_ids=dec_in).logits[:, -1]
                    prev = st.running.clone()
                    _, rows, _ = st.step(logits)
                    rows_seen.append(rows)
                    # the row map is what the KV caches are gathered by: new prefixes = old prefixes of those rows
                    assert torch.equal(st.running.view(B * k, -1)[:, :st.cur - 1], prev.view(B * k, -1)[rows][:, :st.cur - 1])
                mine = st.result()
                assert ref.shape == mine.shape and torch.equal(ref, mine), (kind, k, L, eos, lp, es, ref.tolist(), mine.tolist())
                assert all(int(r.min()) >= 0 and int(r.max()) < B * k for r in rows_seen)


def test_bridge_projector_composition_algebra():
    """The identities ops.BridgeProjFn relies on (ref:speechmix/hf_model.py:253-272, 426-430: bare Conv1d(k 2, s 2) then
    nn.Linear, no non-linearity): forward W_eff = Wp.pack(W), b_eff = bp + Wp b over the frame-pair view; backward
    dWp = G pack(W)^T + db_eff (x) b, dpack(W) = Wp^T G, db = Wp^T db_eff, dbp = db_eff with G = dW_eff -- checked in fp64
    against torch autograd of the unfused pair, odd T (the tail frame gets no gradient)."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    B, T, C, D = 2, 7, 8, 6
    x = torch.randn(B, T, C, dtype=torch.float64, requires_grad=True)
    cw = torch.randn(C, C, 2, dtype=torch.float64, requires_grad=True)
    cb = torch.randn(C, dtype=torch.float64, requires_grad=True)
    pw = torch.randn(D, C, dtype=torch.float64, requires_grad=True)
    pb = torch.randn(D, dtype=torch.float64, requires_grad=True)
    y = F.linear(F.conv1d(x.transpose(1, 2), cw, cb, stride=2).transpose(1, 2), pw, pb)
    dy = torch.randn_like(y)
    y.backward(dy)
    P = cw.detach().permute(0, 2, 1).reshape(C, 2 * C)                 # kernels.pack_conv_weight: tap-major
    w_eff, b_eff = pw.detach() @ P, pb.detach() + pw.detach() @ cb.detach()
    t_out = T // 2
    xv = x.detach()[:, :t_out * 2].reshape(B, t_out, 2 * C)            # the copy-free frame-pair view
    assert torch.allclose(xv @ w_eff.t() + b_eff, y.detach(), atol=1e-12)
    G, db_eff = torch.einsum("btd,btk->dk", dy, xv), dy.sum((0, 1))
    assert torch.allclose(G @ P.t() + torch.outer(db_eff, cb.detach()), pw.grad, atol=1e-12)
    assert torch.allclose((pw.detach().t() @ G).view(C, 2, C).permute(0, 2, 1), cw.grad, atol=1e-12)   # unpack_conv_wgrad
    assert torch.allclose(pw.detach().t() @ db_eff, cb.grad, atol=1e-12) and torch.allclose(db_eff, pb.grad, atol=1e-12)
    dx = torch.zeros(B, T, C, dtype=torch.float64)
    dx[:, :t_out * 2] = (dy @ w_eff).reshape(B, t_out * 2, C)
    assert torch.allclose(dx, x.grad, atol=1e-12) and float(x.grad[:, -1].abs().max()) == 0.0


def test_bench_flop_model_matches_survey_table():
    """bench.py's algorithmic FLOP model (the numerator of ``step_tflops`` / ``roofline.step_frac_of_peak``) against
    the per-sample forward GFLOP table of SURVEY.md section 8(d) / BASELINE.md section 3, and the workload table against
    BASELINE.json's configurations."""
    import bench
    want = {"cfg2": 281.95, "cfg3": 660.26, "cfg5": 1486.35}
    for name, gflop in want.items():
        got = bench.fwd_flops_per_sample(bench.WORKLOADS[name]) / 1e9
        assert abs(got - gflop) < 0.02 * gflop, (name, got, gflop)
    assert bench.frames(15.0) == 749 and bench.frames(30.0) == 1499 and bench.frames(5.0) == 249
    # training step: 3 x forward where everything trains; frozen configurations count less than that
    w2 = bench.WORKLOADS["cfg2"]
    assert abs(bench.step_flops_per_sample(w2) - 3 * bench.fwd_flops_per_sample(w2)) < 1.0
    assert bench.step_flops_per_sample(bench.WORKLOADS["cfg3"]) < 3 * bench.fwd_flops_per_sample(bench.WORKLOADS["cfg3"])
    assert bench.step_flops_per_sample(bench.WORKLOADS["gan"]) > bench.step_flops_per_sample(w2)
    assert (w2["batch"], w2["seconds"], w2["t_dec"], w2["kwargs"]["down_scale"]) == (32, 15.0, 64, 2)
    assert bench.WORKLOADS["cfg5"]["batch"] * 8 == 64 and bench.WORKLOADS["cfg5"]["seconds"] == 30.0
    # --dropout switches every site of both backbones on (0 = the measurement plan's default leaves the presets alone)
    from speechmix_b200 import presets
    spc, txc = presets.speech_config("base"), presets.text_config("bart-base")
    bench.set_dropout(spc, txc, 0.0)
    assert spc.hidden_dropout == 0.0 and txc.dropout == 0.0
    bench.set_dropout(spc, txc, 0.1)
    assert (spc.hidden_dropout, spc.attention_dropout, spc.activation_dropout, spc.feat_proj_dropout) == (0.1,) * 4
    assert (txc.dropout, txc.attention_dropout, txc.activation_dropout) == (0.1,) * 3
