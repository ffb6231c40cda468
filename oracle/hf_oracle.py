"""CPU restatement of the reference's HF glue for the SpeechMix hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference (``ref:speechmix/hf_model.py``) is ~700 lines of Python glue that
wires ``transformers`` model classes together; every FLOP of the path runs in
``transformers`` (unpinned ``>=4.12.3`` in ``ref:requirements.txt:2``; the
version installed in this image and on the GPU box is 5.5.0).  ``transformers``
is present on both machines, ``/root/reference`` only exists in the build
container, so this module restates the *glue* and calls the same
``transformers`` classes the reference calls:

* ``OracleEED.__init__``  follows ``ref:speechmix/hf_model.py:188-302``
* ``OracleEED.forward``   follows ``ref:speechmix/hf_model.py:378-447``
* ``OracleEED.cal_loss``  follows ``ref:speechmix/hf_model.py:343-376``
* ``shift_tokens_right``  follows ``ref:speechmix/hf_model.py:25-34``
* ``OracleAdapter``       follows ``ref:speechmix/hf_model.py:465-502`` with the
  4.x-effective hook semantics documented in SURVEY.md section 8(c) caveat A
* ``OracleSelf``          follows ``ref:speechmix/hf_model.py:505-583`` literally
  (SURVEY.md section 8(c) caveat S)
* ``greedy_full_recompute`` follows ``ref:eval.ipynb`` cell 6 (SURVEY.md
  section 8(c) caveat G)

Parity pin: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference
classes from ``/root/reference`` (with an ``s3prl`` import stub) in the build
container on the same random-init checkpoints and asserts that this restatement
reproduces them bit-for-bit; the committed fixtures under ``tests/golden/``
carry those reference outputs so that the GPU box (no ``/root/reference``) can
re-check the oracle before trusting it.

Unlike the reference ctor this restatement takes already-built ``transformers``
modules (or configs + a seed) instead of hub names, because there is no
network; the arithmetic is identical.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

# --------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------


def speech_config(kind: str = "base", deterministic: bool = True, model_type: str = "wav2vec2"):
    """Wav2Vec2/HuBERT config objects for the named size.

    ``deterministic`` zeroes every stochastic knob (dropout, LayerDrop,
    SpecAugment) as SURVEY.md section 8(c) item 3 prescribes for parity runs.
    """
    from transformers import HubertConfig, Wav2Vec2Config

    cls = HubertConfig if model_type == "hubert" else Wav2Vec2Config
    if kind == "base":
        cfg = cls()
    elif kind == "large":  # hubert-large / wav2vec2-large-lv60 style
        cfg = cls(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                  intermediate_size=4096, feat_extract_norm="layer", conv_bias=True,
                  do_stable_layer_norm=True)
    elif kind == "large_group":  # original wav2vec2-large
        cfg = cls(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                  intermediate_size=4096)
    elif kind == "mini":  # small shapes that keep every structural feature of base
        cfg = cls(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                  intermediate_size=512, conv_dim=(128,) * 7,
                  num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
                  vocab_size=32)
    elif kind == "mini_large":  # layer-norm conv stack + stable layer norm, small
        cfg = cls(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                  intermediate_size=512, conv_dim=(128,) * 7,
                  num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
                  vocab_size=32, feat_extract_norm="layer", conv_bias=True,
                  do_stable_layer_norm=True)
    else:
        raise ValueError(kind)
    if deterministic:
        make_speech_config_deterministic(cfg)
    return cfg


def make_speech_config_deterministic(cfg):
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout",
              "feat_proj_dropout", "layerdrop", "mask_time_prob", "mask_feature_prob",
              "final_dropout", "feat_quantizer_dropout"):
        if hasattr(cfg, k):
            setattr(cfg, k, 0.0)
    cfg.apply_spec_augment = False
    return cfg


def text_config(kind: str = "bart-base", deterministic: bool = True):
    from transformers import BartConfig, MBartConfig, T5Config

    if kind == "bart-base":
        cfg = BartConfig(d_model=768, encoder_layers=6, decoder_layers=6,
                         encoder_attention_heads=12, decoder_attention_heads=12,
                         encoder_ffn_dim=3072, decoder_ffn_dim=3072, vocab_size=50265)
    elif kind == "bart-large":
        cfg = BartConfig()
    elif kind == "bart-mini":
        cfg = BartConfig(d_model=256, encoder_layers=2, decoder_layers=2,
                         encoder_attention_heads=4, decoder_attention_heads=4,
                         encoder_ffn_dim=512, decoder_ffn_dim=512, vocab_size=1000,
                         max_position_embeddings=256)
    elif kind == "mbart-large-50":
        cfg = MBartConfig(vocab_size=250054, scale_embedding=True, d_model=1024,
                          encoder_layers=12, decoder_layers=12, encoder_attention_heads=16,
                          decoder_attention_heads=16, encoder_ffn_dim=4096,
                          decoder_ffn_dim=4096, decoder_start_token_id=2)
    elif kind == "mbart-mini":
        cfg = MBartConfig(vocab_size=1000, scale_embedding=True, d_model=256,
                          encoder_layers=2, decoder_layers=2, encoder_attention_heads=4,
                          decoder_attention_heads=4, encoder_ffn_dim=512,
                          decoder_ffn_dim=512, decoder_start_token_id=2,
                          max_position_embeddings=256)
    elif kind == "t5-base":
        cfg = T5Config(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12,
                       vocab_size=32128, feed_forward_proj="relu")
    elif kind == "t5v11-mini":    # t5 v1.1 / mT5 / flan-T5 style: gated-GELU feed-forward, untied LM head
        cfg = T5Config(d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, vocab_size=1000,
                       feed_forward_proj="gated-gelu", tie_word_embeddings=False, decoder_start_token_id=0)
    elif kind == "t5-mini":
        cfg = T5Config(d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4,
                       vocab_size=1000, feed_forward_proj="relu", decoder_start_token_id=0)
    else:
        raise ValueError(kind)
    if deterministic:
        make_text_config_deterministic(cfg)
    return cfg


def make_text_config_deterministic(cfg):
    for k in ("dropout", "attention_dropout", "activation_dropout", "encoder_layerdrop",
              "decoder_layerdrop", "classifier_dropout", "dropout_rate"):
        if hasattr(cfg, k):
            setattr(cfg, k, 0.0)
    return cfg


def build_backbones(speech_cfg, text_cfg, seed: int = 0):
    """Random-init backbones under ``torch.manual_seed(seed)``; speech first,
    text second (the order of SURVEY.md section 8(c) item 5)."""
    from transformers import (BartForConditionalGeneration, HubertModel,
                              MBartForConditionalGeneration, T5ForConditionalGeneration,
                              Wav2Vec2Model)

    torch.manual_seed(seed)
    sp_cls = HubertModel if speech_cfg.model_type == "hubert" else Wav2Vec2Model
    speech = sp_cls(speech_cfg)
    tx_cls = {"bart": BartForConditionalGeneration, "mbart": MBartForConditionalGeneration,
              "t5": T5ForConditionalGeneration}[text_cfg.model_type]
    text = tx_cls(text_cfg)
    return speech, text


# --------------------------------------------------------------------------
# glue restatement
# --------------------------------------------------------------------------


def handle_decoder_input_none(decoder_config, batch=1, device="cpu"):
    """ref:speechmix/hf_model.py:20-22"""
    return torch.tensor([[decoder_config.decoder_start_token_id]] * batch).to(device)


def shift_tokens_right(input_ids: torch.Tensor, pad_token_id: int, decoder_start_token_id: int):
    """ref:speechmix/hf_model.py:25-34"""
    shifted = input_ids.new_zeros(input_ids.shape)
    shifted[:, 1:] = input_ids[:, :-1].clone()
    shifted[:, 0] = decoder_start_token_id
    assert pad_token_id is not None
    shifted.masked_fill_(shifted == -100, pad_token_id)
    return shifted


DEFAULT_FIXED_EXCEPT = ["layer_norm", "encoder_attn", "enc_to_dec_proj", "length_adapter",
                        "layernorm_embedding", "attention"]


class OracleEED(nn.Module):
    """Restates ``HFSpeechMixEED`` (ref:speechmix/hf_model.py:185-447)."""

    def __init__(self, encoder_model, decoder_model, share_layer_ratio=0, down_scale=8,
                 weighted_sum=False, fixed_parameters=False, fixed_except=None, **kwargs):
        super().__init__()
        self.encoder_model = encoder_model
        self.decoder_model = decoder_model
        self.weighted_sum = weighted_sum
        # ref :222-229
        num_nlp_encoder_layers = 0
        enc = self.decoder_model.base_model.encoder
        if hasattr(enc, "layers"):
            num_nlp_encoder_layers = len(enc.layers)
        elif hasattr(enc, "block"):
            num_nlp_encoder_layers = len(enc.block)
        # ref :235-240  drop the LAST int(L*ratio) speech layers
        n_layers = len(self.encoder_model.encoder.layers)
        remove_layers = int(n_layers * share_layer_ratio) if share_layer_ratio != 0 else 0
        self.encoder_model.encoder.layers = self.encoder_model.encoder.layers[:n_layers - remove_layers]
        self.num_speech_encoder_layers = len(self.encoder_model.encoder.layers)
        # ref :253-266
        self.downsize = down_scale
        self.downloop = int(math.log(self.downsize, 2))
        hs = self.encoder_model.config.hidden_size
        if self.downsize > 1:
            self.length_adapters = nn.Sequential(*[
                nn.Conv1d(in_channels=hs, out_channels=hs, kernel_size=2, stride=2)
                for _ in range(self.downloop)])
        else:
            self.length_adapters = nn.Sequential(nn.Identity())
        # ref :268-272
        if self.weighted_sum:
            self.weights_sum = nn.Parameter(torch.zeros(self.num_speech_encoder_layers + 1))
        self.enc_to_dec_proj = nn.Linear(hs, self.decoder_model.config.hidden_size)
        self.custom_modules(**kwargs)
        # ref :274-286
        if fixed_parameters:
            fixed_except = DEFAULT_FIXED_EXCEPT if fixed_except is None else fixed_except
            self.encoder_model.eval()
            self.decoder_model.eval()
            for xcoder in (self.encoder_model.named_parameters, self.decoder_model.named_parameters):
                for name, param in xcoder():
                    if param.requires_grad:
                        param.requires_grad = any(k in name for k in fixed_except)
        # ref :288-302
        self.list_grad = [n for n, p in self.named_parameters() if p.requires_grad]
        self.list_no_grad = [n for n, p in self.named_parameters() if not p.requires_grad]
        self.nlp_emb = self.decoder_model.get_input_embeddings()
        self.speech_encoder_layer = len(self.encoder_model.encoder.layers)
        self.nlp_encoder_layer = num_nlp_encoder_layers
        self.decoder_outputs = None

    @property
    def device(self):
        return next(self.parameters()).device

    def custom_modules(self, **kwargs):
        return None

    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None,
                 decoder_outputs=None, decoder_input_ids=None, labels=None,
                 past_key_values=None, use_cache=None):
        """ref:speechmix/hf_model.py:343-376"""
        if past_key_values is None:
            self.decoder_outputs = None
        if inputs_embeds is not None:
            output = self.decoder_model(
                inputs_embeds=inputs_embeds,
                encoder_outputs=decoder_outputs if decoder_outputs else self.decoder_outputs,
                attention_mask=attention_mask, decoder_input_ids=decoder_input_ids,
                labels=labels, past_key_values=past_key_values, use_cache=use_cache)
        elif text_input_ids is not None:
            output = self.decoder_model(
                input_ids=text_input_ids,
                encoder_outputs=decoder_outputs if decoder_outputs else self.decoder_outputs,
                decoder_input_ids=decoder_input_ids, labels=labels,
                past_key_values=past_key_values, use_cache=use_cache)
        self.decoder_outputs = [output.encoder_last_hidden_state]
        return output

    def bridge(self, encoder_outputs, detail=None):
        """weighted sum + length adapters + projector (ref :410-430)."""
        inputs_embeds = encoder_outputs.last_hidden_state
        if self.weighted_sum:
            stacked = torch.stack(encoder_outputs["hidden_states"], dim=0)
            _, *origin_shape = stacked.shape
            stacked = stacked.view(self.num_speech_encoder_layers + 1, -1)
            norm_weights = F.softmax(self.weights_sum, dim=-1)
            if detail is not None:
                detail["weighted_sum"] = norm_weights
            inputs_embeds = (norm_weights.unsqueeze(-1) * stacked).sum(dim=0).view(*origin_shape)
        if detail is not None:
            detail["shape_before_length_adapter"] = inputs_embeds.shape
        inputs_embeds = self.length_adapters(inputs_embeds.transpose(1, 2)).transpose(1, 2)
        if detail is not None:
            detail["shape_before_enc_dec_projector"] = inputs_embeds.shape
        inputs_embeds = self.enc_to_dec_proj(inputs_embeds)
        if detail is not None:
            detail["shape_after_enc_dec_projector"] = inputs_embeds.shape
        return inputs_embeds

    def forward(self, input_values=None, decoder_text_prompt_ids=None, text_input_ids=None,
                decoder_input_ids=None, labels=None, encoder_outputs=None, decoder_outputs=None,
                past_key_values=None, use_cache=None, return_model_detail=True,
                keep_full_logits=False, attention_mask=None, **kwargs):
        """ref:speechmix/hf_model.py:378-447.  ``attention_mask`` is NOT a reference argument: the reference calls the
        speech encoder without one (:397); given, it is forwarded to exactly that call (the oracle of the
        true-length extension is the reference with the mask passed on to HF's Wav2Vec2Model / HubertModel).
        ``decoder_text_prompt_ids`` takes
        already-tokenised prompt ids (the reference tokenises a string at :433-435;
        there is no real tokenizer offline).  ``keep_full_logits`` additionally
        returns the pre-argmax logits under ``full_logits`` for parity checks."""
        detail = {}
        if encoder_outputs is None:
            if attention_mask is None:
                encoder_outputs = self.encoder_model(input_values, output_hidden_states=True)
            else:
                encoder_outputs = self.encoder_model(input_values, attention_mask=attention_mask,
                                                     output_hidden_states=True)
        if decoder_input_ids is None and labels is None:
            decoder_input_ids = handle_decoder_input_none(
                self.decoder_model.config, encoder_outputs.last_hidden_state.shape[0], device=self.device)
        elif decoder_input_ids is None and labels is not None:
            decoder_input_ids = shift_tokens_right(
                labels, self.decoder_model.config.pad_token_id,
                self.decoder_model.config.decoder_start_token_id)
        inputs_embeds = self.bridge(encoder_outputs, detail if return_model_detail else None)
        if decoder_text_prompt_ids is not None:
            text_prompt = self.nlp_emb(decoder_text_prompt_ids.to(self.device))
            inputs_embeds = torch.cat((text_prompt.expand(inputs_embeds.shape[0], -1, -1), inputs_embeds), 1)
        outputs = self.cal_loss(inputs_embeds=inputs_embeds, decoder_outputs=decoder_outputs,
                                text_input_ids=text_input_ids, decoder_input_ids=decoder_input_ids,
                                labels=labels, past_key_values=past_key_values, use_cache=use_cache)
        if keep_full_logits:
            outputs["full_logits"] = outputs["logits"]
        outputs["speech_last_hidden_state"] = encoder_outputs.last_hidden_state
        outputs["inputs_embeds"] = inputs_embeds
        outputs["logits"] = torch.argmax(outputs["logits"], -1)
        outputs["detail"] = detail
        return outputs


class OracleFixed(OracleEED):
    """ref:speechmix/hf_model.py:450-462"""

    def custom_modules(self, fixed_speech=False, fixed_nlp=True, **kwargs):
        self.encoder_model.eval()
        self.decoder_model.eval()
        if fixed_speech:
            for _, p in self.encoder_model.named_parameters():
                p.requires_grad = False
        if fixed_nlp:
            for _, p in self.decoder_model.named_parameters():
                p.requires_grad = False


class OracleAdapter(OracleEED):
    """ref:speechmix/hf_model.py:465-502 with the 4.x-effective semantics:
    every hooked layer's output is REPLACED by ``adapters[-1](output)`` (the
    lambda at :499-502 late-binds its indices; SURVEY.md section 8(c) caveat A).
    ``adapter_indexing='per_layer'`` gives the presumably intended behaviour."""

    def custom_modules(self, adapter_indexing="reference", **kwargs):
        self.encoder_model.eval()
        self.decoder_model.eval()
        base = self.decoder_model.base_model
        if hasattr(base.encoder, "layers"):
            stacks = [base.encoder.layers, base.decoder.layers]
        else:
            stacks = [base.encoder.block, base.decoder.block]
        for stack in stacks:
            for _, p in stack.named_parameters():
                p.requires_grad = False
        d = self.decoder_model.config.d_model
        self.adapters = nn.ModuleList()
        for stack in stacks:
            for _ in stack:
                self.adapters.append(nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d // 2),
                                                   nn.ReLU(), nn.Linear(d // 2, d)))
        idx = 0
        for stack in stacks:
            for layer in stack:
                j = idx if adapter_indexing == "per_layer" else len(self.adapters) - 1

                def hook(m, i, o, j=j):
                    if isinstance(o, tuple):
                        return (self.adapters[j](o[0]),) + tuple(o[1:])
                    return self.adapters[j](o)

                layer.register_forward_hook(hook)
                idx += 1


class OracleSelf(OracleEED):
    """ref:speechmix/hf_model.py:505-583 followed literally (caveat S)."""

    def custom_modules(self, **kwargs):
        self.encoder_model.eval()
        self.decoder_model.eval()
        for _, p in self.decoder_model.named_parameters():
            p.requires_grad = False

    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None,
                 decoder_input_ids=None, labels=None, **ignored):
        self.decoder_model.eval()
        outputs = self.decoder_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                                     output_hidden_states=True,
                                     decoder_input_ids=decoder_input_ids, labels=labels)
        if labels is not None:
            nlp_outputs = self.decoder_model(input_ids=text_input_ids, output_hidden_states=True,
                                             decoder_input_ids=decoder_input_ids, labels=labels)
            nlp_hidden = nlp_outputs["encoder_hidden_states"][-1]
            speech_hidden = outputs["encoder_hidden_states"][-1]
            hidden = self.decoder_model.config.hidden_size
            attn = torch.bmm(nlp_hidden, speech_hidden.view(nlp_hidden.shape[0], hidden, -1))
            attn = torch.softmax(attn / math.sqrt(hidden), dim=-1)
            projected = torch.bmm(attn, speech_hidden)
            mse = F.mse_loss(projected, nlp_hidden)
            kld = F.kl_div(F.log_softmax(outputs.logits, dim=-1),
                           F.softmax(nlp_outputs.logits, dim=-1), reduction="batchmean")
            outputs["ce_loss"] = outputs.loss
            outputs["mse_loss"] = mse
            outputs["kld_loss"] = kld
            outputs["loss"] = (kld + outputs.loss + mse).mean()
        outputs["encoder_last_hidden_state"] = outputs["encoder_last_hidden_state"]
        return outputs


class OracleGAN(OracleEED):
    """ref:speechmix/hf_model.py:586-694 (HFSpeechMixGAN) followed literally: no cross-entropy term -- the loss is four
    BCE-with-logits terms of ONE Linear(D*D, 1) discriminator over flatten(X.view(D, T) . X.view(T, D)) for X = the speech
    embeddings fed to the text encoder (target 1), the text encoder's states on the label ids (0), the decoder's last
    states on the speech path (1) and on the text path (0).  ``.view`` is a memory reinterpretation, not a transpose.
    The update-phase counters (:609-626) only set ``p.grad = None`` on one parameter family BEFORE this step's backward
    (they clear gradients left over from earlier micro-steps); they are restated as they stand."""

    def custom_modules(self, **kwargs):
        self.discriminator = nn.Linear(self.decoder_model.config.hidden_size ** 2, 1)
        self.des_update = 1000
        self.update_count = 1
        self.keep_update = 1000
        return None

    def gram_features(self, x):
        hidden = self.decoder_model.config.hidden_size
        return torch.bmm(x.view(x.shape[0], hidden, -1), x.view(x.shape[0], -1, hidden)).flatten(start_dim=1)

    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None,
                 decoder_input_ids=None, labels=None, **ignored):
        outputs = self.decoder_model(inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                                     output_hidden_states=True, decoder_input_ids=decoder_input_ids)
        loss = 0
        if labels is not None:
            if self.training:                                                   # ref :609-626
                if self.update_count % self.des_update == 0:
                    if self.keep_update > 0:
                        self.keep_update -= 1
                        for name, p in self.named_parameters():
                            if "discriminator" in name:
                                p.grad = None
                    else:
                        self.keep_update = 1000
                        self.update_count += 1
                else:
                    self.update_count += 1
                    for name, p in self.named_parameters():
                        if "discriminator" not in name:
                            p.grad = None
            nlp_outputs = self.decoder_model(labels, output_hidden_states=True, decoder_input_ids=decoder_input_ids)
            voice_hidden = outputs["decoder_hidden_states"][-1]
            nlp_hidden = nlp_outputs["decoder_hidden_states"][-1]
            nlp_encoder_hidden = nlp_outputs["encoder_hidden_states"][-1]
            bce = torch.nn.BCEWithLogitsLoss()
            terms = {}
            for key, x, target in (("vt_enc_loss", inputs_embeds, 1.0), ("nt_enc_loss", nlp_encoder_hidden, 0.0),
                                   ("vt_loss", voice_hidden, 1.0), ("nt_loss", nlp_hidden, 0.0)):
                logit = self.discriminator(self.gram_features(x)).flatten()
                terms[key] = bce(logit, torch.full((x.shape[0],), target))
                outputs[key.replace("loss", "logit")] = logit
            loss = loss + (terms["vt_loss"] + terms["nt_loss"] + terms["nt_enc_loss"] + terms["vt_enc_loss"])   # ref :692
            for key, v in terms.items():
                outputs[key] = v
        outputs["loss"] = loss
        return outputs


ED_FIXED_EXCEPT = ["layer_norm", "encoder_attn", "enc_to_dec_proj", "length_adapter", "layernorm_embedding",
                   "attention", "encoder"]


class OracleED(nn.Module):
    """ref:speechmix/hf_model.py:82-182 (HFSpeechMixED): ``SpeechEncoderDecoderModel`` over the speech encoder and the
    DECODER half of the text model as a causal LM with cross-attention, feature encoder frozen.  Takes built modules
    instead of hub names: ``text_model`` is the seq2seq model whose decoder weights (and shared embedding) the causal
    decoder receives -- what ``from_encoder_decoder_pretrained`` loads from a BART checkpoint (under transformers 5.x the
    token embedding is reported MISSING and re-drawn; golden and tests equalise it to the checkpoint's shared embedding)."""

    def __init__(self, encoder_model, text_model, fixed_parameters=False, fixed_except=None, **kwargs):
        super().__init__()
        from transformers import AutoModelForCausalLM, SpeechEncoderDecoderConfig, SpeechEncoderDecoderModel
        dec_cfg = copy.deepcopy(text_model.config)
        dec_cfg.is_decoder, dec_cfg.add_cross_attention = True, True
        decoder = AutoModelForCausalLM.from_config(dec_cfg)
        src = text_model.model.decoder.state_dict()
        decoder.model.decoder.load_state_dict(src)
        with torch.no_grad():
            decoder.model.decoder.embed_tokens.weight.copy_(text_model.model.shared.weight)
        cfg = SpeechEncoderDecoderConfig.from_encoder_decoder_configs(encoder_model.config, dec_cfg)
        self.model = SpeechEncoderDecoderModel(config=cfg, encoder=encoder_model, decoder=decoder)
        self.model.config.decoder_start_token_id = self.model.config.decoder.decoder_start_token_id
        self.model.config.pad_token_id = self.model.config.decoder.pad_token_id
        self.model.freeze_feature_encoder()
        if fixed_parameters:
            fixed_except = ED_FIXED_EXCEPT if fixed_except is None else fixed_except
            for name, param in self.model.named_parameters():
                if param.requires_grad:
                    param.requires_grad = any(k in name for k in fixed_except)

    def forward(self, input_values, attention_mask=None, decoder_input_ids=None, labels=None):
        if decoder_input_ids is None and labels is None:
            decoder_input_ids = handle_decoder_input_none(self.model.config.decoder)
        outputs = self.model(input_values=input_values, attention_mask=attention_mask,
                             decoder_input_ids=decoder_input_ids, labels=labels)
        return {"loss": outputs["loss"] if "loss" in outputs else None, "logits": outputs.logits,
                "encoder_last_hidden_state": outputs.encoder_last_hidden_state}


@torch.no_grad()
def greedy_full_recompute(model, input_values, max_length=32, eos_token_id=None):
    """Greedy decode the way ``ref:eval.ipynb`` cell 6 does: no KV cache, every
    step re-runs ``forward`` with ``decoder_input_ids=[start]+previous argmax``."""
    cfg = model.decoder_model.config
    start = cfg.decoder_start_token_id
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    B = input_values.shape[0]
    dec = torch.full((B, 1), start, dtype=torch.long)
    done = torch.zeros(B, dtype=torch.bool)
    enc = model.encoder_model(input_values, output_hidden_states=True)
    for _ in range(max_length - 1):
        ids = model(encoder_outputs=enc, decoder_input_ids=dec)["logits"]
        nxt = ids[:, -1]
        dec = torch.cat([dec, nxt[:, None]], dim=1)
        done |= nxt == eos
        if bool(done.all()):
            break
    return dec


GLUE_PREFIXES = ("length_adapters", "enc_to_dec_proj", "weights_sum", "adapters", "discriminator")


def reinit_glue(model, seed=1):
    """Deterministically re-draw the glue parameters (length adapters, projector,
    weighted-sum logits, adapters).  The reference draws them from the global RNG
    after ``from_pretrained`` has consumed an unspecified amount of it, so the
    golden generator and every test call this on both sides instead."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith(GLUE_PREFIXES):
                scale = 0.5 if name == "weights_sum" else (1.0 if "adapters" in name and name.endswith("0.weight") else 0.05)
                if name.startswith("discriminator"):    # Gram features are sums over frames of products: keep logits O(1)
                    scale = 1e-3
                p.copy_(torch.randn(p.shape, generator=g) * scale)
    return model


def synthetic_batch(batch, seconds, t_dec, vocab, seed=0, ignore_tail=False, rate=16000):
    """SURVEY.md section 8(d) synthetic inputs: N(0,1) audio, labels U[4, vocab)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, int(seconds * rate), generator=g)
    labels = torch.randint(4, vocab, (batch, t_dec), generator=g)
    if ignore_tail:
        labels[:, t_dec - t_dec // 4:] = -100
    return x, labels
