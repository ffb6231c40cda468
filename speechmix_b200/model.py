"""SpeechMix model classes on the B200-native kernels -- the drop-in boundary.

Same constructor surface, attributes, ``forward`` contract and ``state_dict`` layout as the
reference's ``HFSpeechMixEED`` / ``HFSpeechMixFixed`` / ``HFSpeechMixAdapter`` / ``HFSpeechMixSelf`` /
``HFSpeechMixGAN`` / ``HFSpeechMixED`` (ref:speechmix/hf_model.py:82-694); every FLOP runs in libspeechmix_sm100.so.  There is no CPU
path: constructing a model without the built library, or running it without a B200, raises.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops
from .speech import SpeechOutput, speech_from_pretrained
from .text import causal_from_pretrained, text_from_pretrained


def handle_decoder_input_none(decoder_config, batch=1, device="cpu"):
    """ref:speechmix/hf_model.py:20-22"""
    return torch.tensor([[decoder_config.decoder_start_token_id]] * batch).to(device)


def shift_tokens_right(input_ids, pad_token_id, decoder_start_token_id):
    """ref:speechmix/hf_model.py:25-34"""
    shifted = input_ids.new_zeros(input_ids.shape)
    shifted[:, 1:] = input_ids[:, :-1].clone()
    shifted[:, 0] = decoder_start_token_id
    assert pad_token_id is not None, "self.model.config.pad_token_id has to be defined."
    shifted.masked_fill_(shifted == -100, pad_token_id)
    return shifted


def _sub_config(c):
    """sub-config from a config object or from the reference's dict form (``to_dict()`` incl. ``model_type``)"""
    from transformers import AutoConfig, PretrainedConfig
    if isinstance(c, PretrainedConfig):
        return c
    c = dict(c)
    return AutoConfig.for_model(c.pop("model_type"), **c)


try:
    from transformers import PretrainedConfig as _PretrainedConfig
except Exception:   # pragma: no cover - transformers is a hard dependency of the checkpoints, not of the kernels
    _PretrainedConfig = object


class SpeechMixConfig(_PretrainedConfig):
    """Composite config (ref:speechmix/hf_model.py:37-79): ``encoder`` / ``decoder`` sub-configs, ``from_configs``,
    ``to_dict``.  ``encoder`` / ``decoder`` may be config objects or the dicts the reference passes."""

    model_type = "speechmix"
    is_composition = True
    has_no_defaults_at_init = True

    def __init__(self, encoder=None, decoder=None, **kwargs):
        assert encoder is not None and decoder is not None, "Config has to be initialized with encoder and decoder config"
        enc, dec = _sub_config(encoder), _sub_config(decoder)
        if _PretrainedConfig is not object:
            super().__init__(**kwargs)
        self.encoder = enc
        self.decoder = dec
        self.is_encoder_decoder = True
        self.pad_token_id = dec.pad_token_id
        self.decoder_start_token_id = dec.decoder_start_token_id

    @classmethod
    def from_configs(cls, encoder_config, decoder_config, **kwargs):
        """ref:speechmix/hf_model.py:57-72: both arguments are checkpoint names / directories (or config objects)."""
        from transformers import AutoConfig, PretrainedConfig
        from .speech import resolve_checkpoint
        if not isinstance(encoder_config, PretrainedConfig):
            encoder_config = AutoConfig.from_pretrained(resolve_checkpoint(encoder_config))
        if not isinstance(decoder_config, PretrainedConfig):
            decoder_config = AutoConfig.from_pretrained(resolve_checkpoint(decoder_config))
        decoder_config.is_decoder = True
        decoder_config.add_cross_attention = True
        return cls(encoder=encoder_config.to_dict(), decoder=decoder_config.to_dict(), **kwargs)

    def to_dict(self):
        import copy
        output = {k: copy.deepcopy(v) for k, v in self.__dict__.items() if k not in ("encoder", "decoder")}
        output["encoder"] = self.encoder.to_dict()
        output["decoder"] = self.decoder.to_dict()
        output["model_type"] = self.__class__.model_type
        return output


DEFAULT_FIXED_EXCEPT = ["layer_norm", "encoder_attn", "enc_to_dec_proj", "length_adapter", "layernorm_embedding",
                        "attention"]


def _dropout_knobs(speech_cfg, text_cfg):
    """names of the dropout probabilities that are switched on in the two backbone configs"""
    knobs = [("speech", speech_cfg, ("hidden_dropout", "attention_dropout", "activation_dropout", "feat_proj_dropout")),
             ("text", text_cfg, ("dropout", "attention_dropout", "activation_dropout", "dropout_rate"))]
    return ["%s.%s=%g" % (side, k, getattr(cfg, k)) for side, cfg, ks in knobs for k in ks
            if isinstance(getattr(cfg, k, 0.0), float) and getattr(cfg, k, 0.0) > 0.0]


class SpeechMixEED(nn.Module):
    """ref:speechmix/hf_model.py:185-447 (HFSpeechMixEED)."""

    main_input_name = "input_values"

    def __init__(self, speech_model_config, nlp_model_config, share_layer_ratio=0, down_scale=8, weighted_sum=False,
                 fixed_parameters=False, fixed_except=None, tokenizer=None, **kwargs):
        super().__init__()
        _lib.load()  # fail loudly when the CUDA library has not been built
        # ref :206-220 -- `speech_model_config` / `nlp_model_config` are checkpoint directories / hub
        # names in the reference; config objects (random init) are accepted too for offline use.
        self.encoder_model = speech_from_pretrained(speech_model_config)
        self.decoder_model = text_from_pretrained(nlp_model_config)
        # train-mode dropout of the backbones (the reference trains under .train(): ref:train.py:315-330) is applied by
        # the kernels with counter-based masks; this only records whether any site is live
        self.dropout_sites = _dropout_knobs(self.encoder_model.config, self.decoder_model.config)
        self.config = SpeechMixConfig(self.encoder_model.config, self.decoder_model.config)
        self.tokenizer = tokenizer
        if tokenizer is None and isinstance(nlp_model_config, str):
            try:
                from transformers import AutoTokenizer
                from .speech import resolve_checkpoint
                self.tokenizer = AutoTokenizer.from_pretrained(resolve_checkpoint(nlp_model_config))
            except Exception:  # tokenizer files are optional for training on pre-tokenised labels
                self.tokenizer = None
        self.weighted_sum = weighted_sum

        # ref :222-229
        enc = self.decoder_model.base_model.encoder
        num_nlp_encoder_layers = len(enc.layers) if hasattr(enc, "layers") else len(getattr(enc, "block", []))
        # ref :231-251
        n_layers = len(self.encoder_model.encoder.layers)
        print("Before layer sharing num_speech_encoder_layers", n_layers)
        remove_layers = int(n_layers * share_layer_ratio) if share_layer_ratio != 0 else 0
        self.encoder_model.encoder.layers = self.encoder_model.encoder.layers[:n_layers - remove_layers]
        self.num_speech_encoder_layers = len(self.encoder_model.encoder.layers)
        print("After layer sharing ", "num_speech_encoder_layers", self.num_speech_encoder_layers,
              "num_nlp_encoder_layers", num_nlp_encoder_layers, "share_layer_ratio", share_layer_ratio,
              "remove_layers", remove_layers)
        # ref :253-266
        self.downsize = down_scale
        self.downloop = int(math.log(self.downsize, 2))
        hs = self.encoder_model.config.hidden_size
        if self.downsize > 1:
            self.length_adapters = nn.Sequential(*[nn.Conv1d(hs, hs, kernel_size=2, stride=2)
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

    # ------------------------------------------------------------------ reference API surface
    def train(self, mode=True):
        ops.CACHE.invalidate()   # bf16 working copies are re-derived from the fp32 masters after a mode switch
        return super().train(mode)

    @property
    def device(self):
        return next(self.parameters()).device

    def get_encoder(self):
        return self.encoder_model

    def get_decoder(self):
        return self.decoder_model

    def prepare_decoder_input_ids_from_labels(self, labels):
        return shift_tokens_right(labels, self.config.pad_token_id, self.config.decoder_start_token_id)

    def prepare_inputs_for_generation(self, input_ids, past=None, attention_mask=None, use_cache=None,
                                      encoder_outputs=None, **kwargs):
        d = {"encoder_outputs": encoder_outputs, "attention_mask": attention_mask, "use_cache": use_cache,
             "past_key_values": past, "decoder_input_ids": input_ids}
        d.update(kwargs)
        return d

    def custom_modules(self, **kwargs):
        return None

    # ------------------------------------------------------------------ hot path
    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None, decoder_outputs=None,
                 decoder_input_ids=None, labels=None, past_key_values=None, use_cache=None):
        """ref:speechmix/hf_model.py:343-376"""
        if past_key_values is None:
            self.decoder_outputs = None
        cached = decoder_outputs if decoder_outputs else self.decoder_outputs
        if inputs_embeds is not None:
            output = self.decoder_model(inputs_embeds=inputs_embeds, encoder_outputs=cached,
                                        decoder_input_ids=decoder_input_ids, labels=labels)
        elif text_input_ids is not None:
            output = self.decoder_model(input_ids=text_input_ids, encoder_outputs=cached,
                                        decoder_input_ids=decoder_input_ids, labels=labels)
        else:
            raise ValueError("cal_loss needs inputs_embeds or text_input_ids")
        self.decoder_outputs = [output.encoder_last_hidden_state]
        return output

    def bridge(self, encoder_outputs, detail=None):
        """weighted layer sum -> down_scale length adapters -> projector (ref :410-430)."""
        x = encoder_outputs.last_hidden_state
        if self.weighted_sum:
            norm_weights = torch.softmax(self.weights_sum, dim=-1)  # L+1 scalars: torch autograd
            if detail is not None:
                detail["weighted_sum"] = norm_weights
            x = ops.WeightedSumFn.apply(norm_weights, *encoder_outputs.hidden_states)
        if detail is not None:
            detail["shape_before_length_adapter"] = x.shape
        proj = self.enc_to_dec_proj
        fused = False
        if self.downsize > 1:
            convs = list(self.length_adapters)
            if ops.K.FP32_MODE:                      # fp32 verification mode: the plain per-layer kernels
                for conv in convs:
                    x = ops.ConvS2Fn.apply(x, conv.weight, conv.bias, 2)
            else:
                for conv in convs[:-1]:
                    x = ops.ConvK2Fn.apply(x, conv.weight, conv.bias)
                last = convs[-1]
                # last length adapter + projector: one linear map of a frame pair -> ONE GEMM launch (ops.BridgeProjFn)
                fused = ops.bridge_fusion_pays(x.shape[0] * (x.shape[1] // 2), x.shape[2])
                if detail is not None:
                    detail["bridge_fused"] = fused
                if not fused:
                    x = ops.ConvK2Fn.apply(x, last.weight, last.bias)
        if detail is not None:
            detail["shape_before_enc_dec_projector"] = (torch.Size((x.shape[0], x.shape[1] // 2, x.shape[2])) if fused
                                                        else x.shape)
        if fused:
            x = ops.BridgeProjFn.apply(x, last.weight, last.bias, proj.weight, proj.bias)
        else:
            x = ops.linear(x, proj.weight, proj.bias)
        if detail is not None:
            detail["shape_after_enc_dec_projector"] = x.shape
        return x

    def forward(self, input_values=None, decoder_text_prompt=None, text_input_ids=None, decoder_input_ids=None,
                labels=None, encoder_outputs=None, decoder_outputs=None, past_key_values=None, use_cache=None,
                return_model_detail=True, output_attentions=None, output_hidden_states=None, return_dict=None,
                precision=None, attention_mask=None, **kwargs):
        if precision == "fp32":   # verification run: fp32 activations + fp32 arithmetic, inference only
            if attention_mask is not None:
                raise NotImplementedError("fp32 verification mode has no key-padding mask")
            with torch.no_grad(), ops.fp32_verification():
                return self.forward(input_values, decoder_text_prompt, text_input_ids, decoder_input_ids, labels,
                                    encoder_outputs, decoder_outputs, past_key_values, use_cache, return_model_detail)
        """ref:speechmix/hf_model.py:378-447.  Returns a mapping with ``loss`` and ``logits`` (= argmax
        token ids, as the reference returns them at :446) plus the model-detail breadcrumbs."""
        detail = {} if return_model_detail else None
        dev = self.device
        if dev.type != "cuda" or (input_values is not None and not input_values.is_cuda):
            raise RuntimeError("speechmix_b200 runs on a B200 only (model and input_values must be on the GPU); "
                               "there is no CPU fallback")
        # ids may arrive on the host (the reference moves them itself: ref:speechmix/hf_model.py:541-542)
        labels = labels.to(dev) if labels is not None else None
        decoder_input_ids = decoder_input_ids.to(dev) if decoder_input_ids is not None else None
        text_input_ids = text_input_ids.to(dev) if text_input_ids is not None else None
        if self.training and self.dropout_sites:
            ops.DROPOUT.begin_step(dev)      # new masks for this pass (device-side counter: survives CUDA-graph replay)
        if encoder_outputs is None and (torch.is_grad_enabled() or ops.CACHE.dirty):
            # a training pass: the optimizer may have moved the fp32 masters since the last pass (fused
            # optimizers do not bump tensor versions) -> refresh all bf16 working copies in one launch.  A no_grad
            # pass (evaluation in the middle of training, generate) does the same when an optimizer step has run
            # since the last refresh (``dirty`` is set by a global optimizer-step hook and by graph replays).
            ops.CACHE.new_step()
        if encoder_outputs is None:
            # attention_mask is an extension: the reference calls the speech encoder without one
            # (ref:speechmix/hf_model.py:397); given, it is forwarded to that call and to nothing else
            encoder_outputs = self.encoder_model(input_values, attention_mask=attention_mask, output_hidden_states=True)
        if decoder_input_ids is None and labels is None:
            decoder_input_ids = handle_decoder_input_none(self.decoder_model.config,
                                                          encoder_outputs.last_hidden_state.shape[0], device=self.device)
        elif decoder_input_ids is None and labels is not None:
            decoder_input_ids = shift_tokens_right(labels, self.decoder_model.config.pad_token_id,
                                                   self.decoder_model.config.decoder_start_token_id)
        inputs_embeds = self.bridge(encoder_outputs, detail)
        if decoder_text_prompt is not None:
            inputs_embeds = self._prepend_prompt(inputs_embeds, decoder_text_prompt)
        outputs = self.cal_loss(inputs_embeds=inputs_embeds, decoder_outputs=decoder_outputs,
                                text_input_ids=text_input_ids, decoder_input_ids=decoder_input_ids, labels=labels,
                                past_key_values=past_key_values, use_cache=use_cache)
        outputs["speech_last_hidden_state"] = encoder_outputs.last_hidden_state
        outputs["inputs_embeds"] = inputs_embeds
        if detail:
            outputs.update(detail)
            outputs["detail"] = detail
        return outputs

    @torch.no_grad()
    def generate(self, input_values, max_length=32, decoder_text_prompt=None, eos_token_id=None, use_cache=True,
                 precision=None, cuda_graph=False, attention_mask=None, num_beams=1, length_penalty=1.0,
                 early_stopping=False, forced_eos_token_id="config", **kwargs):
        """Greedy decode (ref:eval.py:12-13; loop semantics of ref:eval.ipynb cell 6).  The speech encoder, bridge
        and text encoder run once.  ``use_cache=True`` (default): KV-cached decoder, one pass per new token
        (the role of ref:speechmix/hf_model.py:314-338 ``prepare_inputs_for_generation`` + ``past_key_values``);
        ``use_cache=False``: the notebook's full-prefix recompute.  Both return the same ids.
        ``cuda_graph=True`` replays the whole cached decode loop as one CUDA graph (captured once per input shape
        and ``max_length``).  ``attention_mask``: the variable-length extension of ``forward`` (speech encoder only).
        ``num_beams > 1``: KV-cached beam search with HF's semantics (``length_penalty``, ``early_stopping``,
        ``forced_eos_token_id`` -- default: the text config's value, as GenerationMixin would apply it); the caches are
        gathered per step by ``_reorder_cache`` (ref:speechmix/hf_model.py:337-338)."""
        if precision == "fp32":
            if attention_mask is not None:
                raise NotImplementedError("fp32 verification mode has no key-padding mask")
            with ops.fp32_verification():
                return self.generate(input_values, max_length, decoder_text_prompt, eos_token_id, use_cache,
                                     cuda_graph=cuda_graph, num_beams=num_beams, length_penalty=length_penalty,
                                     early_stopping=early_stopping, forced_eos_token_id=forced_eos_token_id)
        cfg = self.decoder_model.config
        eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
        if ops.CACHE.dirty:          # an optimizer step ran since the working copies were made
            ops.CACHE.new_step()
        enc = self.encoder_model(input_values, attention_mask=attention_mask, output_hidden_states=True)
        B = input_values.shape[0]
        if use_cache:
            inputs_embeds = self.bridge(enc)
            if decoder_text_prompt is not None:
                inputs_embeds = self._prepend_prompt(inputs_embeds, decoder_text_prompt)
            text_enc, _ = self.decoder_model.encode(inputs_embeds=inputs_embeds)
            if int(num_beams) > 1:
                feos = getattr(cfg, "forced_eos_token_id", None) if forced_eos_token_id == "config" else forced_eos_token_id
                return self.decoder_model.beam_decode(text_enc, max_length, int(num_beams), eos_token_id=eos,
                                                      length_penalty=length_penalty, early_stopping=early_stopping,
                                                      forced_eos_token_id=feos)
            if cuda_graph:
                return self.decoder_model.greedy_decode_graph(text_enc, max_length, eos_token_id=eos)
            return self.decoder_model.greedy_decode(text_enc, max_length, eos_token_id=eos)
        if int(num_beams) > 1:
            raise NotImplementedError("beam search runs on the KV-cached decoder (use_cache=True)")
        dec = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long, device=self.device)
        done = torch.zeros(B, dtype=torch.bool, device=self.device)
        text_enc = None
        for _ in range(max_length - 1):
            out = self.forward(encoder_outputs=enc, decoder_input_ids=dec, decoder_text_prompt=decoder_text_prompt,
                               decoder_outputs=text_enc, return_model_detail=False)
            text_enc = [out.encoder_last_hidden_state]
            nxt = out["logits"][:, -1]
            dec = torch.cat([dec, nxt[:, None]], dim=1)
            done |= nxt == eos
            if bool(done.all()):
                break
        return dec

    def _reorder_cache(self, past, beam_idx):
        """ref:speechmix/hf_model.py:337-338 (delegates to the decoder model's cache reorder): ``past`` = list of
        per-layer self-attention caches [rows, T, 2D]; row i of the result continues row ``beam_idx[i]``."""
        from .beam import reorder_cache
        return reorder_cache(past, beam_idx)

    def _prepend_prompt(self, inputs_embeds, decoder_text_prompt):
        """ref:speechmix/hf_model.py:433-436"""
        if isinstance(decoder_text_prompt, str):
            ids = self.tokenizer(decoder_text_prompt, return_tensors="pt")["input_ids"].to(self.device)
        else:
            ids = decoder_text_prompt.to(self.device)
        prompt = ops.EmbedFn.apply(ids, None, self.nlp_emb.weight, None, self.decoder_model.get_encoder().embed_scale, 0, 0)
        return torch.cat((prompt.expand(inputs_embeds.shape[0], -1, -1), inputs_embeds), 1)


ED_FIXED_EXCEPT = ["layer_norm", "encoder_attn", "enc_to_dec_proj", "length_adapter", "layernorm_embedding", "attention",
                   "encoder"]


class _EncoderDecoder(nn.Module):
    """attribute layout (and therefore parameter names) of hf SpeechEncoderDecoderModel: ``encoder``, ``decoder``,
    ``enc_to_dec_proj`` when the hidden sizes differ"""

    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder
        if encoder.config.hidden_size != decoder.config.hidden_size:
            self.enc_to_dec_proj = nn.Linear(encoder.config.hidden_size, decoder.config.hidden_size)


class SpeechMixED(nn.Module):
    """ref:speechmix/hf_model.py:82-182 (HFSpeechMixED): the speech encoder feeding the DECODER half of the text model
    directly (hf SpeechEncoderDecoderModel: no text encoder, no length adapters), feature encoder frozen.
    ``forward(input_values, labels=...)`` returns ``loss`` (mean CE over labels != -100) and -- as the reference does for
    this class -- the FULL-vocabulary ``logits`` (detached; the loss itself comes from the fused LM-head kernel that never
    materialises them), plus ``argmax_ids``.  State-dict keys are those of the reference (``model.encoder.*``,
    ``model.decoder.model.decoder.*``, ``model.decoder.lm_head.weight``)."""

    main_input_name = "input_values"

    def __init__(self, speech_model_config, nlp_model_config, fixed_parameters=False, fixed_except=None, tokenizer=None,
                 **kwargs):
        super().__init__()
        _lib.load()
        self.model = _EncoderDecoder(speech_from_pretrained(speech_model_config), causal_from_pretrained(nlp_model_config))
        self.dropout_sites = _dropout_knobs(self.model.encoder.config, self.model.decoder.config)
        self.config = SpeechMixConfig(self.model.encoder.config, self.model.decoder.config)
        self.tokenizer = tokenizer
        for p in self.model.encoder.feature_extractor.parameters():      # ref :115 freeze_feature_encoder()
            p.requires_grad = False
        if fixed_parameters:                                             # ref :116-122
            fixed_except = ED_FIXED_EXCEPT if fixed_except is None else fixed_except
            for name, param in self.model.named_parameters():
                if param.requires_grad:
                    param.requires_grad = any(k in name for k in fixed_except)
        self.list_grad = [n for n, p in self.named_parameters() if p.requires_grad]
        self.list_no_grad = [n for n, p in self.named_parameters() if not p.requires_grad]

    device = property(lambda self: next(self.parameters()).device)
    encoder_model = property(lambda self: self.model.encoder)
    decoder_model = property(lambda self: self.model.decoder)

    def get_encoder(self):
        return self.model.encoder

    def get_decoder(self):
        return self.model.decoder

    def prepare_decoder_input_ids_from_labels(self, labels):
        cfg = self.model.decoder.config
        return shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)

    def forward(self, input_values, attention_mask=None, decoder_input_ids=None, labels=None, return_full_logits=True):
        if attention_mask is not None:
            raise NotImplementedError("SpeechMixED: the reference call passes no attention mask (ref :157-169)")
        dev = self.device
        if dev.type != "cuda" or not input_values.is_cuda:
            raise RuntimeError("speechmix_b200 runs on a B200 only (model and input_values must be on the GPU); "
                               "there is no CPU fallback")
        dec = self.model.decoder
        labels = labels.to(dev) if labels is not None else None
        if self.training and self.dropout_sites:
            ops.DROPOUT.begin_step(dev)
        if torch.is_grad_enabled() or ops.CACHE.dirty:
            ops.CACHE.new_step()
        if decoder_input_ids is None and labels is None:
            decoder_input_ids = handle_decoder_input_none(dec.config, device=dev)
        elif decoder_input_ids is None:
            decoder_input_ids = self.prepare_decoder_input_ids_from_labels(labels)
        enc = self.model.encoder(input_values).last_hidden_state
        if hasattr(self.model, "enc_to_dec_proj"):
            enc = ops.linear(enc, self.model.enc_to_dec_proj.weight, self.model.enc_to_dec_proj.bias)
        hidden = dec.decode_hidden(decoder_input_ids.to(dev), enc)
        B, T, _ = hidden.shape
        lab = labels if labels is not None else torch.full((B, T), -100, device=dev, dtype=torch.long)
        w, b, scale = dec.lm_head_params()
        loss, ids = ops.LMHeadCEFn.apply(hidden, w, b, lab, scale)
        out = SpeechOutput(loss=loss if labels is not None else None, argmax_ids=ids, encoder_last_hidden_state=enc,
                           decoder_last_hidden_state=hidden)
        with torch.no_grad():
            out["logits"] = dec.full_logits(hidden.detach()) if return_full_logits else ids
        return out


class SpeechMixFixed(SpeechMixEED):
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


class SpeechMixAdapter(SpeechMixEED):
    """ref:speechmix/hf_model.py:465-502.  Text encoder/decoder layers are frozen and every layer's output is
    REPLACED by ``adapter(output)`` with adapter = LayerNorm -> Linear(D, D/2) -> ReLU -> Linear(D/2, D).

    ``adapter_indexing="reference"`` (default) reproduces the reference's effective behaviour: its hook lambda
    late-binds the loop indices, so every layer runs ``adapters[-1]`` (SURVEY.md section 8c caveat A);
    ``"per_layer"`` gives each layer its own adapter."""

    def custom_modules(self, adapter_indexing="reference", **kwargs):
        self.encoder_model.eval()
        self.decoder_model.eval()
        base = self.decoder_model.base_model
        stacks = [base.encoder, base.decoder]

        def layers_of(st):      # BART / mBART: .layers, T5: .block (the reference branches the same way, :470-478)
            return st.layers if hasattr(st, "layers") else st.block
        for st in stacks:
            for _, p in layers_of(st).named_parameters():
                p.requires_grad = False
        d = self.decoder_model.config.d_model
        self.adapters = nn.ModuleList()
        for st in stacks:
            for _ in layers_of(st):
                self.adapters.append(nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d // 2), nn.ReLU(),
                                                   nn.Linear(d // 2, d)))
        self.adapter_indexing = adapter_indexing
        self._adapter_cfg = dict(pre_ln=True, eps=1e-5, act="relu", no_residual=True)
        offset = 0
        for st in stacks:
            st.layer_output_hook = self._make_hook(offset)
            offset += len(layers_of(st))

    def _make_hook(self, offset):
        def hook(layer_index, hidden):
            j = offset + layer_index if self.adapter_indexing == "per_layer" else len(self.adapters) - 1
            a = self.adapters[j]
            return ops.FFNBlockFn.apply(hidden, self._adapter_cfg, a[1].weight, a[1].bias, a[3].weight, a[3].bias,
                                        a[0].weight, a[0].bias)
        return hook


class SpeechMixSelf(SpeechMixEED):
    """ref:speechmix/hf_model.py:505-583 (HFSpeechMixSelf): the text model is frozen and used twice per step --
    on the speech embeddings (student) and on ``text_input_ids`` (teacher) -- and the loss is
    ``CE(student) + KLDiv(student || teacher, batchmean) + MSE(attention-projected speech states, text states)``.

    The reference's ``cal_loss`` rejects the keyword arguments its own ``forward`` passes (SURVEY.md section 8c
    caveat S), so this follows the body of that method literally, including the
    ``.view`` memory reinterpretation of the speech states at :563-565."""

    def custom_modules(self, **kwargs):
        self.encoder_model.eval()
        self.decoder_model.eval()
        for _, p in self.decoder_model.named_parameters():
            p.requires_grad = False

    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None, decoder_outputs=None,
                 decoder_input_ids=None, labels=None, past_key_values=None, use_cache=None):
        if labels is None or text_input_ids is None:
            return super().cal_loss(inputs_embeds=inputs_embeds, text_input_ids=None, attention_mask=attention_mask,
                                    decoder_outputs=decoder_outputs, decoder_input_ids=decoder_input_ids, labels=labels,
                                    past_key_values=past_key_values, use_cache=use_cache)
        lm = self.decoder_model
        lm.eval()                                                              # ref :543
        enc_s, hs_s = lm.encode(inputs_embeds=inputs_embeds, output_hidden_states=True)
        hid_s = lm.decode_hidden(decoder_input_ids, enc_s)
        with torch.no_grad():                                                  # frozen teacher on token ids: no gradient path
            enc_t, _ = lm.encode(input_ids=text_input_ids, output_hidden_states=True)
            hid_t = lm.decode_hidden(decoder_input_ids, enc_t)
        w, b, scale = lm.lm_head_params()
        B = hid_s.shape[0]
        ce, kld, ids = ops.SelfDistillHeadFn.apply(hid_s, hid_t, w, b, labels, scale, B)
        mse = ops.SelfMSEFn.apply(enc_t, enc_s)
        loss = kld + ce + mse
        self.decoder_outputs = [enc_s]
        return SpeechOutput(loss=loss, logits=ids, argmax_ids=ids, ce_loss=ce, kld_loss=kld, mse_loss=mse,
                            encoder_last_hidden_state=enc_s, decoder_last_hidden_state=hid_s,
                            teacher_decoder_last_hidden_state=hid_t, encoder_hidden_states=tuple(hs_s))


class SpeechMixGAN(SpeechMixEED):
    """ref:speechmix/hf_model.py:586-694 (HFSpeechMixGAN).  No cross-entropy: the loss is four BCE-with-logits terms of
    one ``Linear(D*D, 1)`` discriminator over flatten(X.view(D, T) . X.view(T, D)) for X = the speech embeddings fed to
    the text encoder (target 1), the text encoder's states on the label ids (0), and the decoder's last states on the
    speech (1) and text (0) paths.  The labels double as the text model's input ids (:630-633), so they carry no -100.
    The D x D Gram features are never formed (``ops.GramLogitFn``).

    Like ``HFSpeechMixSelf`` the reference's ``cal_loss`` rejects three keyword arguments its own ``forward`` passes; this
    follows the body of the method, including the update-phase counters (:609-626), which only set ``p.grad = None`` on
    one parameter family BEFORE this step's backward."""

    def custom_modules(self, **kwargs):
        self.discriminator = nn.Linear(self.decoder_model.config.hidden_size ** 2, 1)
        self.des_update = 1000
        self.update_count = 1
        self.keep_update = 1000
        return None

    def _phase(self):
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

    def cal_loss(self, inputs_embeds=None, text_input_ids=None, attention_mask=None, decoder_outputs=None,
                 decoder_input_ids=None, labels=None, past_key_values=None, use_cache=None):
        lm = self.decoder_model
        enc_s, hs_s = lm.encode(inputs_embeds=inputs_embeds, output_hidden_states=True)
        hid_s = lm.decode_hidden(decoder_input_ids, enc_s)
        w, b, scale = lm.lm_head_params()
        ids = ops.lm_head_argmax(hid_s.reshape(-1, hid_s.shape[-1]), w, b, scale).view(hid_s.shape[:2])
        out = SpeechOutput(loss=0, logits=ids, argmax_ids=ids, encoder_last_hidden_state=enc_s,
                           decoder_last_hidden_state=hid_s, encoder_hidden_states=tuple(hs_s))
        self.decoder_outputs = [enc_s]
        if labels is None:
            return out
        if self.training:
            self._phase()
        enc_t, _ = lm.encode(input_ids=labels, output_hidden_states=True)
        hid_t = lm.decode_hidden(decoder_input_ids, enc_t)
        dw, db = self.discriminator.weight, self.discriminator.bias
        terms = {}
        for key, x, target in (("vt_enc", inputs_embeds, 1.0), ("nt_enc", enc_t, 0.0), ("vt", hid_s, 1.0),
                               ("nt", hid_t, 0.0)):
            logit = ops.GramLogitFn.apply(x, dw, db)
            out[key + "_logit"] = logit
            terms[key] = F.binary_cross_entropy_with_logits(logit, torch.full_like(logit, target))   # B scalars: torch
            out[key + "_loss"] = terms[key]
        out["loss"] = 0 + (terms["vt"] + terms["nt"] + terms["nt_enc"] + terms["vt_enc"])               # ref :692
        out["teacher_decoder_last_hidden_state"] = hid_t
        out["teacher_encoder_last_hidden_state"] = enc_t
        return out


HFSpeechMixEED = SpeechMixEED
HFSpeechMixGAN = SpeechMixGAN
HFSpeechMixSelf = SpeechMixSelf
HFSpeechMixFixed = SpeechMixFixed
HFSpeechMixAdapter = SpeechMixAdapter
HFSpeechMixED = SpeechMixED
