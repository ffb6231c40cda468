"""BART / mBART sequence-to-sequence LM on the sm_100a kernels: text encoder fed with
speech embeddings, decoder with causal self-attention and cross-attention, tied LM head
with fused cross-entropy (no [rows, vocab] logits in HBM).

Module / parameter names mirror ``BartForConditionalGeneration`` /
``MBartForConditionalGeneration`` (hf:models/bart/modeling_bart.py,
hf:models/mbart/modeling_mbart.py) so reference state dicts load unchanged.
"""
import math

import torch
from torch import nn

from . import kernels as K
from . import ops
from .speech import SpeechOutput, _Attention, load_checkpoint_state, resolve_checkpoint


def _with_dropout(cfg, drop, training):
    """block configuration with the train-mode dropout probabilities of the HF layer (p_attn on the attention
    probabilities, p_hidden on the block outputs before the residual add, p_act after the FFN activation)"""
    if training and drop and any(v > 0 for v in drop.values()):
        return dict(cfg, **drop)
    return cfg


def _drop_probs(config):
    if config.model_type == "t5":
        p = float(config.dropout_rate)
        return dict(p_attn=p, p_hidden=p, p_act=p)
    return dict(p_attn=float(config.attention_dropout), p_hidden=float(config.dropout), p_act=float(config.activation_dropout))


class _EncoderLayer(nn.Module):
    """hf:...bart.py:261-309 (post-LN) / hf:...mbart.py:274-328 (pre-LN)"""

    def __init__(self, d, heads, ffn, act, pre_ln, drop=None):
        super().__init__()
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)
        self.cfg = dict(heads=heads, causal=False, pre_ln=pre_ln, eps=1e-5, act=act)
        self.drop = drop or {}

    def forward(self, x):
        cfg = _with_dropout(self.cfg, self.drop, self.training)
        x = ops.AttnBlockFn.apply(x, None, cfg, *self.self_attn.params(), self.self_attn_layer_norm.weight,
                                  self.self_attn_layer_norm.bias)
        return ops.FFNBlockFn.apply(x, cfg, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                    self.final_layer_norm.weight, self.final_layer_norm.bias)


class _DecoderLayer(nn.Module):
    """hf:...bart.py:312-391 / hf:...mbart.py:331-430"""

    def __init__(self, d, heads, ffn, act, pre_ln, drop=None):
        super().__init__()
        self.drop = drop or {}
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = _Attention(d)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)
        self.cfg_self = dict(heads=heads, causal=True, pre_ln=pre_ln, eps=1e-5, act=act)
        self.cfg_cross = dict(heads=heads, causal=False, pre_ln=pre_ln, eps=1e-5, act=act)

    def forward(self, x, enc):
        cfg_self = _with_dropout(self.cfg_self, self.drop, self.training)
        cfg_cross = _with_dropout(self.cfg_cross, self.drop, self.training)
        x = ops.AttnBlockFn.apply(x, None, cfg_self, *self.self_attn.params(), self.self_attn_layer_norm.weight,
                                  self.self_attn_layer_norm.bias)
        x = ops.AttnBlockFn.apply(x, enc, cfg_cross, *self.encoder_attn.params(),
                                  self.encoder_attn_layer_norm.weight, self.encoder_attn_layer_norm.bias)
        return ops.FFNBlockFn.apply(x, cfg_cross, self.fc1.weight, self.fc1.bias, self.fc2.weight,
                                    self.fc2.bias, self.final_layer_norm.weight, self.final_layer_norm.bias)


class _Stack(nn.Module):
    POS_OFFSET = 2  # BartLearnedPositionalEmbedding / MBartLearnedPositionalEmbedding (hf:...bart.py:74-98)

    def __init__(self, config, shared, is_decoder):
        super().__init__()
        d = config.d_model
        self.config = config
        self.is_decoder = is_decoder
        self.embed_tokens = shared
        self.embed_scale = math.sqrt(d) if config.scale_embedding else 1.0
        self.embed_positions = nn.Embedding(config.max_position_embeddings + self.POS_OFFSET, d)
        pre_ln = config.model_type == "mbart"
        n = config.decoder_layers if is_decoder else config.encoder_layers
        heads = config.decoder_attention_heads if is_decoder else config.encoder_attention_heads
        ffn = config.decoder_ffn_dim if is_decoder else config.encoder_ffn_dim
        if d // heads != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")
        cls = _DecoderLayer if is_decoder else _EncoderLayer
        self.layers = nn.ModuleList([cls(d, heads, ffn, config.activation_function, pre_ln, _drop_probs(config))
                                     for _ in range(n)])
        self.layernorm_embedding = nn.LayerNorm(d)
        if pre_ln:
            self.layer_norm = nn.LayerNorm(d)
        self.pre_ln = pre_ln
        self.layer_output_hook = None  # callable(layer_index, hidden) -> hidden (SpeechMixAdapter)

    def embed(self, input_ids=None, inputs_embeds=None, t_start=0):
        """tokens*scale (or given embeddings, unscaled: hf:...bart.py:520-524) + learned positions, then LN."""
        x = ops.EmbedFn.apply(input_ids, inputs_embeds, self.embed_tokens.weight if input_ids is not None else None,
                              self.embed_positions.weight, self.embed_scale, self.POS_OFFSET, t_start)
        x = ops.layer_norm(x, self.layernorm_embedding.weight, self.layernorm_embedding.bias, 1e-5)
        return ops.dropout(x, float(self.config.dropout), self.training)    # hf:...bart.py:530 / :640

    def forward(self, input_ids=None, inputs_embeds=None, encoder_hidden_states=None, output_hidden_states=False):
        x = self.embed(input_ids, inputs_embeds)
        hs = [x] if output_hidden_states else None
        for li, layer in enumerate(self.layers):
            x = layer(x, encoder_hidden_states) if self.is_decoder else layer(x)
            if self.layer_output_hook is not None:
                x = self.layer_output_hook(li, x)
            if output_hidden_states:
                hs.append(x)
        if self.pre_ln:
            x = ops.layer_norm(x, self.layer_norm.weight, self.layer_norm.bias, 1e-5)
            if output_hidden_states:
                hs[-1] = x
        return x, hs


class _Body(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.shared = nn.Embedding(config.vocab_size, config.d_model, config.pad_token_id)
        self.encoder = _Stack(config, self.shared, is_decoder=False)
        self.decoder = _Stack(config, self.shared, is_decoder=True)


class Seq2SeqLM(nn.Module):
    """Drop-in for ``AutoModelForSeq2SeqLM.from_pretrained(<bart|mbart>)`` on the SpeechMix path
    (call site ref:speechmix/hf_model.py:357-374)."""

    def __init__(self, config):
        super().__init__()
        if config.model_type not in ("bart", "mbart"):
            raise NotImplementedError("text backbone %r is not wired to the sm_100a kernels yet" % config.model_type)
        self.config = config
        self.model = _Body(config)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.lm_head.weight = self.model.shared.weight  # tied (hf:...bart.py:806-812)
        self.register_buffer("final_logits_bias", torch.zeros((1, config.vocab_size)))

    @property
    def base_model(self):
        return self.model

    @property
    def device(self):
        return self.model.shared.weight.device

    def get_input_embeddings(self):
        return self.model.shared

    def get_encoder(self):
        return self.model.encoder

    def lm_head_params(self):
        """(weight [V, D], bias [1, V], logit scale)  hf:...bart.py:940-942"""
        return self.model.shared.weight, self.final_logits_bias, 1.0

    def encode(self, input_ids=None, inputs_embeds=None, output_hidden_states=False):
        return self.model.encoder(input_ids=input_ids, inputs_embeds=inputs_embeds,
                                  output_hidden_states=output_hidden_states)

    def decode_hidden(self, decoder_input_ids, encoder_hidden_states):
        x, _ = self.model.decoder(input_ids=decoder_input_ids, encoder_hidden_states=encoder_hidden_states)
        return x

    @torch.no_grad()
    def greedy_decode(self, enc, max_length, eos_token_id=None, sync_eos=True):
        """KV-cached greedy decode over encoder states ``enc`` [B, Ts, D]: one decoder pass per new token
        (hf:...bart.py:143-258 cache branch; loop semantics of ref:eval.ipynb cell 6).  Returns ids [B, <=max_length].
        ``sync_eos=False`` never reads the device (no early exit): the whole loop is CUDA-graph capturable."""
        cfg, dec = self.config, self.model.decoder
        B, dev = enc.shape[0], enc.device
        eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
        ids = torch.full((B, max_length), cfg.decoder_start_token_id, dtype=torch.long, device=dev)
        caches, cross = self._decode_state(enc, B, max_length)
        w, b, scale = self.lm_head_params()
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        for t in range(max_length - 1):
            x = self._decode_step(ids[:, t:t + 1].contiguous(), t, caches, cross)
            nxt = ops.lm_head_argmax(x, w, b, scale)
            ids[:, t + 1] = nxt
            if sync_eos:
                done |= nxt == eos
                if bool(done.all()):
                    return ids[:, :t + 2]
        return ids

    def _decode_step(self, tok, t, caches, cross):
        """one KV-cached decoder pass: tok [B, 1] at position t -> last hidden state [B, D] (writes row t of the caches)"""
        dec = self.model.decoder
        B, D = tok.shape[0], self.config.d_model
        x = dec.embed(tok, None, t_start=t).view(B, D)
        for li, l in enumerate(dec.layers):
            x = ops.decode_self_attn_step(x, l.cfg_self, *l.self_attn.params(), l.self_attn_layer_norm.weight,
                                          l.self_attn_layer_norm.bias, caches[li], t)
            ea = l.encoder_attn
            x = ops.decode_cross_attn_step(x, l.cfg_cross, ea.q_proj.weight, ea.q_proj.bias, ea.out_proj.weight,
                                           ea.out_proj.bias, l.encoder_attn_layer_norm.weight,
                                           l.encoder_attn_layer_norm.bias, cross[li])
            x = ops.decode_ffn_step(x, l.cfg_cross, l.fc1.weight, l.fc1.bias, l.fc2.weight, l.fc2.bias,
                                    l.final_layer_norm.weight, l.final_layer_norm.bias)
            if dec.layer_output_hook is not None:
                x = dec.layer_output_hook(li, x.view(B, 1, D)).reshape(B, D)
        if dec.pre_ln:
            x = ops._ln_maybe(x, dec.layer_norm.weight, dec.layer_norm.bias, 1e-5, False)
        return x

    def _decode_state(self, enc, rows, max_length):
        """(self-attention caches [rows, max_length, 2D] per layer, cross-attention k | v of ``enc`` per layer)"""
        dec, D = self.model.decoder, self.config.d_model
        caches = [torch.empty(rows, max_length, 2 * D, device=enc.device, dtype=K.act_dtype()) for _ in dec.layers]
        cross = [ops.cross_kv(enc, l.encoder_attn.k_proj.weight, l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.weight,
                              l.encoder_attn.v_proj.bias) for l in dec.layers]
        return caches, cross

    def full_logits(self, hidden):
        """fp32 [.., V] logits, materialised -- parity tests / debugging only, never on the training path."""
        h2 = hidden.reshape(-1, hidden.shape[-1]).contiguous()
        lg = K.linear_fwd(h2, ops.w16(self.model.shared.weight), self.final_logits_bias.reshape(-1).float().contiguous(),
                          out_f32=True)
        return lg.view(*hidden.shape[:-1], -1)

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, decoder_input_ids=None, labels=None,
                encoder_outputs=None, output_hidden_states=False, past_key_values=None, use_cache=None, **kwargs):
        if attention_mask is not None:
            raise NotImplementedError("the SpeechMix path never passes an attention mask (SURVEY section 8)")
        cfg = self.config
        if decoder_input_ids is None and labels is not None:
            from .model import shift_tokens_right
            decoder_input_ids = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        enc_hs = None
        if encoder_outputs is None:
            enc, enc_hs = self.encode(input_ids, inputs_embeds, output_hidden_states)
        else:
            enc = encoder_outputs[0] if isinstance(encoder_outputs, (list, tuple)) else encoder_outputs
        hidden = self.decode_hidden(decoder_input_ids, enc)
        B, T, _ = hidden.shape
        lab = labels if labels is not None else torch.full((B, T), -100, device=hidden.device, dtype=torch.long)
        w, b, scale = self.lm_head_params()
        loss, ids = ops.LMHeadCEFn.apply(hidden, w, b, lab, scale)
        out = SpeechOutput(loss=loss if labels is not None else None, logits=ids, argmax_ids=ids,
                           encoder_last_hidden_state=enc, decoder_last_hidden_state=hidden)
        if output_hidden_states:
            out["encoder_hidden_states"] = tuple(enc_hs) if enc_hs is not None else None
        return out



# =============================================================================================
# T5 (hf:models/t5/modeling_t5.py): RMSNorm pre-norm blocks, bias-free linears, UNSCALED attention
# scores plus a bucketed relative position bias owned by block 0 of each stack and shared by all of
# its blocks, ReLU feed-forward, tied LM head scaled by d_model**-0.5.
# =============================================================================================
class _CausalBody(nn.Module):
    """hf BartDecoderWrapper: ``model.decoder`` of a causal LM"""

    def __init__(self, config):
        super().__init__()
        self.decoder = _Stack(config, nn.Embedding(config.vocab_size, config.d_model, config.pad_token_id), is_decoder=True)


class CausalLM(nn.Module):
    """The decoder half of a BART / mBART model as a causal LM WITH cross-attention and a tied LM head
    (hf:models/bart/modeling_bart.py BartForCausalLM, built by ``SpeechEncoderDecoderModel.from_encoder_decoder_pretrained``
    with ``is_decoder=True, add_cross_attention=True``): the decoder of ``SpeechMixED`` (ref:speechmix/hf_model.py:104).
    Same parameter names as HF (``model.decoder.*``, ``lm_head.weight``)."""

    def __init__(self, config):
        super().__init__()
        if config.model_type not in ("bart", "mbart"):
            raise NotImplementedError("causal decoder %r is not wired to the sm_100a kernels" % config.model_type)
        self.config = config
        self.model = _CausalBody(config)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.lm_head.weight = self.model.decoder.embed_tokens.weight

    device = property(lambda self: self.lm_head.weight.device)

    def lm_head_params(self):
        return self.lm_head.weight, None, 1.0

    def decode_hidden(self, decoder_input_ids, encoder_hidden_states):
        x, _ = self.model.decoder(input_ids=decoder_input_ids, encoder_hidden_states=encoder_hidden_states)
        return x

    def full_logits(self, hidden):
        h2 = hidden.reshape(-1, hidden.shape[-1]).contiguous()
        return K.linear_fwd(h2, ops.w16(self.lm_head.weight), None, out_f32=True).view(*hidden.shape[:-1], -1)


def causal_from_pretrained(path_or_config):
    """config object (random init) or a local / cached SEQ2SEQ checkpoint whose decoder (``model.decoder.*``) and shared
    token embedding are loaded -- what ``AutoModelForCausalLM.from_pretrained(<bart checkpoint>)`` is meant to give
    (transformers 5.x reports the token embedding MISSING there and re-draws it; we take ``model.shared``)."""
    from transformers import AutoConfig, PretrainedConfig
    if isinstance(path_or_config, PretrainedConfig):
        return CausalLM(path_or_config)
    path = resolve_checkpoint(path_or_config)
    model = CausalLM(AutoConfig.from_pretrained(path))
    sd = dict(load_checkpoint_state(path))
    emb = next((sd[k] for k in ("model.decoder.embed_tokens.weight", "model.shared.weight", "lm_head.weight") if k in sd), None)
    for k in ("model.decoder.embed_tokens.weight", "lm_head.weight"):
        sd.setdefault(k, emb)
    own = model.state_dict()
    missing = [k for k in own if k not in sd or sd[k] is None]
    if missing:
        raise RuntimeError("checkpoint %s lacks %d decoder tensors, e.g. %s" % (path, len(missing), missing[:3]))
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    return model


class _RMSNorm(nn.Module):
    """hf:...t5.py:46-69 (T5LayerNorm): weight only."""

    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class _T5Attention(nn.Module):
    """hf:...t5.py:153-345; parameter container (q, k, v, o without bias, optional bucket embedding)."""

    def __init__(self, config, has_relative_attention_bias):
        super().__init__()
        inner = config.num_heads * config.d_kv
        self.q = nn.Linear(config.d_model, inner, bias=False)
        self.k = nn.Linear(config.d_model, inner, bias=False)
        self.v = nn.Linear(config.d_model, inner, bias=False)
        self.o = nn.Linear(inner, config.d_model, bias=False)
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(config.relative_attention_num_buckets, config.num_heads)

    def params(self):
        return (self.q.weight, None, self.k.weight, None, self.v.weight, None, self.o.weight, None)


class _T5LayerSelfAttention(nn.Module):
    def __init__(self, config, has_relative_attention_bias):
        super().__init__()
        self.SelfAttention = _T5Attention(config, has_relative_attention_bias)
        self.layer_norm = _RMSNorm(config.d_model)


class _T5LayerCrossAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.EncDecAttention = _T5Attention(config, False)
        self.layer_norm = _RMSNorm(config.d_model)


class _T5DenseActDense(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.wi = nn.Linear(config.d_model, config.d_ff, bias=False)
        self.wo = nn.Linear(config.d_ff, config.d_model, bias=False)


class _T5DenseGatedActDense(nn.Module):
    """hf:...t5.py T5DenseGatedActDense (t5 v1.1, mT5, flan-T5): wo(act(wi_0 x) * wi_1 x)"""

    def __init__(self, config):
        super().__init__()
        self.wi_0 = nn.Linear(config.d_model, config.d_ff, bias=False)
        self.wi_1 = nn.Linear(config.d_model, config.d_ff, bias=False)
        self.wo = nn.Linear(config.d_ff, config.d_model, bias=False)


class _T5LayerFF(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gated = bool(config.is_gated_act)
        self.DenseReluDense = _T5DenseGatedActDense(config) if self.gated else _T5DenseActDense(config)
        self.layer_norm = _RMSNorm(config.d_model)

    def forward(self, x, cfg):
        d = self.DenseReluDense
        if self.gated:
            return ops.GatedFFNBlockFn.apply(x, cfg, d.wi_0.weight, d.wi_1.weight, d.wo.weight, self.layer_norm.weight)
        return ops.FFNBlockFn.apply(x, cfg, d.wi.weight, None, d.wo.weight, None, self.layer_norm.weight, None)

    def decode_step(self, x2, cfg):
        d = self.DenseReluDense
        if self.gated:
            return ops.decode_gated_ffn_step(x2, cfg, d.wi_0.weight, d.wi_1.weight, d.wo.weight, self.layer_norm.weight)
        return ops.decode_ffn_step(x2, cfg, d.wi.weight, None, d.wo.weight, None, self.layer_norm.weight, None)


class _T5Block(nn.Module):
    """hf:...t5.py:411-500"""

    def __init__(self, config, is_decoder, has_relative_attention_bias):
        super().__init__()
        self.is_decoder = is_decoder
        self.layer = nn.ModuleList([_T5LayerSelfAttention(config, has_relative_attention_bias)])
        if is_decoder:
            self.layer.append(_T5LayerCrossAttention(config))
        self.layer.append(_T5LayerFF(config))
        eps = config.layer_norm_epsilon
        base = dict(heads=config.num_heads, pre_ln=True, rms=True, eps=eps, scale=1.0, act=config.dense_act_fn)
        self.cfg_self = dict(base, causal=is_decoder)
        self.cfg_cross = dict(base, causal=False)
        self.drop = _drop_probs(config)

    def forward(self, x, pos_bias, enc=None):
        cfg_self = _with_dropout(self.cfg_self, self.drop, self.training)
        cfg_cross = _with_dropout(self.cfg_cross, self.drop, self.training)
        sa = self.layer[0]
        x = ops.AttnBlockFn.apply(x, None, cfg_self, *sa.SelfAttention.params(), sa.layer_norm.weight, None, pos_bias)
        if self.is_decoder:
            ca = self.layer[1]
            x = ops.AttnBlockFn.apply(x, enc, cfg_cross, *ca.EncDecAttention.params(), ca.layer_norm.weight, None,
                                      None)
        return self.layer[-1](x, cfg_cross)


class _T5Stack(nn.Module):
    """hf:...t5.py:617-770"""

    def __init__(self, config, shared, is_decoder):
        super().__init__()
        self.config = config
        self.is_decoder = is_decoder
        self.embed_tokens = shared
        self.embed_scale = 1.0
        n = config.num_decoder_layers if is_decoder else config.num_layers
        self.block = nn.ModuleList([_T5Block(config, is_decoder, i == 0) for i in range(n)])
        for i, blk in enumerate(self.block):
            blk._index = i
        self.final_layer_norm = _RMSNorm(config.d_model)
        self.layer_output_hook = None

    def forward(self, input_ids=None, inputs_embeds=None, encoder_hidden_states=None, output_hidden_states=False):
        cfg = self.config
        if input_ids is not None:
            x = ops.EmbedFn.apply(input_ids, None, self.embed_tokens.weight, None, 1.0, 0, 0)
        else:
            x = inputs_embeds if inputs_embeds.dtype == K.act_dtype() else ops.EmbedFn.apply(None, inputs_embeds, None, None, 1.0, 0, 0)
        x = ops.dropout(x, float(cfg.dropout_rate), self.training)          # hf:...t5.py:700 (embedding dropout)
        T = x.shape[1]
        table = ops.t5_bucket_table(T, T, not self.is_decoder, cfg.relative_attention_num_buckets,
                                    cfg.relative_attention_max_distance, x.device)
        rel = self.block[0].layer[0].SelfAttention.relative_attention_bias.weight
        pos_bias = ops.RelPosBiasFn.apply(rel, table, T, T, 0)
        hs = [x] if output_hidden_states else None
        for li, blk in enumerate(self.block):
            x = blk(x, pos_bias, encoder_hidden_states)
            if self.layer_output_hook is not None:
                x = self.layer_output_hook(li, x)
            if output_hidden_states:
                hs.append(x)
        x = ops.layer_norm(x, self.final_layer_norm.weight, None, cfg.layer_norm_epsilon, rms_only=True)
        x = ops.dropout(x, float(cfg.dropout_rate), self.training)          # hf:...t5.py:765 (after the final norm)
        if output_hidden_states:
            hs[-1] = x
        return x, hs


class T5Seq2SeqLM(nn.Module):
    """Drop-in for ``AutoModelForSeq2SeqLM.from_pretrained(<t5>)`` on the SpeechMix path
    (hf:...t5.py:940-1130, call site ref:speechmix/hf_model.py:357-374)."""

    def __init__(self, config):
        super().__init__()
        if config.dense_act_fn not in ("relu", "gelu", "gelu_new"):
            raise NotImplementedError("T5 feed-forward activation %r is not wired to the kernels" % (config.dense_act_fn,))
        if config.d_kv != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")
        self.config = config
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = _T5Stack(config, self.shared, is_decoder=False)
        self.decoder = _T5Stack(config, self.shared, is_decoder=True)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        if config.tie_word_embeddings:
            self.lm_head.weight = self.shared.weight

    base_model = property(lambda self: self)
    device = property(lambda self: self.shared.weight.device)

    def get_input_embeddings(self):
        return self.shared

    def get_encoder(self):
        return self.encoder

    def lm_head_params(self):
        """(weight [V, D], bias or None, logit scale)  hf:...t5.py:1105-1110"""
        # original T5 scales the decoder output by d_model^-0.5 before the (tied) LM head, v1.1 / mT5 do not.  transformers
        # >= 5 carries that as ``scale_decoder_outputs`` (set from the checkpoint's ``tie_word_embeddings``, which it then
        # forces to True); older configs only have ``tie_word_embeddings``.
        cfg = self.config
        scaled = getattr(cfg, "scale_decoder_outputs", None)
        if scaled is None:
            scaled = bool(cfg.tie_word_embeddings)
        return self.lm_head.weight, None, (cfg.d_model ** -0.5) if scaled else 1.0

    def encode(self, input_ids=None, inputs_embeds=None, output_hidden_states=False):
        return self.encoder(input_ids=input_ids, inputs_embeds=inputs_embeds, output_hidden_states=output_hidden_states)

    def decode_hidden(self, decoder_input_ids, encoder_hidden_states):
        x, _ = self.decoder(input_ids=decoder_input_ids, encoder_hidden_states=encoder_hidden_states)
        return x

    def full_logits(self, hidden):
        w, _, scale = self.lm_head_params()
        h2 = hidden.reshape(-1, hidden.shape[-1]).contiguous()
        return K.linear_fwd(h2, ops.w16(w), None, out_f32=True, alpha=scale).view(*hidden.shape[:-1], -1)

    @torch.no_grad()
    def greedy_decode(self, enc, max_length, eos_token_id=None, sync_eos=True):
        """KV-cached greedy decode (hf:...t5.py:248-345 cache branch): the relative position bias of step t is
        row t of the causal bucket table (query offset t, keys 0..t)."""
        cfg, dec = self.config, self.decoder
        B, dev = enc.shape[0], enc.device
        eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
        ids = torch.full((B, max_length), cfg.decoder_start_token_id, dtype=torch.long, device=dev)
        caches, cross = self._decode_state(enc, B, max_length)
        w, b, scale = self.lm_head_params()
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        for t in range(max_length - 1):
            x = self._decode_step(ids[:, t:t + 1].contiguous(), t, caches, cross)
            nxt = ops.lm_head_argmax(x, w, b, scale)
            ids[:, t + 1] = nxt
            if sync_eos:
                done |= nxt == eos
                if bool(done.all()):
                    return ids[:, :t + 2]
        return ids

    def _decode_step(self, tok, t, caches, cross):
        """one KV-cached decoder pass: tok [B, 1] at position t -> last hidden state [B, D] (writes row t of the caches)"""
        cfg, dec = self.config, self.decoder
        B, D, dev = tok.shape[0], cfg.d_model, tok.device
        rel = dec.block[0].layer[0].SelfAttention.relative_attention_bias.weight.detach().float().contiguous()
        x = ops.EmbedFn.apply(tok, None, dec.embed_tokens.weight, None, 1.0, 0, 0).view(B, D)
        table = ops.t5_bucket_table(1, t + 1, False, cfg.relative_attention_num_buckets,
                                    cfg.relative_attention_max_distance, dev, q_offset=t)
        pos_bias = K.relpos_bias_fwd(rel, table, cfg.num_heads, 1, t + 1, q_offset=t)
        for blk in dec.block:
            sa, ca, ff = blk.layer[0], blk.layer[1], blk.layer[2]
            x = ops.decode_self_attn_step(x, blk.cfg_self, *sa.SelfAttention.params(), sa.layer_norm.weight, None,
                                          caches[blk._index], t, pos_bias)
            a = ca.EncDecAttention
            x = ops.decode_cross_attn_step(x, blk.cfg_cross, a.q.weight, None, a.o.weight, None, ca.layer_norm.weight,
                                           None, cross[blk._index])
            x = ff.decode_step(x, blk.cfg_cross)
            if dec.layer_output_hook is not None:       # SpeechMixAdapter (ref:speechmix/hf_model.py:486-502)
                x = dec.layer_output_hook(blk._index, x.view(B, 1, D)).reshape(B, D)
        return ops._ln_maybe(x, dec.final_layer_norm.weight, None, cfg.layer_norm_epsilon, True)

    def _decode_state(self, enc, rows, max_length):
        cfg, dec = self.config, self.decoder
        Hi = cfg.num_heads * cfg.d_kv
        caches = [torch.empty(rows, max_length, 2 * Hi, device=enc.device, dtype=K.act_dtype()) for _ in dec.block]
        cross = [ops.cross_kv(enc, blk.layer[1].EncDecAttention.k.weight, None, blk.layer[1].EncDecAttention.v.weight, None)
                 for blk in dec.block]
        return caches, cross

    forward = None  # assigned below (shared with the BART-family class)


T5Seq2SeqLM.forward = Seq2SeqLM.forward


@torch.no_grad()
def _beam_decode(self, enc, max_length, num_beams, eos_token_id=None, length_penalty=1.0, early_stopping=False,
                 forced_eos_token_id=None):
    """KV-cached beam search over encoder states ``enc`` [B, Ts, D] (the search of hf:generation/utils.py
    ``_beam_search``, restated in beam.BeamState): every utterance runs ``num_beams`` decoder rows; per step the fp32
    next-token logits of all rows are materialised ([B k, V] -- inference only), the state machine picks the surviving
    rows and the self-attention caches are gathered accordingly (``_reorder_cache``, ref:speechmix/hf_model.py:337-338);
    the cross-attention keys / values are projected once per utterance and repeated per beam."""
    from .beam import BeamState, reorder_cache
    cfg = self.config
    B, k = enc.shape[0], int(num_beams)
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    caches, cross = self._decode_state(enc, B * k, max_length)        # cross: [B, Ts, 2 Hi] per layer
    cross = [c.repeat_interleave(k, 0) for c in cross]
    state = BeamState(B, k, max_length, [cfg.decoder_start_token_id] * B, eos_token_id=eos, pad_token_id=cfg.pad_token_id,
                      length_penalty=length_penalty, early_stopping=early_stopping, forced_eos_token_id=forced_eos_token_id,
                      device=enc.device)
    w, b, scale = self.lm_head_params()
    bias = None if b is None else b.detach().reshape(-1).float().contiguous()
    t = 0
    while not state.done:
        x = self._decode_step(state.current_tokens().view(B * k, 1).contiguous(), t, caches, cross)
        logits = K.linear_fwd(x.contiguous(), ops.w16(w), bias, out_f32=True, alpha=scale)
        _, rows, _ = state.step(logits)
        caches = reorder_cache(caches, rows)
        t += 1
    return state.result()


@torch.no_grad()
def _greedy_decode_graph(self, enc, max_length, eos_token_id=None):
    """The whole KV-cached decode loop (max_length - 1 decoder passes, ~85 launches each) as ONE CUDA graph, cached
    per (shape, max_length, weight-cache epoch); the eos early exit becomes a host-side truncation of the result."""
    cfg = self.config
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    key = (tuple(enc.shape), enc.dtype, int(max_length), ops.CACHE.epoch, K.FP32_MODE)
    store = self.__dict__.setdefault("_decode_graphs", {})
    hit = store.get(key)
    if hit is None:
        static_enc = enc.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # lazy initialisation (working copies of the weights, bucket tables)
            self.greedy_decode(static_enc, max_length, eos_token_id=eos, sync_eos=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out_ids = self.greedy_decode(static_enc, max_length, eos_token_id=eos, sync_eos=False)
        store.clear()                          # one resident decode graph per model
        hit = store[key] = (graph, static_enc, out_ids)
    graph, static_enc, out_ids = hit
    static_enc.copy_(enc)
    graph.replay()
    ids = out_ids.clone()
    done = (ids[:, 1:] == eos).cumsum(1).clamp_(max=1).bool()        # [B, L-1]: eos seen at or before this position
    all_done = done.all(0)
    if bool(all_done.any()):
        ids = ids[:, :int(all_done.float().argmax()) + 2]
    return ids


Seq2SeqLM.greedy_decode_graph = _greedy_decode_graph
T5Seq2SeqLM.greedy_decode_graph = _greedy_decode_graph
Seq2SeqLM.beam_decode = _beam_decode
T5Seq2SeqLM.beam_decode = _beam_decode


def text_from_pretrained(path_or_config):
    from transformers import AutoConfig, PretrainedConfig

    def build(config):
        return T5Seq2SeqLM(config) if config.model_type == "t5" else Seq2SeqLM(config)

    if isinstance(path_or_config, PretrainedConfig):
        return build(path_or_config)
    path_or_config = resolve_checkpoint(path_or_config)
    config = AutoConfig.from_pretrained(path_or_config)
    model = build(config)
    sd = dict(load_checkpoint_state(path_or_config))
    # tied aliases may be stored once
    pre = "" if config.model_type == "t5" else "model."
    aliases = [pre + "shared.weight", pre + "encoder.embed_tokens.weight", pre + "decoder.embed_tokens.weight"]
    if getattr(config, "tie_word_embeddings", True):
        aliases.append("lm_head.weight")
    base = None
    for k in aliases:
        if k in sd:
            base = sd[k]
            break
    own = model.state_dict()
    for k in aliases:
        sd.setdefault(k, base)
    missing = [k for k in own if k not in sd]
    if missing:
        raise RuntimeError("checkpoint %s lacks %d tensors, e.g. %s" % (path_or_config, len(missing), missing[:3]))
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    return model
