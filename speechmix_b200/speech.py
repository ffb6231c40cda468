"""wav2vec2 / HuBERT speech encoder on the sm_100a kernels.

Module / parameter names mirror ``transformers``' ``Wav2Vec2Model`` / ``HubertModel``
(hf:models/wav2vec2/modeling_wav2vec2.py, hf:models/hubert/modeling_hubert.py) so that a
reference ``state_dict`` loads unchanged; the ``torch.nn`` leaf modules are used as
parameter containers only -- all arithmetic goes through ``ops.py`` -> libspeechmix_sm100.
"""
import os

import torch
from torch import nn

from . import ops


class SpeechOutput(dict):
    """Mapping with attribute access (stands in for transformers' BaseModelOutput)."""

    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class _ConvLayer(nn.Module):
    def __init__(self, cin, cout, k, stride, bias, norm):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, stride=stride, bias=bias)
        if norm == "group":
            self.layer_norm = nn.GroupNorm(num_groups=cout, num_channels=cout, affine=True)
        elif norm == "layer":
            self.layer_norm = nn.LayerNorm(cout, elementwise_affine=True)


class FeatureEncoder(nn.Module):
    """hf:...wav2vec2.py:382-419"""

    def __init__(self, config):
        super().__init__()
        self.config = config
        n = config.num_feat_extract_layers
        dims, ks, ss = list(config.conv_dim), list(config.conv_kernel), list(config.conv_stride)
        layers = []
        for i in range(n):
            if config.feat_extract_norm == "group":
                norm = "group" if i == 0 else None
            elif config.feat_extract_norm == "layer":
                norm = "layer"
            else:
                raise ValueError(config.feat_extract_norm)
            layers.append(_ConvLayer(1 if i == 0 else dims[i - 1], dims[i], ks[i], ss[i], config.conv_bias, norm))
        self.conv_layers = nn.ModuleList(layers)
        self._ks, self._ss = ks, ss

    def forward(self, input_values):
        cfg = self.config
        if cfg.feat_extract_activation != "gelu":
            raise NotImplementedError("only GELU feature encoders are supported")
        if self._ks[0] != 10 or self._ss[0] != 5 or any(s != 2 for s in self._ss[1:]) or \
                any(k not in (2, 3) for k in self._ks[1:]):
            raise NotImplementedError("unsupported conv feature-encoder geometry %r / %r" % (self._ks, self._ss))
        if cfg.feat_extract_norm == "layer":
            if not cfg.conv_bias:
                raise NotImplementedError("feat_extract_norm='layer' is only wired for conv_bias=True (the HF large presets)")
            flat = []
            for l in self.conv_layers:
                flat += [l.conv.weight, l.conv.bias, l.layer_norm.weight, l.layer_norm.bias]
            return ops.FeatureEncoderLayerNormFn.apply(input_values, tuple(self._ks[1:]), *flat)
        if cfg.conv_bias:
            raise NotImplementedError("feat_extract_norm='group' with conv_bias=True is not wired to the kernels")
        l0 = self.conv_layers[0]
        ws = [l.conv.weight for l in self.conv_layers[1:]]
        return ops.FeatureEncoderGroupFn.apply(input_values, tuple(self._ks[1:]), l0.conv.weight, l0.layer_norm.weight,
                                               l0.layer_norm.bias, *ws)


class FeatureProjection(nn.Module):
    """hf:...wav2vec2.py:422-434 / hf:...hubert.py:216-232"""

    def __init__(self, config, with_layer_norm=True):
        super().__init__()
        self.with_layer_norm = with_layer_norm
        if with_layer_norm:
            self.layer_norm = nn.LayerNorm(config.conv_dim[-1], eps=config.layer_norm_eps)
        self.projection = nn.Linear(config.conv_dim[-1], config.hidden_size)
        self.eps = config.layer_norm_eps
        self.p_drop = float(getattr(config, "feat_proj_dropout", 0.0) or 0.0)

    def forward(self, x):
        if self.with_layer_norm:
            x = ops.layer_norm(x, self.layer_norm.weight, self.layer_norm.bias, self.eps)
        x = ops.linear(x, self.projection.weight, self.projection.bias)
        return ops.dropout(x, self.p_drop, self.training)      # hf:...wav2vec2.py:432 (feat_proj_dropout)


class PositionalConvEmbedding(nn.Module):
    """hf:...wav2vec2.py:326-379; the weight-norm factors keep the reference's parameter names
    (conv.parametrizations.weight.original0/1)."""

    def __init__(self, config):
        super().__init__()
        conv = nn.Conv1d(config.hidden_size, config.hidden_size, kernel_size=config.num_conv_pos_embeddings,
                         padding=config.num_conv_pos_embeddings // 2, groups=config.num_conv_pos_embedding_groups)
        self.conv = nn.utils.parametrizations.weight_norm(conv, name="weight", dim=2)
        self.groups = config.num_conv_pos_embedding_groups
        if config.num_conv_pos_embeddings % 2 != 0:
            raise NotImplementedError("odd positional-conv kernels are not supported")

    def forward(self, x):
        """returns x + GELU(conv(x)) (the residual add is fused into the kernel epilogue)."""
        p = self.conv.parametrizations.weight
        weight = ops.WeightNormFn.apply(p.original0, p.original1)  # g * v / ||v|| per tap
        return ops.PosConvFn.apply(x, weight, self.conv.bias, self.groups, (p.original0, p.original1))


class _Attention(nn.Module):
    def __init__(self, dim, bias=True):
        super().__init__()
        # registration order follows hf BartAttention / Wav2Vec2Attention (k, v, q, out) so that
        # named_parameters() -- and with it list_grad / list_no_grad -- enumerate like the reference
        self.k_proj = nn.Linear(dim, dim, bias=bias)
        self.v_proj = nn.Linear(dim, dim, bias=bias)
        self.q_proj = nn.Linear(dim, dim, bias=bias)
        self.out_proj = nn.Linear(dim, dim, bias=bias)

    def params(self):
        return (self.q_proj.weight, self.q_proj.bias, self.k_proj.weight, self.k_proj.bias, self.v_proj.weight,
                self.v_proj.bias, self.out_proj.weight, self.out_proj.bias)


class _FeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.intermediate_dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.output_dense = nn.Linear(config.intermediate_size, config.hidden_size)


class EncoderLayer(nn.Module):
    """hf:...wav2vec2.py:576-609 (post-LN) and :612-655 (stable / pre-LN)."""

    def __init__(self, config):
        super().__init__()
        self.attention = _Attention(config.hidden_size)
        self.layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.feed_forward = _FeedForward(config)
        self.final_layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.cfg = dict(heads=config.num_attention_heads, causal=False, pre_ln=bool(config.do_stable_layer_norm),
                        eps=config.layer_norm_eps, act=config.hidden_act)
        # train-mode dropout (hf:...wav2vec2.py:466-573): attention probabilities / block outputs / FFN activation
        self.drop = dict(p_attn=float(config.attention_dropout), p_hidden=float(config.hidden_dropout),
                         p_act=float(config.activation_dropout))
        if config.hidden_size // config.num_attention_heads != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")

    def forward(self, x, kv_len=None):
        base = dict(self.cfg, **self.drop) if (self.training and any(v > 0 for v in self.drop.values())) else self.cfg
        cfg = base if kv_len is None else dict(base, kv_len=kv_len)
        x = ops.AttnBlockFn.apply(x, None, cfg, *self.attention.params(), self.layer_norm.weight,
                                  self.layer_norm.bias)
        ff = self.feed_forward
        return ops.FFNBlockFn.apply(x, base, ff.intermediate_dense.weight, ff.intermediate_dense.bias,
                                    ff.output_dense.weight, ff.output_dense.bias, self.final_layer_norm.weight,
                                    self.final_layer_norm.bias)


class Encoder(nn.Module):
    """hf:...wav2vec2.py:658-803"""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.pos_conv_embed = PositionalConvEmbedding(config)
        self.layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.layers = nn.ModuleList([EncoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.stable = bool(config.do_stable_layer_norm)

    def forward(self, x, output_hidden_states=False, frame_len=None):
        """frame_len: optional int32 [B] valid-frame counts (hf :669-681): padded frames are zeroed before the
        positional conv and masked out as attention keys in every layer."""
        hs = []
        if frame_len is not None:
            x = ops.mask_rows(x, frame_len)
        x = self.pos_conv_embed(x)
        if not self.stable:
            x = ops.layer_norm(x, self.layer_norm.weight, self.layer_norm.bias, self.config.layer_norm_eps)
        x = ops.dropout(x, float(self.config.hidden_dropout), self.training)   # hf:...wav2vec2.py:695 / :770
        for layer in self.layers:
            if output_hidden_states:
                hs.append(x)
            # LayerDrop (hf :702-704) is a training-time regulariser of the reference recipe; it is
            # honoured on the host exactly like the reference does.
            if self.training and self.config.layerdrop > 0 and float(torch.rand([])) < self.config.layerdrop:
                continue
            x = layer(x, kv_len=frame_len)
        if self.stable:
            x = ops.layer_norm(x, self.layer_norm.weight, self.layer_norm.bias, self.config.layer_norm_eps)
        if output_hidden_states:
            hs.append(x)
        return x, tuple(hs)


class SpeechEncoderModel(nn.Module):
    """Drop-in for ``Wav2Vec2Model`` / ``HubertModel`` on the SpeechMix path
    (``encoder_model(input_values, output_hidden_states=True)``, ref:speechmix/hf_model.py:397)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.feature_extractor = FeatureEncoder(config)
        is_hubert = config.model_type == "hubert"
        self.feature_projection = FeatureProjection(
            config, with_layer_norm=(not is_hubert) or bool(getattr(config, "feat_proj_layer_norm", True)))
        if config.mask_time_prob > 0.0 or config.mask_feature_prob > 0.0:  # hf:...wav2vec2.py:1262-1264
            self.masked_spec_embed = nn.Parameter(torch.empty(config.hidden_size).uniform_())
        self.encoder = Encoder(config)
        if getattr(config, "add_adapter", False):
            raise NotImplementedError("wav2vec2 adapter stacks are out of scope")

    @property
    def device(self):
        return next(self.parameters()).device

    def feat_extract_output_lengths(self, input_lengths):
        """hf:...wav2vec2.py:1005-1018: frames left by the conv stack for each sample count."""
        n = torch.as_tensor(input_lengths).to(torch.long)
        for k, s in zip(self.config.conv_kernel, self.config.conv_stride):
            n = torch.div(n - k, s, rounding_mode="floor") + 1
        return n

    def _mask_hidden_states(self, x, frame_len=None):
        """SpecAugment, hf:...wav2vec2.py:1280-1324 (hubert: same code): training only, ``apply_spec_augment`` and a
        positive ``mask_time_prob`` / ``mask_feature_prob``.  The span indices are drawn on the host by the SAME
        function the reference's backbone calls (transformers' ``_compute_mask_indices``, numpy's global RNG), so a
        seeded run masks the same frames as the reference; the replacement itself is a CUDA kernel."""
        cfg = self.config
        if not self.training or not getattr(cfg, "apply_spec_augment", True):
            return x
        if cfg.mask_time_prob <= 0 and cfg.mask_feature_prob <= 0:
            return x
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("SpecAugment draws a new mask on the host every step and cannot be captured in a CUDA "
                               "graph; set apply_spec_augment=False for graphed steps")
        from transformers.models.wav2vec2.modeling_wav2vec2 import _compute_mask_indices
        B, T, H = x.shape
        time_mask = feat_mask = None
        if cfg.mask_time_prob > 0:
            frames = None
            if frame_len is not None and cfg.model_type != "hubert":   # HubertModel draws over all T frames (hf hubert :1020)
                frames = (torch.arange(T)[None, :] < frame_len.cpu()[:, None]).long()
            idx = _compute_mask_indices((B, T), mask_prob=cfg.mask_time_prob, mask_length=cfg.mask_time_length,
                                        attention_mask=frames, min_masks=cfg.mask_time_min_masks)
            time_mask = torch.from_numpy(idx).to(device=x.device, dtype=torch.uint8)
        if cfg.mask_feature_prob > 0:
            idx = _compute_mask_indices((B, H), mask_prob=cfg.mask_feature_prob, mask_length=cfg.mask_feature_length,
                                        min_masks=cfg.mask_feature_min_masks)
            feat_mask = torch.from_numpy(idx).to(device=x.device, dtype=torch.uint8)
        return ops.spec_augment(x, self.masked_spec_embed, time_mask, feat_mask)

    def forward(self, input_values, attention_mask=None, output_hidden_states=False, **kwargs):
        """attention_mask ([B, n] 1 = audio sample, 0 = padding; the reference itself never passes one --
        SURVEY section 8f row 1): HF semantics, hf:...wav2vec2.py:1026-1044, :669-681 -- the conv stack runs over the
        padded signal, frames past each sample's own length are zeroed after the projection and masked as keys.
        (HF advises against it for feat_extract_norm="group" checkpoints, whose GroupNorm sees the padding.)"""
        x = self.feature_extractor(input_values)          # [B, T, C] channels-last (HF: [B, C, T] + transpose)
        x = self.feature_projection(x)
        frame_len = None
        if attention_mask is not None:
            if ops.K.FP32_MODE:
                raise NotImplementedError("fp32 verification mode has no key-padding mask")
            lens = self.feat_extract_output_lengths(attention_mask.to(torch.long).sum(-1))
            frame_len = lens.clamp(1, x.shape[1]).to(device=x.device, dtype=torch.int32)
        x = self._mask_hidden_states(x, frame_len)
        x, hs = self.encoder(x, output_hidden_states=output_hidden_states, frame_len=frame_len)
        return SpeechOutput(last_hidden_state=x, hidden_states=hs if output_hidden_states else None)


# ---------------------------------------------------------------------------
def resolve_checkpoint(name_or_path):
    """Local checkpoint directory for ``name_or_path``: a directory is returned as is; anything else is treated as a hub
    id the way the reference's ``from_pretrained`` calls do (ref:speechmix/hf_model.py:206-220, ref:eval.py, ref:train.py)
    and resolved through the local hub cache / ``snapshot_download`` (which honours TRANSFORMERS_OFFLINE / HF_HUB_OFFLINE)."""
    if os.path.isdir(name_or_path):
        return name_or_path
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(name_or_path, allow_patterns=["*.json", "*.safetensors", "*.bin", "*.model", "*.txt"])
    except Exception as e:
        raise FileNotFoundError("%r is neither a local checkpoint directory nor a hub id available in the local cache "
                                "(%s: %s)" % (name_or_path, type(e).__name__, e)) from e


def load_checkpoint_state(path):
    """state dict of a transformers checkpoint directory (safetensors or .bin)."""
    path = resolve_checkpoint(path)
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    pt = os.path.join(path, "pytorch_model.bin")
    if os.path.exists(pt):
        return torch.load(pt, map_location="cpu")
    raise FileNotFoundError("no model.safetensors / pytorch_model.bin under %s" % path)


def speech_from_pretrained(path_or_config):
    """Build from a local transformers checkpoint directory (config + weights) or a config object
    (random init, HF initialisers are not reproduced -- load a state dict afterwards)."""
    from transformers import AutoConfig, PretrainedConfig

    if isinstance(path_or_config, PretrainedConfig):
        return SpeechEncoderModel(path_or_config)
    path_or_config = resolve_checkpoint(path_or_config)
    config = AutoConfig.from_pretrained(path_or_config)
    model = SpeechEncoderModel(config)
    sd = load_checkpoint_state(path_or_config)
    fixed = {}
    for k, v in sd.items():
        for prefix in ("wav2vec2.", "hubert."):
            if k.startswith(prefix):
                k = k[len(prefix):]
        k = k.replace("pos_conv_embed.conv.weight_g", "pos_conv_embed.conv.parametrizations.weight.original0")
        k = k.replace("pos_conv_embed.conv.weight_v", "pos_conv_embed.conv.parametrizations.weight.original1")
        fixed[k] = v
    own = model.state_dict()
    missing = [k for k in own if k not in fixed]
    model.load_state_dict({k: v for k, v in fixed.items() if k in own}, strict=False)
    if missing:
        raise RuntimeError("checkpoint %s lacks %d tensors, e.g. %s" % (path_or_config, len(missing), missing[:3]))
    return model
