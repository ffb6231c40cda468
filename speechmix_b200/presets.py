"""Named backbone shapes (transformers config objects, no weights) for synthetic runs: the configurations
BASELINE.json quotes its metric on.  Random-init models of these shapes are what
bench.py times; checkpoints load through ``speech_from_pretrained`` / ``text_from_pretrained`` instead."""


def _deterministic(cfg, keys):
    for k in keys:
        if hasattr(cfg, k):
            setattr(cfg, k, 0.0)
    return cfg


def speech_config(kind="base", model_type="wav2vec2", deterministic=True):
    from transformers import HubertConfig, Wav2Vec2Config
    cls = HubertConfig if model_type == "hubert" else Wav2Vec2Config
    if kind == "base":            # wav2vec2-base / hubert-base: H768, L12, FF3072, group-norm conv stack, post-LN
        cfg = cls()
    elif kind == "large":         # hubert-large / wav2vec2-large-lv60: layer-norm conv stack, stable LN
        cfg = cls(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)
    elif kind == "large_group":   # original wav2vec2-large
        cfg = cls(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
    else:
        raise ValueError(kind)
    if deterministic:             # both bench arms run with dropout / LayerDrop / SpecAugment zeroed
        _deterministic(cfg, ("hidden_dropout", "activation_dropout", "attention_dropout", "feat_proj_dropout", "layerdrop",
                             "mask_time_prob", "mask_feature_prob", "final_dropout", "feat_quantizer_dropout"))
        cfg.apply_spec_augment = False
    return cfg


def text_config(kind="bart-base", deterministic=True):
    from transformers import BartConfig, MBartConfig, T5Config
    if kind == "bart-base":
        cfg = BartConfig(d_model=768, encoder_layers=6, decoder_layers=6, encoder_attention_heads=12,
                         decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072, vocab_size=50265)
    elif kind == "bart-large":
        cfg = BartConfig()
    elif kind == "mbart-large-50":
        cfg = MBartConfig(vocab_size=250054, scale_embedding=True, d_model=1024, encoder_layers=12, decoder_layers=12,
                          encoder_attention_heads=16, decoder_attention_heads=16, encoder_ffn_dim=4096,
                          decoder_ffn_dim=4096, decoder_start_token_id=2)
    elif kind == "t5-base":
        cfg = T5Config(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, vocab_size=32128,
                       feed_forward_proj="relu", decoder_start_token_id=0)
    else:
        raise ValueError(kind)
    if deterministic:
        _deterministic(cfg, ("dropout", "attention_dropout", "activation_dropout", "encoder_layerdrop", "decoder_layerdrop",
                             "classifier_dropout", "dropout_rate"))
    return cfg
