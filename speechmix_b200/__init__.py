"""speechmix_b200 -- B200-native (sm_100a) implementation of the SpeechMix speech-to-text hot path.

    from speechmix_b200 import SpeechMixEED
    model = SpeechMixEED(speech_ckpt_dir, text_ckpt_dir, down_scale=2).cuda()
    out = model(input_values, labels=labels)      # out["loss"], out["logits"] (argmax ids)

The CUDA library must be built first (``python -m speechmix_b200.build``); there is no fallback.
"""
from .model import (HFSpeechMixAdapter, HFSpeechMixED, HFSpeechMixEED, HFSpeechMixFixed, HFSpeechMixGAN,  # noqa: F401
                    HFSpeechMixSelf, SpeechMixAdapter, SpeechMixConfig, SpeechMixED, SpeechMixEED, SpeechMixFixed, SpeechMixGAN, SpeechMixSelf,
                    handle_decoder_input_none, shift_tokens_right)
from .optim import FusedAdafactor  # noqa: F401

__all__ = ["SpeechMixEED", "SpeechMixED", "HFSpeechMixED", "SpeechMixFixed", "SpeechMixAdapter", "SpeechMixSelf", "SpeechMixGAN", "HFSpeechMixGAN", "HFSpeechMixEED", "HFSpeechMixFixed",
           "HFSpeechMixAdapter", "HFSpeechMixSelf", "SpeechMixConfig", "shift_tokens_right",
           "handle_decoder_input_none", "FusedAdafactor"]
