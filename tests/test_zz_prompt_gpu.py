"""the decoder text prompt (ref:speechmix/hf_model.py:433-436) -- prompt token embeddings,
expanded over the batch, are prepended to the bridged speech states before the text encoder.  The oracle takes
already-tokenised ids (there is no real tokenizer offline); the product takes ids or a string."""
import pytest
import torch

from tests._cases import build_oracle, load_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["mini_eed_ds2", "mini_large_mbart"])
def test_decoder_text_prompt_matches_oracle(name, cuda_device):
    from tests.test_model_gpu import _mine_from, _rel
    fx = load_fixture(name)
    ora, x, labels = build_oracle(fx)
    ora.train(True)
    mine = _mine_from(ora, dict(fx, train_mode=True), cuda_device)
    prompt = torch.tensor([[5, 9, 4, 17, 6]])
    ref = ora(x, labels=labels, decoder_text_prompt_ids=prompt, keep_full_logits=True)
    out = mine(x.to(cuda_device), labels=labels.to(cuda_device), decoder_text_prompt=prompt.to(cuda_device))
    assert out["inputs_embeds"].shape[1] == ref["inputs_embeds"].shape[1] == 5 + ora(x, labels=labels)["inputs_embeds"].shape[1]
    assert abs(float(out["loss"]) - float(ref["loss"])) < 5e-3      # 16 target tokens (toy batch)
    assert _rel(out["inputs_embeds"], ref["inputs_embeds"]) < 2e-2
    assert _rel(out["encoder_last_hidden_state"], ref["encoder_last_hidden_state"]) < 4e-2
    assert _rel(mine.decoder_model.full_logits(out["decoder_last_hidden_state"]), ref["full_logits"]) < 2e-2
    ref["loss"].backward()
    out["loss"].backward()
    k = "decoder_model.model.shared.weight"       # the tied embedding also receives the prompt's gradient
    po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
    assert float((pm[k].grad.cpu() - po[k].grad).norm()) <= 5e-2 * float(po[k].grad.norm())
