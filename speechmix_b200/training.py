"""Host-side pieces either side of the hot path (SURVEY.md section 8f "next" rows): the batch / collation
contract of ref:train.py:90-133, the text-teacher target construction of ref:train.py:18-34 on the KV-cached
decoder, and the gradual-unfreeze policy of ref:speechmix/module/utility.py:6-34 without the HF Trainer.
Nothing here launches a kernel except ``create_self_decoder_input`` (through ``Seq2SeqLM.greedy_decode``)."""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch.nn.utils.rnn import pad_sequence


@dataclass
class DataCollatorWithPadding:
    """ref:train.py:90-133.  Audio is padded with the VALUE -100 and no mask is built (the reference's contract:
    SURVEY.md headline fact 4), label padding becomes -100 (ignored by the CE), a leading bos shared by every row
    is cut because ``shift_tokens_right`` prepends the start token anyway.  ``tokenizer`` is only used for its
    pad / bos ids; pre-tokenised features work with ``pad_token_id`` / ``bos_token_id`` given directly."""

    tokenizer: Optional[object] = None
    pad_token_id: Optional[int] = None
    bos_token_id: Optional[int] = None
    pad_to_multiple_of_labels: Optional[int] = None
    max_length_labels: Optional[int] = None
    # extension, off by default = the reference's contract: also emit ``attention_mask`` (1 = audio
    # sample) and pad the audio with 0.0 the way HF's feature extractor does, for ``forward(..., attention_mask=...)``
    return_attention_mask: bool = False

    def _ids(self):
        pad = self.pad_token_id if self.pad_token_id is not None else getattr(self.tokenizer, "pad_token_id", None)
        bos = self.bos_token_id if self.bos_token_id is not None else getattr(self.tokenizer, "bos_token_id", None)
        if pad is None:
            raise ValueError("DataCollatorWithPadding needs a pad token id")
        return pad, bos

    def _pad_ids(self, rows: Sequence[Sequence[int]], pad: int):
        n = max(len(r) for r in rows)     # tokenizer.pad(padding=True) pads to the longest row (max_length unused)
        if self.pad_to_multiple_of_labels:
            m = self.pad_to_multiple_of_labels
            n = (n + m - 1) // m * m
        ids = torch.full((len(rows), n), pad, dtype=torch.long)
        mask = torch.zeros((len(rows), n), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.as_tensor(list(r), dtype=torch.long)
            mask[i, :len(r)] = 1
        return ids, mask

    def __call__(self, features: List[Dict]) -> Dict[str, torch.Tensor]:
        pad, bos = self._ids()
        audio = [torch.as_tensor(f["input_values"], dtype=torch.float32) for f in features]
        batch = {"input_values": pad_sequence(audio, batch_first=True,
                                              padding_value=0.0 if self.return_attention_mask else -100)}
        if self.return_attention_mask:
            batch["attention_mask"] = pad_sequence([torch.ones(len(a), dtype=torch.long) for a in audio], batch_first=True)
        labels, mask = self._pad_ids([f["labels"] for f in features], pad)
        if "text_input_ids" in features[0]:
            batch["text_input_ids"], _ = self._pad_ids([f["text_input_ids"] for f in features], pad)
        labels = labels.masked_fill(mask.ne(1), -100)
        if bos and bool((labels[:, 0] == bos).all()):
            labels = labels[:, 1:]
        batch["labels"] = labels
        return batch


@torch.no_grad()
def create_self_decoder_input(decoder_model, gen_input: Sequence[int], device=None, max_length: Optional[int] = None):
    """ref:train.py:18-34: greedy text-teacher targets -- the frozen text model decodes its own continuation of the
    (already tokenised) input sentence; returns ``(gen_input, predicted)`` without the start token and the eos.
    The reference re-runs the whole model for every token; here the text encoder runs once and the decoder is
    KV-cached (``Seq2SeqLM.greedy_decode``)."""
    cfg = decoder_model.config
    device = decoder_model.device if device is None else device
    ids = torch.as_tensor([list(gen_input)], dtype=torch.long, device=device)
    steps = max(int(getattr(cfg, "max_length", 20) or 20), len(gen_input)) if max_length is None else max_length
    was_training = decoder_model.training
    decoder_model.eval()
    enc, _ = decoder_model.encode(input_ids=ids)
    out = decoder_model.greedy_decode(enc, steps + 1, eos_token_id=cfg.eos_token_id)[0].tolist()
    decoder_model.train(was_training)
    pred = out[1:]
    if cfg.eos_token_id in pred:
        pred = pred[:pred.index(cfg.eos_token_id)]
    return list(gen_input), pred


class FreezingPolicy:
    """ref:speechmix/module/utility.py:6-34 (FreezingCallback) as a plain object: call ``on_epoch_begin(epoch)``.
    During the first ``freeze_epoch`` epochs only the last ``epoch * n_params / freeze_epoch`` parameters (in
    registration order) of ``freeze_model`` keep their original ``requires_grad``; afterwards all are restored.
    Which weight-gradient GEMMs run follows from ``requires_grad`` automatically.  A data-parallel
    ``parallel.GradientAllReducer`` fixes its hooked parameter set when it is built: pass it as ``reducer`` and the
    policy rebuilds its buckets whenever the trainable set changes (otherwise newly unfrozen parameters would never be
    averaged and the replicas would drift apart)."""

    def __init__(self, freeze_model, freeze_epoch=3, reducer=None, reducer_module=None):
        self.freeze_model, self.freeze_epoch = freeze_model, freeze_epoch
        self.reducer, self.reducer_module = reducer, reducer_module if reducer_module is not None else freeze_model
        self.default = {n: p.requires_grad for n, p in freeze_model.named_parameters()}
        self.names = list(self.default)
        self.freeze_layers = int(len(self.names) / freeze_epoch)

    def on_epoch_begin(self, epoch):
        before = [p.requires_grad for p in self.freeze_model.parameters()]
        self._apply(epoch)
        if self.reducer is not None and before != [p.requires_grad for p in self.freeze_model.parameters()]:
            self.reducer.rebuild(self.reducer_module)

    def _apply(self, epoch):
        if epoch < self.freeze_epoch:
            k = int(self.freeze_layers * epoch)
            release = set(self.names[-k:]) if k > 0 else set(self.names)   # names[-0:] is the whole list, as in the reference
            for n, p in self.freeze_model.named_parameters():
                p.requires_grad = self.default[n] if n in release else False
        else:
            for n, p in self.freeze_model.named_parameters():
                p.requires_grad = self.default[n]

    def on_save(self, model):
        for _, p in model.named_parameters():
            p.requires_grad = True


class TrainStep:
    """The optimizer-step policy of the reference recipe (ref:train.py:291-311 -> HF ``TrainingArguments``):
    ``gradient_accumulation_steps`` micro-batches per update (loss scaled by 1 / steps, as the Trainer does), global
    gradient-norm clipping (``max_grad_norm``, Trainer default 1.0) and one optimizer step; under data parallelism the
    micro-steps before the last run inside ``reducer.no_sync()`` so that ONE bucketed all-reduce carries the accumulated
    gradients.  ``model(**batch)`` must return a mapping with ``loss`` (every SpeechMix class does)."""

    def __init__(self, model, optimizer, grad_accum=1, max_grad_norm=1.0, reducer=None):
        self.model, self.opt, self.reducer = model, optimizer, reducer
        self.grad_accum, self.max_grad_norm = max(int(grad_accum), 1), max_grad_norm
        self.params = [p for g in optimizer.param_groups for p in g["params"]]

    def __call__(self, micro_batches):
        """micro_batches: sequence of ``grad_accum`` keyword dicts; returns (mean loss, gradient norm before clipping)."""
        assert len(micro_batches) == self.grad_accum, "one update = grad_accum micro-batches"
        self.opt.zero_grad(set_to_none=True)
        total = 0.0
        for i, batch in enumerate(micro_batches):
            last = i + 1 == self.grad_accum
            ctx = self.reducer.no_sync() if (self.reducer is not None and not last) else _NullCtx()
            with ctx:
                loss = self.model(**batch)["loss"] / self.grad_accum
                loss.backward()
            total = total + loss.detach()
        if self.reducer is not None:
            self.reducer.finish()
        norm = None
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            norm = torch.nn.utils.clip_grad_norm_([p for p in self.params if p.grad is not None], self.max_grad_norm)
        self.opt.step()
        return total, norm


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
