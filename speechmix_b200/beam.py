"""Beam-search bookkeeping for ``SpeechMixEED.generate(num_beams=k)``.

The reference reaches beam search through HF's ``GenerationMixin`` (``ref:eval.py:12-13`` ``model.generate(...)`` with
the checkpoint's generation defaults; ``ref:speechmix/hf_model.py:304-338`` supplies ``prepare_inputs_for_generation``
and ``_reorder_cache``).  This module restates the search HF runs (hf:generation/utils.py ``_beam_search``: keep
2 x num_beams candidates per utterance so that finished hypotheses cannot starve the live beams, move a candidate to
the finished pool only if it ranks inside the first num_beams, length-normalise finished scores by
``generated_length ** length_penalty``, stop when no running beam can still beat the worst finished one) as a small
state machine over score tensors.  It is device-agnostic tensor logic -- no model code -- so it is checked on the CPU
against ``transformers``' own ``generate`` (tests/test_host_cpu.py) and drives the KV-cached decoder kernels on the GPU.
"""
import torch

NEG = -1.0e9


class BeamState:
    """One search over ``batch`` utterances.  Per decoder step call ``step(logits)`` with the fp32 next-token logits of
    the ``batch * num_beams`` running rows; it returns ``(tokens, rows, done)``: the token every running row continues
    with, the PREVIOUS row each new row descends from (flat index -- what the KV caches must be gathered by, the role of
    ``_reorder_cache``), and whether the search is over.  ``result()`` gives the best hypothesis per utterance."""

    def __init__(self, batch, num_beams, max_length, start_ids, eos_token_id=None, pad_token_id=None,
                 length_penalty=1.0, early_stopping=False, forced_eos_token_id=None, device="cpu"):
        self.B, self.k, self.L = int(batch), int(num_beams), int(max_length)
        self.eos = [] if eos_token_id is None else ([int(e) for e in eos_token_id] if isinstance(eos_token_id, (list, tuple))
                                                      else [int(eos_token_id)])
        self.eos = [e for e in self.eos if e >= 0]
        self.length_penalty, self.early_stopping = float(length_penalty), early_stopping
        # hf ForcedEOSTokenLogitsProcessor (on by default in BART-family generation configs): the token written at
        # position max_length - 1 can only be this one
        self.forced_eos = None if forced_eos_token_id is None else int(forced_eos_token_id)
        fill = ((pad_token_id or self.eos[0]) if self.eos else -1)   # hf: `pad or eos[0] if eos is not None else -1`
        dev = torch.device(device)
        self.keep = max(2, 1 + len(self.eos)) * self.k
        self.top_mask = torch.zeros(self.keep, dtype=torch.bool, device=dev)
        self.top_mask[:self.k] = True
        self.running = torch.full((self.B, self.k, self.L), int(fill), dtype=torch.long, device=dev)
        self.running[:, :, 0] = torch.as_tensor(start_ids, device=dev).view(-1, 1)
        self.finished = self.running.clone()
        self.run_scores = torch.zeros(self.B, self.k, device=dev)
        self.run_scores[:, 1:] = NEG                     # all beams start identical: only beam 0 may seed the search
        self.fin_scores = torch.full((self.B, self.k), NEG, device=dev)
        self.fin_len = torch.zeros(self.B, self.k, dtype=torch.long, device=dev)      # generated tokens of a finished row
        self.is_finished = torch.zeros(self.B, self.k, dtype=torch.bool, device=dev)
        self.can_improve = torch.ones(self.B, 1, dtype=torch.bool, device=dev)
        self.cur = 1                                     # tokens per running row so far (decoder prompt = 1 start token)
        self.done = False

    def current_tokens(self):
        """[batch * num_beams] last token of every running row (the decoder's next input)"""
        return self.running[:, :, self.cur - 1].reshape(-1)

    @staticmethod
    def _gather(t, idx):
        """t[b, idx[b, j], ...]"""
        while idx.dim() < t.dim():
            idx = idx.unsqueeze(-1)
        return torch.gather(t, 1, idx.expand(-1, -1, *t.shape[2:]))

    def step(self, logits):
        B, k, V = self.B, self.k, logits.shape[-1]
        logp = torch.log_softmax(logits.float(), dim=-1)
        if self.forced_eos is not None and self.cur == self.L - 1:
            forced = torch.full_like(logp, float("-inf"))
            forced[:, self.forced_eos] = 0.0
            logp = forced
        logp = logp.view(B, k, V) + self.run_scores[:, :, None]
        top_scores, top_idx = torch.topk(logp.view(B, k * V), self.keep)
        src = top_idx // V                                # beam each candidate extends
        tok = top_idx % V
        cand = self._gather(self.running, src)
        cand[:, :, self.cur] = tok
        new_len = self.cur + 1
        # a candidate stops here when it emitted eos or filled max_length
        stops = torch.zeros_like(tok, dtype=torch.bool)
        for e in self.eos:
            stops |= tok == e
        if new_len >= self.L:
            stops |= True
        # live beams of the next step: best num_beams candidates that did not stop
        live_scores = top_scores + stops.float() * NEG
        nxt = torch.topk(live_scores, k)[1]
        self.running = self._gather(cand, nxt)
        self.run_scores = self._gather(live_scores, nxt)
        rows = (self._gather(src, nxt) + torch.arange(B, device=src.device).view(-1, 1) * k).reshape(-1)
        # finished pool: a stopping candidate enters only from the first num_beams ranks, with its length-normalised score
        just = stops & self.top_mask[None, :]
        fs = top_scores / float((new_len - 1) ** self.length_penalty)
        full = torch.all(self.is_finished, dim=-1, keepdim=True) & (self.early_stopping is True)
        fs = fs + full.float() * NEG + (~self.can_improve).float() * NEG + (~just).float() * NEG
        m_seq = torch.cat((self.finished, cand), 1)
        m_sc = torch.cat((self.fin_scores, fs), 1)
        m_len = torch.cat((self.fin_len, torch.full_like(tok, new_len - 1)), 1)
        m_fin = torch.cat((self.is_finished, just), 1)
        best = torch.topk(m_sc, k)[1]
        self.finished, self.fin_scores = self._gather(m_seq, best), self._gather(m_sc, best)
        self.fin_len, self.is_finished = self._gather(m_len, best), self._gather(m_fin, best)
        self.cur = new_len
        # can a running beam still beat the worst finished hypothesis?
        if self.early_stopping == "never" and self.length_penalty > 0.0:
            hyp_len = self.L - 1
        else:
            hyp_len = self.cur - 1
        best_running = self.run_scores[:, :1] / float(hyp_len ** self.length_penalty)
        worst_fin = torch.where(self.is_finished, self.fin_scores.min(dim=1, keepdim=True)[0],
                                torch.full_like(self.fin_scores, NEG))
        self.can_improve = self.can_improve & torch.any(best_running > worst_fin, dim=-1, keepdim=True)
        open_beam = not (bool(torch.all(self.is_finished)) and self.early_stopping is True)
        self.done = not (bool(torch.any(self.can_improve)) and open_beam and not bool(torch.all(stops)))
        return self.current_tokens(), rows, self.done

    def result(self):
        """[batch, <= max_length] best finished hypothesis of every utterance, cropped to the longest one"""
        seq = self.finished[:, 0]
        n = 1 + int(self.fin_len[:, 0].max())
        return seq[:, :n]


def reorder_cache(caches, rows):
    """KV caches of the running rows after a beam step: new row i continues previous row ``rows[i]``
    (ref:speechmix/hf_model.py:337-338 ``_reorder_cache`` -> the decoder model's cache reorder)."""
    return [c.index_select(0, rows) for c in caches]
