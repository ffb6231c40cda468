"""Train-mode dropout (the regulariser of the reference recipe: the HF modules run under ``.train()`` with
hidden / attention / activation dropout 0.1 -- ref:speechmix/hf_model.py:397,357-365, ref:train.py:291-330).

The kernels draw counter-based masks (seed, step, call, element); the parity test exports exactly those masks and feeds
them to the CPU oracle by patching ``torch.nn.functional.dropout`` (every HF dropout site goes through it, in the same
order as our call indices), so loss, logits and ALL gradients can be compared like in the deterministic tests."""
import pytest
import torch

from tests._cases import load_fixture

pytestmark = pytest.mark.gpu


def test_dropout_kernel_statistics_and_determinism(cuda_device):
    from speechmix_b200 import kernels as K
    from speechmix_b200 import ops
    st = torch.tensor([1234, 7], dtype=torch.int64, device=cuda_device)
    x = torch.ones(4096, 1024, device=cuda_device, dtype=torch.bfloat16)
    for p in (0.1, 0.5):
        y = K.dropout(x, st, 3, p)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 2e-3, (p, keep)                       # 4 M samples: 3 sigma ~ 5e-4
        assert abs(y.float().mean().item() - 1.0) < 6e-3                    # unbiased (bf16 rounding of 1 / (1 - p))
        assert torch.equal(y, K.dropout(x, st, 3, p))                       # pure function of (state, call, element)
        assert not torch.equal(y, K.dropout(x, st, 4, p))                   # another site
        st2 = st.clone()
        st2[1] += 1
        assert not torch.equal(y, K.dropout(x, st2, 3, p))                  # another step
        m = K.dropout_mask(x.shape, st, 3, p)
        assert torch.equal(m.bool(), y != 0)                                # the exported mask is the applied mask
        # rows / columns are not correlated: every row and every column keeps ~ (1 - p)
        assert (m.float().mean(0) - (1 - p)).abs().max() < 0.05 and (m.float().mean(1) - (1 - p)).abs().max() < 0.08
    # autograd: the backward pass regenerates the same mask; residual and activation-gradient variants
    ops.DROPOUT.manual_seed(5)
    ops.DROPOUT.begin_step(cuda_device)
    xr = torch.randn(64, 256, device=cuda_device).to(torch.bfloat16).requires_grad_(True)
    y = ops.dropout(xr, 0.3, True)
    y.backward(torch.ones_like(y))
    assert torch.equal(xr.grad != 0, y != 0)
    res = torch.randn(64, 256, device=cuda_device).to(torch.bfloat16)
    aux = torch.randn(64, 256, device=cuda_device).to(torch.bfloat16)
    out, a1 = K.dropout(xr.detach(), st, 9, 0.25, residual=res, aux_in=aux, aux_mode=1)
    mk = K.dropout_mask(xr.shape, st, 9, 0.25).bool()
    ref = res.float() + torch.where(mk, xr.detach().float() / 0.75, torch.zeros_like(res, dtype=torch.float32))
    assert (out.float() - ref).abs().max() < 0.05
    assert torch.equal(a1 != 0, mk & (aux != 0))
    _, a2 = K.dropout(xr.detach(), st, 9, 0.25, aux_in=aux, aux_mode=2)
    assert torch.equal(a2 != 0, mk & (aux > 0)) and abs(float(a2.max()) - 1 / 0.75) < 0.01


@pytest.mark.parametrize("shape", [(300, 768), (77, 256), (1000, 1024), (64, 96)])
def test_fused_dropout_layernorm_tail_equals_the_separate_launches(shape, cuda_device):
    """smx_layernorm_dropout_fwd / _bwd (dropout + residual add + LayerNorm of a post-LN block as one launch each way)
    against the three / four separate launches they replace, same (state, call, p): the stored sum and both gradient
    tensors must be the same (same arithmetic, same rounding points, same mask), the normalised output / statistics / parameter
    gradients / column sums equal up to summation order."""
    from speechmix_b200 import kernels as K
    R, C = shape
    g = torch.Generator(device=cuda_device).manual_seed(R + C)
    x, res, dy = (torch.randn(R, C, device=cuda_device, generator=g).to(torch.bfloat16) for _ in range(3))
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_device, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_device, generator=g)
    st = torch.tensor([77, 5], dtype=torch.int64, device=cuda_device)
    call, p = 21, 0.1
    s_ref = K.dropout(x, st, call, p, residual=res)
    y_ref, _, mean_ref, rstd_ref = K.layernorm_fwd(s_ref, gamma, beta, 1e-5)
    y, s, mean, rstd = K.layernorm_dropout_fwd(x, res, gamma, beta, 1e-5, st, call, p)
    def same(a, b, what):      # identical up to FMA-contraction choices of the compiler: <= 0.1 % of elements, one bf16 ulp
        diff = (a.float() - b.float()).abs()
        assert float((diff > 0).float().mean()) < 1e-3 and float((diff / (b.float().abs() + 1e-3)).max()) < 2 ** -6, what
        assert torch.equal(a == 0, b == 0), what                              # the same mask
    same(s, s_ref, "sum")
    assert float((mean - mean_ref).abs().max()) < 1e-5 and float((rstd / rstd_ref - 1).abs().max()) < 1e-5
    assert float((y.float() - y_ref.float()).abs().max()) <= 2 ** -6          # one bf16 ulp at |y| < 4
    dx_ref, dg_ref, db_ref, _ = K.layernorm_bwd(dy, s_ref, gamma, mean_ref, rstd_ref, want_colsum=True)
    dxd_ref = K.dropout(dx_ref, st, call, p)
    cs_ref = K.colsum(dxd_ref)
    dx, dxd, dg, db, cs = K.layernorm_dropout_bwd(dy, s_ref, gamma, mean_ref, rstd_ref, st, call, p)
    same(dx, dx_ref, "dx")
    same(dxd, dxd_ref, "dx_drop")
    for name, a, b in (("dgamma", dg, dg_ref), ("dbeta", db, db_ref), ("colsum", cs, cs_ref)):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-5, name
    dx2, dxd2, dg2, db2, cs2 = K.layernorm_dropout_bwd(dy, s_ref, gamma, mean_ref, rstd_ref, st, call, p, want_dbeta=False,
                                                       want_colsum=False)
    assert db2 is None and cs2 is None and torch.equal(dx2, dx) and torch.equal(dxd2, dxd)


def _attention_mask_reference(q, k, v, heads, scale, mask, p, causal):
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    qh, kh, vh = (t.float().view(B, -1, heads, 64).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        s = s.masked_fill(~torch.ones(Tq, Tk, device=q.device, dtype=torch.bool).tril(Tk - Tq), float("-inf"))
    pr = torch.softmax(s, -1) * mask.float() / (1 - p)
    return (pr @ vh).transpose(1, 2).reshape(B, Tq, heads * 64)


@pytest.mark.parametrize("shape", [(2, 300, 300, 2, False), (2, 64, 200, 2, False), (2, 96, 96, 2, True), (1, 749, 749, 2, False)])
def test_attention_probability_dropout_forward_backward(shape, cuda_device):
    """dropout on the attention probabilities inside the flash kernels (both forward kernels, both backward kernels):
    against an fp32 reference that applies the exported mask."""
    from speechmix_b200 import kernels as K
    B, Tq, Tk, H, causal = shape
    g = torch.Generator(device=cuda_device).manual_seed(0)
    q, k, v, do = (torch.randn(B, t, H * 64, device=cuda_device, generator=g).mul(0.7).to(torch.bfloat16)
                   for t in (Tq, Tk, Tk, Tq))
    st = torch.tensor([99, 3], dtype=torch.int64, device=cuda_device)
    p = 0.2
    drop = (st, 11, p)
    o, lse = K.attn_fwd(q, k, v, H, causal=causal, scale=0.125, dropout=drop)
    dq, dk, dv = K.attn_bwd(do, q, k, v, o, lse, H, causal=causal, scale=0.125, dropout=drop)
    mask = K.dropout_mask((B, H, Tq, Tk), st, 11, p, attention=True)
    assert abs(mask.float().mean().item() - (1 - p)) < 0.01
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _attention_mask_reference(qr, kr, vr, H, 0.125, mask, p, causal)
    ref.backward(do.float())
    for name, got, want in (("o", o, ref.detach()), ("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        err = float((got.float() - want).abs().max() / (want.abs().max() + 1e-9))
        assert err < 2e-2, (name, err)
    o2, _ = K.attn_fwd(q, k, v, H, causal=causal, scale=0.125)
    assert float((o2.float() - o.float()).abs().max()) > 1e-2          # the mask matters


class _MaskFeeder:
    """patches torch.nn.functional.dropout: every live call pops the next exported mask"""

    def __init__(self, masks):
        self.masks, self.i = masks, 0

    def __call__(self, input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        rec, m = self.masks[self.i]
        self.i += 1
        assert abs(rec[1] - p) < 1e-9 and m.numel() == input.numel(), (self.i - 1, rec, tuple(input.shape), p)
        return input * m.view(input.shape).to(input.dtype) / (1.0 - p)


@pytest.mark.parametrize("speech_kind,text_kind", [("mini", "bart-mini"), ("mini_large", "mbart-mini"), ("mini", "t5-mini")])
def test_training_with_dropout_matches_mask_fed_oracle(speech_kind, text_kind, cuda_device, monkeypatch):
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED, kernels as K, ops
    sp_cfg = O.speech_config(speech_kind)
    sp_cfg.hidden_dropout, sp_cfg.attention_dropout, sp_cfg.activation_dropout, sp_cfg.feat_proj_dropout = 0.1, 0.1, 0.1, 0.05
    tx_cfg = O.text_config(text_kind)
    if tx_cfg.model_type == "t5":
        tx_cfg.dropout_rate = 0.1
    else:
        tx_cfg.dropout, tx_cfg.attention_dropout, tx_cfg.activation_dropout = 0.1, 0.1, 0.1
    sp_cfg._attn_implementation = "eager"      # the eager HF attention applies F.dropout to the probabilities
    tx_cfg._attn_implementation = "eager"
    speech, text = O.build_backbones(sp_cfg, tx_cfg, seed=0)
    ora = O.OracleEED(speech, text, down_scale=2)
    O.reinit_glue(ora, 1)
    ora.train()
    mine = SpeechMixEED(sp_cfg, tx_cfg, down_scale=2)
    mine.load_state_dict(ora.state_dict())
    mine = mine.to(cuda_device).train()
    assert mine.dropout_sites
    x, labels = O.synthetic_batch(2, 1.0, 8, tx_cfg.vocab_size, seed=0)
    ops.DROPOUT.manual_seed(77)
    ops.DROPOUT.trace = []
    out = mine(x.to(cuda_device), labels=labels.to(cuda_device))
    sites, ops.DROPOUT.trace = ops.DROPOUT.trace, None
    assert len(sites) > 10
    st = ops.DROPOUT.state(cuda_device)
    masks = [(rec, K.dropout_mask(rec[3], st, rec[0], rec[1], attention=(rec[2] == "attention")).cpu()) for rec in sites]
    feeder = _MaskFeeder(masks)
    monkeypatch.setattr(torch.nn.functional, "dropout", feeder)
    ref = ora(x, labels=labels, keep_full_logits=True)
    assert feeder.i == len(masks), (feeder.i, len(masks))       # same number of live dropout sites, same order
    monkeypatch.undo()
    assert abs(float(out["loss"]) - float(ref["loss"])) < 6e-3
    logits = mine.decoder_model.full_logits(out["decoder_last_hidden_state"])
    assert float((logits.float().cpu() - ref["full_logits"]).abs().max() / ref["full_logits"].abs().max()) < 2e-2
    ref["loss"].backward()
    out["loss"].backward()
    po, pm = dict(ora.named_parameters()), dict(mine.named_parameters())
    scale = max(float(p.grad.norm()) for p in po.values() if p.grad is not None)
    tol = 1.2e-1 if tx_cfg.model_type == "t5" else 6e-2
    for k, p in po.items():
        if p.grad is None:
            continue
        err = float((pm[k].grad.cpu() - p.grad).norm())
        assert err <= tol * float(p.grad.norm()) + 3e-4 * scale, (k, err, float(p.grad.norm()))
    # and the masks matter: a deterministic forward gives a different loss
    mine.eval()
    with torch.no_grad():
        l_eval = float(mine(x.to(cuda_device), labels=labels.to(cuda_device))["loss"])
    assert abs(l_eval - float(out["loss"])) > 1e-4


def test_cuda_graph_replays_draw_new_masks(cuda_device):
    """the step counter behind the masks lives on the device and is advanced inside the captured step"""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    from speechmix_b200.graph import GraphedTrainStep
    sp_cfg = O.speech_config("mini")
    sp_cfg.hidden_dropout = 0.2
    m = SpeechMixEED(sp_cfg, O.text_config("bart-mini"), down_scale=2).to(cuda_device).train()
    x, y = O.synthetic_batch(2, 1.0, 8, 1000)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    g = GraphedTrainStep(m, opt, x.to(cuda_device), y.to(cuda_device), warmup=2)
    losses = [float(g(x.to(cuda_device), y.to(cuda_device))) for _ in range(4)]
    assert len(set(losses)) == 4, losses        # lr = 0 and the same batch: only the masks change
