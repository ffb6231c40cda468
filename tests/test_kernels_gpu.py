"""GPU parity tests of every kernel family, called through the C ABI (ctypes) and compared with
a plain torch fp32 evaluation of the same op on the same bf16-rounded inputs (tolerance:
max|err| / max|ref| <= 2e-2 for bf16 outputs, <= 1e-3 for fp32 reductions; see tools/probe_*.py)."""
import pytest

pytestmark = pytest.mark.gpu


def _cases(mod):
    return [n for n in mod.CASES if n != "perf"]


def _run(mod, name, cuda_device):
    assert mod.CASES[name]() is True


from tools import probe_attn, probe_gemm, probe_misc  # noqa: E402


@pytest.mark.parametrize("name", _cases(probe_gemm))
def test_gemm_family(name, cuda_device):
    """tcgen05 GEMM: NT / NN / TN modes, fused epilogues, implicit-GEMM strided convs (fwd, dgrad, wgrad)."""
    _run(probe_gemm, name, cuda_device)


@pytest.mark.parametrize("name", _cases(probe_attn))
def test_attention(name, cuda_device):
    """flash-style attention fwd/bwd: ragged lengths, causal, cross-attention shapes, fused-QKV strides,
    full-size (T=749 / 1499) cases."""
    _run(probe_attn, name, cuda_device)


@pytest.mark.parametrize("name", _cases(probe_misc))
def test_rowwise_conv0_posconv_lmhead(name, cuda_device):
    _run(probe_misc, name, cuda_device)


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adafactor_matches_transformers(weight_decay, cuda_device):
    """ref:train.py:298 optim="adafactor" -> transformers.optimization.Adafactor(scale_parameter=False,
    relative_step=False): parameters and second-moment state after 4 steps, every shape family of the model
    (matrices, conv weights with leading dims, tall / wide / tiny, vectors), one launch group per step."""
    import torch
    from transformers.optimization import Adafactor
    from speechmix_b200.optim import FusedAdafactor
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(768,), (1,), (1000, 768), (512, 64, 3), (128, 1, 10), (96, 48, 128), (1, 5000), (3072, 768), (300,),
              (2, 3, 65, 257)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g) * 0.1) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = Adafactor(ref_p, lr=5e-4, scale_parameter=False, relative_step=False, weight_decay=weight_decay)
    mine = FusedAdafactor(my_p, lr=5e-4, weight_decay=weight_decay)
    for step in range(4):
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (10.0 if step == 2 else 0.01)   # clipped / unclipped
            a.grad, b.grad = gr, gr.clone()
        ref.step()
        mine.step()
    for a, b, s in zip(ref_p, my_p, shapes):
        d = float((a - b).abs().max())
        upd = float((a - torch.zeros_like(a)).abs().max()) + 1e-12
        assert d <= 2e-6 + 1e-5 * upd, (s, d)
        sa, sb = ref.state[a], mine.state[b]
        assert sa["step"] == sb["step"] == 4
        for k in ("exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_sq"):
            if k in sa:
                assert sb[k].shape == sa[k].shape
                assert float(((sa[k] - sb[k]).abs() / (sa[k].abs() + 1e-20)).max()) < 1e-4, (s, k)
    # a parameter that starts receiving gradients later has its own step count -> its own launch group
    late = torch.nn.Parameter(torch.randn(64, 64, device="cuda", generator=g))
    late_ref = torch.nn.Parameter(late.detach().clone())
    mine.add_param_group({"params": [late]})
    ref.add_param_group({"params": [late_ref]})
    for a, b in zip(ref_p + [late_ref], my_p + [late]):
        gr = torch.randn(a.shape, device="cuda", generator=g) * 0.01
        a.grad, b.grad = gr, gr.clone()
    ref.step()
    mine.step()
    assert float((late - late_ref).abs().max()) < 2e-6 and mine.state[late]["step"] == 1
    assert float((ref_p[2] - my_p[2]).abs().max()) < 5e-6


@pytest.mark.parametrize("env", [{"SMX_ATTN_POLY": "4"}, {"SMX_ATTN_FWD_V1": "1"}, {"SMX_PDL": "0"}])
def test_attention_forward_env_selected_variants(env, cuda_device):
    """The A/B switches of the attention forward are read once per process, so each runs in its own interpreter:
    SMX_ATTN_POLY=4 (every fourth pair of exponentials on the FMA pipe: Cody-Waite + cubic instead of MUFU.EX2),
    SMX_ATTN_FWD_V1=1 (the general kernel also for the plain non-causal case), SMX_PDL=0 (no programmatic dependent
    launch).  Same parity cases, same bounds."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for case in ("fwd_small", "fwd2_rescale"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "probe_attn.py"), "--case", case],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0 and '"ok": true' in r.stdout, (env, case, r.stdout[-2000:], r.stderr[-2000:])
