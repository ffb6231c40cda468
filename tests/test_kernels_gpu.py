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
