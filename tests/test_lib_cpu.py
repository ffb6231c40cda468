"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads without a
GPU, and exports exactly the symbols include/speechmix_sm100.h declares.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "speechmix_sm100.h")


@pytest.fixture(scope="module")
def lib_path():
    from speechmix_b200 import build
    return build.build()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "smx_gemm" in syms and "smx_attn_fwd" in syms and "smx_lmhead_ce_fwd" in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export: " + s


def test_binding_table_matches_header(lib_path):
    from speechmix_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.smx_abi_version() == 3


def test_struct_layout_matches_c():
    """sizeof(SmxGemm)/sizeof(SmxAttn) as seen by ctypes == as compiled by gcc from the header."""
    from speechmix_b200 import _lib
    prog = '#include <stdio.h>\n#include "speechmix_sm100.h"\nint main(){printf("%zu %zu %zu", sizeof(SmxGemm), sizeof(SmxAttn), sizeof(SmxView3));return 0;}'
    exe = "/tmp/smx_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(_lib.SmxGemm), ctypes.sizeof(_lib.SmxAttn), ctypes.sizeof(_lib.SmxView3)]


def test_kernels_are_blackwell_native(lib_path):
    """SASS evidence: tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM."""
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_no_fallback_without_gpu(lib_path):
    import torch
    from speechmix_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.smx_device_ok() != 1  # no device -> error, never a silent CPU path
    assert lib.smx_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "speechmix_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src, f


def test_state_dict_layout_matches_reference_oracle():
    """drop-in boundary: same parameter names / shapes / requires_grad bookkeeping as HFSpeechMixEED."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    for tx in ("bart-mini", "mbart-mini", "t5-mini", "t5v11-mini"):
        spc, txc = O.speech_config("mini"), O.text_config(tx)
        s, t = O.build_backbones(spc, txc)
        ora = O.OracleEED(s, t, down_scale=4, weighted_sum=True, fixed_parameters=True)
        mine = SpeechMixEED(spc, txc, down_scale=4, weighted_sum=True, fixed_parameters=True)
        a, b = ora.state_dict(), mine.state_dict()
        assert list(a) == list(b) or set(a) == set(b)
        assert all(a[k].shape == b[k].shape for k in a)
        assert ora.list_no_grad == mine.list_no_grad and ora.list_grad == mine.list_grad
        assert mine.speech_encoder_layer == ora.speech_encoder_layer and mine.nlp_encoder_layer == ora.nlp_encoder_layer
        mine.load_state_dict(a)


def test_speechmix_gan_layout_and_update_phases_match_reference_oracle():
    """SpeechMixGAN (ref:speechmix/hf_model.py:586-694): discriminator Linear(D*D, 1) under the reference's state-dict
    key, nothing frozen, and the update-phase state machine (:609-626: counters + which family's ``.grad`` is cleared
    before the step) walks through the same states as the oracle's literal restatement, including the switch at
    ``update_count % des_update == 0`` and the ``keep_update`` countdown."""
    from oracle import hf_oracle as O
    import torch
    from speechmix_b200 import HFSpeechMixGAN, SpeechMixGAN
    assert HFSpeechMixGAN is SpeechMixGAN
    spc, txc = O.speech_config("mini"), O.text_config("bart-mini")
    s, t = O.build_backbones(spc, txc)
    ora = O.OracleGAN(s, t, down_scale=2)
    mine = SpeechMixGAN(spc, txc, down_scale=2)
    a, b = ora.state_dict(), mine.state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
    d = txc.d_model
    assert b["discriminator.weight"].shape == (1, d * d) and b["discriminator.bias"].shape == (1,)
    assert ora.list_grad == mine.list_grad and mine.list_no_grad == [] == ora.list_no_grad
    mine.load_state_dict(a)

    def oracle_phase(m):      # the block of OracleGAN.cal_loss under ``if self.training`` (ref :609-626), no model pass
        if m.update_count % m.des_update == 0:
            if m.keep_update > 0:
                m.keep_update -= 1
                for name, p in m.named_parameters():
                    if "discriminator" in name:
                        p.grad = None
            else:
                m.keep_update = 1000
                m.update_count += 1
        else:
            m.update_count += 1
            for name, p in m.named_parameters():
                if "discriminator" not in name:
                    p.grad = None
    import pytest as _pytest
    from speechmix_b200 import graph
    with _pytest.raises(RuntimeError, match="update phase"):    # host-side phase counters cannot live in a CUDA graph
        graph._check_no_host_randomness(mine)
    for m in (ora, mine):
        m.des_update, m.keep_update = 3, 2          # short cycle: 1 -> 2 -> 3 (hold 2 steps) -> reset -> 4 ...
    for step in range(9):
        for m in (ora, mine):
            for p in m.parameters():
                p.grad = torch.zeros_like(p)
        oracle_phase(ora)
        mine._phase()
        assert (ora.update_count, ora.keep_update) == (mine.update_count, mine.keep_update), step
        ga = {k for k, p in ora.named_parameters() if p.grad is None}
        gb = {k for k, p in mine.named_parameters() if p.grad is None}
        assert ga == gb, step


def test_speechmix_ed_layout_matches_reference_oracle():
    """SpeechMixED (ref:speechmix/hf_model.py:82-182): same state-dict keys in the same order, same shapes, same frozen
    feature encoder as the hf SpeechEncoderDecoderModel the reference builds; fixed_parameters follows the reference's
    substring rule."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixED
    for kw in ({}, {"fixed_parameters": True}):
        spc, txc = O.speech_config("mini"), O.text_config("bart-mini")
        s, t = O.build_backbones(spc, txc)
        ora = O.OracleED(s, t, **kw)
        mine = SpeechMixED(spc, txc, **kw)
        a, b = ora.state_dict(), mine.state_dict()
        assert list(a) == list(b) and all(a[k].shape == b[k].shape for k in a)
        assert [k for k, p in ora.named_parameters() if not p.requires_grad] == mine.list_no_grad
        assert [k for k, p in ora.named_parameters() if p.requires_grad] == mine.list_grad
        mine.load_state_dict(a)


def test_share_layer_ratio_and_helpers():
    """ref:test/test_hf_model.py:18-33 restated on the product classes."""
    import torch
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED, shift_tokens_right
    for ratio, kept in [(1, 0), (0.5, 1), (0, 2)]:
        m = SpeechMixEED(O.speech_config("mini"), O.text_config("bart-mini"), share_layer_ratio=ratio, down_scale=8)
        assert m.speech_encoder_layer == kept and m.nlp_encoder_layer == 2 and len(m.list_no_grad) == 0
    lab = torch.tensor([[5, 6, -100, -100], [7, 8, 9, 10]])
    assert torch.equal(shift_tokens_right(lab, 1, 2), O.shift_tokens_right(lab, 1, 2))


def test_freezing_variants_mark_the_same_parameters_as_the_reference():
    """SpeechMixFixed (ref:speechmix/hf_model.py:450-462) and fixed_parameters (ref :226-244): the product freezes the
    same parameters as the oracle, which the mini_fixed / mini_fixed_params goldens pin to the unmodified reference."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED, SpeechMixFixed
    from tests._cases import load_fixture
    for name, mine_cls in (("mini_fixed", SpeechMixFixed), ("mini_fixed_params", SpeechMixEED)):
        fx = load_fixture(name)
        spc, txc = O.speech_config(fx["speech"], model_type=fx["speech_type"]), O.text_config(fx["text"])
        mine = mine_cls(spc, txc, **fx["kwargs"])
        assert len(mine.list_no_grad) == fx["list_no_grad"] and len(mine.list_grad) == fx["list_grad"]
        s, t = O.build_backbones(spc, txc)
        ora = getattr(O, "Oracle" + fx.get("cls", "EED"))(s, t, **fx["kwargs"])
        assert mine.list_grad == ora.list_grad and mine.list_no_grad == ora.list_no_grad
        assert [k for k, p in mine.named_parameters() if p.requires_grad] == [k for k, p in ora.named_parameters() if p.requires_grad]


def test_t5_layer_sharing_bookkeeping_matches_reference_golden():
    """`.block` branch of the layer bookkeeping (ref:speechmix/hf_model.py:232-251) on the product class, against the
    numbers the unmodified reference produced for the mini_t5_share golden."""
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED
    from tests._cases import build_oracle, load_fixture
    fx = load_fixture("mini_t5_share")
    mine = SpeechMixEED(O.speech_config(fx["speech"], model_type=fx["speech_type"]), O.text_config(fx["text"]), **fx["kwargs"])
    assert mine.speech_encoder_layer == fx["speech_encoder_layer"] and mine.nlp_encoder_layer == fx["nlp_encoder_layer"]
    assert len(mine.state_dict()) == fx["n_state_keys"] and len(mine.list_no_grad) == fx["list_no_grad"]
    assert sum(p.numel() for p in mine.parameters()) == fx["n_params"]
    ora, _, _ = build_oracle(fx)
    mine.load_state_dict(ora.state_dict())
