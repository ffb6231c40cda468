"""ctypes binding of libspeechmix_sm100.so (see include/speechmix_sm100.h).

There is deliberately NO fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMX_LIB: A/B runs of two builds of the same ABI (development aid; never a fallback -- a missing file still raises)
LIB_PATH = os.environ.get("SMX_LIB") or os.path.join(_HERE, "libspeechmix_sm100.so")
SMX_MAX_SEG = 4

GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_DGELU, ACT_DRELU, ACT_GELU_G, ACT_MULAUX = 0, 1, 2, 3, 4, 5, 6
OUT_BF16, OUT_F32 = 0, 1


class SmxView3(Structure):
    _fields_ = [("ptr", c_void_p), ("inner", c_int64), ("rows", c_int64), ("batches", c_int64),
                ("row_stride", c_int64), ("batch_stride", c_int64)]


class SmxGemm(Structure):
    _fields_ = [("mode", c_int32), ("out_dtype", c_int32), ("a", SmxView3), ("b", SmxView3),
                ("m", c_int64), ("n", c_int64), ("k", c_int64), ("batches", c_int64),
                ("nseg", c_int32), ("seg_len", c_int32),
                ("a_row_off", c_int32 * SMX_MAX_SEG), ("a_col_off", c_int32 * SMX_MAX_SEG),
                ("b_row_off", c_int32 * SMX_MAX_SEG), ("b_col_off", c_int32 * SMX_MAX_SEG),
                ("c", c_void_p), ("c_row_stride", c_int64), ("c_batch_stride", c_int64),
                ("act", c_int32), ("split_k", c_int32), ("accumulate", c_int32), ("alpha", c_float),
                ("bias", c_void_p), ("residual", c_void_p),
                ("res_row_stride", c_int64), ("res_batch_stride", c_int64),
                ("aux_out", c_void_p), ("aux_in", c_void_p)]


class SmxAttn(Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p), ("lse", c_void_p),
                ("q_row_stride", c_int64), ("k_row_stride", c_int64), ("v_row_stride", c_int64),
                ("o_row_stride", c_int64),
                ("q_batch_stride", c_int64), ("k_batch_stride", c_int64), ("v_batch_stride", c_int64),
                ("o_batch_stride", c_int64),
                ("batch", c_int32), ("heads", c_int32), ("tq", c_int32), ("tk", c_int32), ("causal", c_int32),
                ("scale", c_float), ("bias", c_void_p),
                ("d_o", c_void_p), ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
                ("delta", c_void_p), ("dbias", c_void_p),
                ("do_row_stride", c_int64), ("do_batch_stride", c_int64),
                ("dq_row_stride", c_int64), ("dk_row_stride", c_int64), ("dv_row_stride", c_int64),
                ("dq_batch_stride", c_int64), ("dk_batch_stride", c_int64), ("dv_batch_stride", c_int64),
                ("kv_len", c_void_p), ("dropout_state", c_void_p), ("dropout_call", ctypes.c_uint32),
                ("dropout_p", c_float)]


class SmxAdafactorTensor(Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("row", c_void_p), ("col", c_void_p), ("row_acc", c_void_p),
                ("col_acc", c_void_p), ("rmean", c_void_p), ("sumsq", c_void_p),
                ("batch", c_int64), ("rows", c_int64), ("cols", c_int64), ("numel", c_int64),
                ("factored", c_int32), ("pad_", c_int32)]


_P = c_void_p
_I64 = c_int64

# name -> (restype, argtypes); mirrors include/speechmix_sm100.h one to one
SIGNATURES = {
    "smx_last_error": (c_char_p, []),
    "smx_abi_version": (c_int, []),
    "smx_device_ok": (c_int, []),
    "smx_debug_attn_trace": (c_int, [_P]),
    "smx_gemm": (c_int, [POINTER(SmxGemm), _P]),
    "smx_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, c_float, c_int, c_int, _P]),
    "smx_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, c_int, c_int, _P]),
    "smx_colsum": (c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "smx_mask_rows": (c_int, [_P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "smx_dropout": (c_int, [_P, _P, _P, _P, _P, c_int, _I64, _P, ctypes.c_uint32, c_float, _P]),
    "smx_dropout_mask": (c_int, [_P, _I64, _I64, c_int, _P, ctypes.c_uint32, c_float, _P]),
    "smx_layernorm_dropout_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, c_float, _P, ctypes.c_uint32, c_float, _P]),
    "smx_layernorm_dropout_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, ctypes.c_uint32, c_float, _P]),
    "smx_adafactor_step": (c_int, [_P, c_int32, _P, c_int32, _P, c_int32, _P, c_int32, c_int32, _P, _I64, c_float, c_float,
                                   c_float, c_float, c_float, _P, c_float, _P]),
    "smx_spec_augment_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "smx_spec_augment_bwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "smx_cast_f32_to_bf16": (c_int, [_P, _P, _I64, _P]),
    "smx_weightnorm_fwd": (c_int, [_P, _P, _P, _P, _I64, _I64, _P]),
    "smx_weightnorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "smx_multi_cast": (c_int, [_P, c_int32, c_int32, _P]),
    "smx_add_bf16": (c_int, [_P, _P, _P, _I64, _P]),
    "smx_mul_bf16": (c_int, [_P, _P, _P, _I64, _P]),
    "smx_act_bf16": (c_int, [_P, _P, _I64, c_int, _P]),
    "smx_dact_bf16": (c_int, [_P, _P, _P, _I64, c_int, _P]),
    "smx_pack_conv_weight": (c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "smx_unpack_conv_wgrad": (c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "smx_conv0_stats": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, c_float, _P]),
    "smx_conv0_gn_gelu_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, _P]),
    "smx_conv0_gn_gelu_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, _P]),
    "smx_conv0_ln_gelu_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, c_float, _P]),
    "smx_conv0_ln_gelu_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, c_float, _P]),
    "smx_conv0_wgrad": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, c_int, c_int, c_int, _P]),
    "smx_posconv_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, c_int, c_int, c_int, c_int, _P]),
    "smx_posconv_dgrad": (c_int, [_P, _P, _P, _P, _I64, _I64, c_int, c_int, c_int, _P]),
    "smx_posconv_wgrad": (c_int, [_P, _P, _P, _I64, _I64, c_int, c_int, c_int, _P]),
    "smx_attn_fwd": (c_int, [POINTER(SmxAttn), _P]),
    "smx_attn_bwd": (c_int, [POINTER(SmxAttn), _P]),
    "smx_embed_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, c_float, _I64, _I64, _P]),
    "smx_embed_bwd": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, c_float, _I64, _P]),
    "smx_lmhead_ws_bytes": (c_size_t, [_I64, _I64]),
    "smx_lmhead_ce_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, c_float, _I64, _P]),
    "smx_lmhead_dlogits": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, c_float, _P]),
    "smx_weighted_sum_fwd": (c_int, [_P, _P, _P, c_int, _I64, _P]),
    "smx_weighted_sum_bwd_w": (c_int, [_P, _P, _P, c_int, _I64, _P]),
    "smx_kl_chunk_fwd": (c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "smx_kl_finalize": (c_int, [_P, _P, _P, _I64, c_float, _P, _P]),
    "smx_kl_chunk_bwd": (c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "smx_gram_dot_fwd": (c_int, [_P, _P, _P, _I64, _I64, _I64, _P]),
    "smx_gram_dot_bwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "smx_self_mse_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "smx_self_mse_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "smx_f32_gemm_nt": (c_int, [_P, _I64, _I64, _P, _P, _P, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I64, _I64, c_int, c_float, _P]),
    "smx_f32_layernorm": (c_int, [_P, _P, _P, _P, _I64, _I64, c_float, c_int, c_int, _P]),
    "smx_f32_groupnorm_gelu": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, c_float, _P]),
    "smx_f32_posconv": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, c_int, _P]),
    "smx_f32_attn": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, c_int, c_float, _P, _P]),
    "smx_f32_embed": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, c_float, _I64, _P]),
    "smx_f32_argmax_chunk": (c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "smx_f32_axpy": (c_int, [_P, _P, c_int32, _P, _I64, c_int, _P]),
    "smx_relpos_bias_fwd": (c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "smx_relpos_bias_bwd": (c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
}

_lib = None

# Optional per-call device timing (tools/profile_step.py): name -> [(start_event, end_event), ...]
PROFILE = None


def _profiled(name, fn):
    def call(*args):
        if PROFILE is None:
            return fn(*args)
        import torch
        key = name
        if name == "smx_gemm":
            g = args[0]._obj
            key = "gemm_%s m%d n%d k%d b%d act%d" % (("NT", "NN", "TN")[g.mode], g.m, g.n, g.k, g.batches, g.act)
        elif name in ("smx_attn_fwd", "smx_attn_bwd"):
            a = args[0]._obj
            key = "%s tq%d tk%d h%d c%d" % (name, a.tq, a.tk, a.heads, a.causal)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.setdefault(key, []).append((e0, e1))
        return rc
    return call


class _Lib:
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libspeechmix_sm100.so is missing (%s). Build it with `python -m speechmix_b200.build`; "
            "there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.smx_abi_version() != 3:
        raise RuntimeError("libspeechmix_sm100.so ABI version mismatch")
    shim = _Lib()
    for name in SIGNATURES:
        fn = getattr(lib, name)
        setattr(shim, name, _profiled(name, fn) if SIGNATURES[name][0] is c_int and name.startswith("smx_") and
                name not in ("smx_abi_version", "smx_device_ok") else fn)
    shim._cdll = lib
    _lib = shim
    return shim


# kernels launched per successful call (bench.py's gpu_launches); smx_gemm is counted by its caller
KERNELS_PER_CALL = {"multi_cast": 1, "weightnorm_fwd": 2, "weightnorm_bwd": 2, "kl_chunk_fwd": 1, "kl_finalize": 1, "kl_chunk_bwd": 1, "self_mse_fwd": 1,
                    "self_mse_bwd": 2, "relpos_fwd": 1, "relpos_bwd": 1, "smx_attn_fwd": 1, "smx_attn_bwd": 2, "dropout": 1, "dropout_mask": 1, "layernorm_fwd": 1, "layernorm_bwd": 1, "colsum": 1,
                    "cast": 1, "add": 1, "dact": 1, "conv0_stats": 3, "conv0_fwd": 1, "conv0_bwd": 2, "conv0_ln_fwd": 1, "conv0_ln_bwd": 1, "conv0_wgrad": 1,
                    "posconv_fwd": 1, "posconv_dgrad": 1, "posconv_wgrad": 1, "embed_fwd": 1, "embed_bwd": 1,
                    "lmhead_ce_fwd": 2, "lmhead_dlogits": 1, "wsum_fwd": 1, "wsum_bwd": 1,
                    "layernorm_dropout_fwd": 1, "layernorm_dropout_bwd": 1, "gram_dot_fwd": 1, "gram_dot_bwd": 1,
                    "mask_rows": 1, "mul": 1, "spec_augment_fwd": 1, "spec_augment_bwd": 1}   # (+1 with a dembed output)
LAUNCHES = [0]


def check(rc, what=""):
    LAUNCHES[0] += KERNELS_PER_CALL.get(what, 0)
    if rc != 0:
        msg = load().smx_last_error()
        raise RuntimeError("libspeechmix_sm100 %s failed: %s" % (what, msg.decode() if msg else "?"))
