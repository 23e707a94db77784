"""ctypes binding of ``libtbg.so`` — the C-ABI boundary declared in ``include/tbg.h``.

The product path has no fallback: if the shared library is missing the import of the first op
raises.  (The reference loads its plugin with ``tf.load_op_library`` and likewise fails hard,
models/custom_stylegan2/layers/upfirdn/custom_ops.py:182-207.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtbg.so"

c_void_p = C.c_void_p
c_int = C.c_int
c_float = C.c_float


class TbgError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status (mirrors a TF ``Status`` error)."""


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("w", c_void_p), ("out", c_void_p),
        ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int),
        ("Ho", c_int), ("Wo", c_int),
        ("n_total", c_int), ("cout", c_int),
        ("taps_h", c_int), ("taps_w", c_int),
        ("pad_h", c_int), ("pad_w", c_int),
        ("stride_h", c_int), ("stride_w", c_int),
        ("up_h", c_int), ("up_w", c_int),
        ("col_scale", c_void_p), ("bias", c_void_p), ("noise", c_void_p), ("noise_strength", c_void_p),
        ("residual", c_void_p), ("res_scale", c_float), ("res_first", c_int),
        ("act", c_int), ("act_gain", c_float), ("out_fp32", c_int),
        ("relu_mask", c_void_p),
        ("tap_mask", C.c_uint64 * 4),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("gy", c_void_p), ("gw", c_void_p),
        ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int),
        ("Ho", c_int), ("Wo", c_int),
        ("n_total", c_int), ("cout", c_int),
        ("taps_h", c_int), ("taps_w", c_int), ("pad_h", c_int), ("pad_w", c_int),
        ("stride_h", c_int), ("stride_w", c_int),
        ("up_h", c_int), ("up_w", c_int),
    ]


class StyleLayer(C.Structure):
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("s", c_void_p), ("gs", c_void_p), ("gw", c_void_p), ("gb", c_void_p),
                ("I", c_int), ("idx", c_int)]


class DecWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("wq", "wqT", "v", "emb", "wg", "wgT", "b", "wd", "wdT", "bd")]


_lib = None

# every symbol include/tbg.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("tbg_last_error", C.c_char_p, []),
    ("tbg_version", c_int, []),
    ("tbg_launch_count", C.c_longlong, []),
    ("tbg_reset_launch_count", None, []),
    ("tbg_crc32c", C.c_uint, [c_void_p, C.c_ulonglong, C.c_uint]),
    ("tbg_conv2d_igemm", c_int, [C.POINTER(ConvArgs), c_void_p]),
    ("tbg_conv2d_wgrad", c_int, [C.POINTER(WgradArgs), c_void_p]),
    ("tbg_upfirdn2d", c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 15 + [c_void_p]),
    ("tbg_adam_step", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_longlong, c_float, c_void_p, c_float,
                              c_float, c_float, c_void_p]),
    ("tbg_ema_step", c_int, [c_void_p, c_void_p, C.c_longlong, c_float, c_void_p]),
    ("tbg_lstm_seq_fwd", c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    ("tbg_lstm_seq_bwd", c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    ("tbg_modulate", c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    ("tbg_modulate_bwd", c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    ("tbg_bias_act_bwd", c_int, [c_void_p] * 9 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    ("tbg_bias_act_rgb_bwd", c_int, [c_void_p] * 11 + [c_int] * 4 + [c_float, c_void_p]),
    ("tbg_bias_act_fwd", c_int, [c_void_p] * 5 + [c_int] * 4 + [c_float, c_void_p]),
    ("tbg_rowdot", c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    ("tbg_torgb_fwd", c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    ("tbg_torgb_bwd", c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    ("tbg_wprep", c_int, [c_void_p, c_void_p, c_float] + [c_int] * 6 + [c_void_p] * 4),
    ("tbg_wprep_job_bytes", c_int, []),
    ("tbg_wprep_make_job", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float] + [c_int] * 6 + [c_void_p] * 3),
    ("tbg_wprep_group", c_int, [c_void_p, c_int, c_int, c_void_p]),
    ("tbg_wfold", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float] + [c_int] * 6 + [c_void_p] * 3 + [c_int, c_int, c_void_p]),
    ("tbg_set_tuning", c_int, [C.c_char_p, c_int]),
    ("tbg_get_tuning", c_int, [C.c_char_p]),
    ("tbg_crop_resize_fwd", c_int, [c_void_p] * 3 + [c_int] * 9 + [c_void_p]),
    ("tbg_crop_resize_bwd", c_int, [c_void_p] * 3 + [c_int] * 9 + [c_void_p]),
    ("tbg_batch_resize_normalize", c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p]),
    ("tbg_fromrgb_fwd", c_int, [c_void_p] * 4 + [c_int] * 3 + [c_float, c_float, c_void_p]),
    ("tbg_fromrgb_bwd", c_int, [c_void_p] * 7 + [c_int] * 3 + [c_float, c_float, c_void_p]),
    ("tbg_fir4", c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_float] + [c_void_p] * 4 + [c_int, c_float, c_void_p]),
    ("tbg_fir4_down", c_int, [c_void_p, c_void_p] + [c_int] * 9 + [c_float, c_void_p]),
    ("tbg_fir4_down_adjoint", c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_float, c_void_p]),
    ("tbg_wfold_adj", c_int, [c_void_p, c_void_p, c_float] + [c_int] * 5 + [c_void_p] * 3 + [c_int, c_int, c_int, c_void_p]),
    ("tbg_style_dense_fwd", c_int, [C.POINTER(StyleLayer), c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    ("tbg_style_dense_bwd", c_int, [C.POINTER(StyleLayer), c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                    c_void_p]),
    ("tbg_demod_coef", c_int, [c_void_p] * 3 + [c_int] * 3 + [c_float, c_void_p]),
    ("tbg_demod_bwd", c_int, [c_void_p] * 12 + [c_int] * 3 + [c_void_p]),
    ("tbg_attn_decoder_fwd", c_int, [c_void_p, c_void_p, C.POINTER(DecWeights)] + [c_void_p] * 7 + [c_int] * 3 + [c_void_p]),
    ("tbg_attn_decoder_bwd", c_int, [c_void_p, c_void_p, C.POINTER(DecWeights)] + [c_void_p] * 8 + [c_int] * 3 + [c_void_p]),
    ("tbg_dense_fwd", c_int, [c_void_p] * 4 + [c_int] * 3 + [c_float, c_float, c_int, c_float, c_void_p]),
    ("tbg_dense_bwd", c_int, [c_void_p] * 8 + [c_int] * 3 + [c_float, c_float, c_int, c_float, c_int, c_void_p]),
    ("tbg_pixel_norm_fwd", c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    ("tbg_pixel_norm_bwd", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    ("tbg_word_encoder_fwd", c_int, [c_void_p] * 4 + [c_float] + [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    ("tbg_word_encoder_bwd", c_int, [c_void_p] * 2 + [c_float] + [c_void_p] * 8 + [c_int] * 7 + [c_void_p]),
    ("tbg_minibatch_std_fwd", c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    ("tbg_minibatch_std_bwd", c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    ("tbg_r1_sqnorm", c_int, [c_void_p, c_void_p, c_int, C.c_longlong, c_void_p]),
    ("tbg_r1_sqnorm_bwd", c_int, [c_void_p, c_void_p, c_void_p, c_int, C.c_longlong, c_void_p]),
    ("tbg_torgb_skip_fwd", c_int, [c_void_p] * 6 + [c_int] * 6 + [c_void_p]),
    ("tbg_image_grad_nhwc", c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
]


def exported_symbols() -> list[str]:
    return [s[0] for s in _SIGNATURES]


def load() -> C.CDLL:
    """Load ``libtbg.so`` (built in-tree by ``textboxgan_b200.build``); raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TbgError(
            f"{LIB_PATH} not found: run `python -m textboxgan_b200.build` (or __graft_entry__.build()). "
            "There is no CPU or PyTorch fallback for the hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def set_tuning(key: str, value: int) -> None:
    """``tbg_set_tuning`` (include/tbg.h): explicit kernel-selection switches for tests and perf scripts."""
    check(load().tbg_set_tuning(key.encode(), int(value)), f"tbg_set_tuning({key})")


def get_tuning(key: str) -> int:
    return int(load().tbg_get_tuning(key.encode()))


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().tbg_last_error().decode("utf-8", "replace")
        raise TbgError(f"{what} failed with status {status}: {msg}")
