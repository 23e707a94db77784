"""Device-side StyleGAN2 layers of TextBoxGAN (NHWC, bf16 activations, fp32 parameters).

Host code only orchestrates: the contractions run on the tcgen05 kernels through
:mod:`textboxgan_b200.conv`; parameter-sized algebra (modulation vectors, demodulation
coefficients, weight re-layouts) is fp32 torch on tensors of at most a few MB.

Layer-by-layer correspondence with the reference (models/custom_stylegan2/layers/):
``dense`` dense.py:13-29 · ``bias_act`` bias_act.py:25-34 · ``modulated_conv2d``
modulated_conv2d.py:66-122 (non-fused algebra :95-96,119-121: scale activations by the style,
convolve with the shared weight, scale by the demodulation coefficient) · ``to_rgb``
to_rgb.py:28-33 · ``upsample_rgb`` upfirdn_2d_v2.py:58-62 · ``minibatch_std``
mini_batch_std.py:10-35.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import conv as C
from . import upfirdn as U

Params = Dict[str, torch.Tensor]
SQRT2 = math.sqrt(2.0)
ACT_DTYPE = torch.bfloat16

# Plain steps run every layer group as a fused first-order Function (textboxgan_b200.fused); code
# that needs gradients of gradients (path-length / R1 regularisers) wraps its forward pass in
# ``double_backward()`` to get the composable, arbitrarily differentiable primitives instead.
FUSED = True
# upsample_conv_2d as transposed conv + separate FIR pass (algorithmic FLOPs) instead of the FIR folded into
# a 4-phase 3x3 GEMM (4x the tensor work, no intermediate)
UNFOLD_UP = True
# conv_downsample_2d: FIR pre-pass + k x k strided conv (forward, weight gradient) instead of the folded
# (k+3) x (k+3) convolution; the input gradient stays folded
UNFOLD_DOWN = True
FUSE_RGB_BACKWARD = True   # conv_1 + ToRGB + skip sum of a synthesis block as one autograd node (fused.ModConvActRGB)
SPEC_SECOND_ORDER = True   # regulariser (double-backward) passes: convolutions as second_order.lin_conv on the master weight
SKIP_SPLIT = True          # residual skip branch: FIR at the strided pixels only + plain 1x1 GEMM (fused.SkipSplit)
_DOUBLE_BACKWARD = False


class double_backward:
    def __enter__(self):
        global _DOUBLE_BACKWARD
        self.prev = _DOUBLE_BACKWARD
        _DOUBLE_BACKWARD = True
        return self

    def __exit__(self, *exc):
        global _DOUBLE_BACKWARD
        _DOUBLE_BACKWARD = self.prev
        return False


def use_fused() -> bool:
    return FUSED and torch.is_grad_enabled() and not _DOUBLE_BACKWARD


def runtime_coef(weight_shape, gain: float = 1.0, lrmul: float = 1.0) -> float:
    """commons.py:4-12"""
    fan_in = 1
    for d in weight_shape[:-1]:
        fan_in *= int(d)
    return gain / math.sqrt(fan_in) * lrmul


def dense(x: torch.Tensor, w: torch.Tensor, gain: float = 1.0, lrmul: float = 1.0) -> torch.Tensor:
    return x.reshape(x.shape[0], -1).float() @ (runtime_coef(w.shape, gain, lrmul) * w)


def lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, 0.2) * SQRT2


def style_scale(style: torch.Tensor, P: Params, prefix: str) -> torch.Tensor:
    """s = mod_bias(mod_dense(y)) + 1   (modulated_conv2d.py:75-76)."""
    return dense(style, P[prefix + "/mod_dense/w"]) + P[prefix + "/mod_bias/b"] + 1.0


def modulated_conv2d(x: torch.Tensor, style: torch.Tensor, P: Params, prefix: str, *, up: bool,
                     demodulate: bool = True, noise: Optional[torch.Tensor] = None,
                     noise_strength: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                     act: bool = False, fused_epilogue: bool = False,
                     s_pre: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [B,H,W,I] bf16 -> [B,H',W',O] bf16, including (optionally) noise, bias and activation.

    ``fused_epilogue`` applies demod/noise/bias/lrelu inside the GEMM epilogue; it is only legal
    when no gradient is needed through those terms (inference), otherwise the terms are applied
    with differentiable element-wise ops."""
    w_raw = P[prefix + "/w"]
    kh, kw, I, O = w_raw.shape
    s = s_pre if s_pre is not None else style_scale(style, P, prefix)       # [B,I]
    B, H, W_, _ = x.shape
    if use_fused() and act and noise is not None and bias is not None and demodulate and not fused_epilogue:
        from .fused import ModConvAct, ModUpConvAct

        if up and UNFOLD_UP:
            spec = C.weight_spec("upT", H, W_, I, O, kh, True, "modconv")
            return ModUpConvAct.apply(x, s, w_raw, noise, noise_strength, bias, spec, SQRT2)
        spec = C.weight_spec("up" if up else "plain", H, W_, I, O, kh, True, "modconv")
        return ModConvAct.apply(x, s, w_raw, noise, noise_strength, bias, spec, SQRT2)
    w = runtime_coef(w_raw.shape) * w_raw                                   # :71
    if up:
        geom = C.up_geom(H, W_, I, O, tag="modconv")
        wmat = C.up_wmat(w)
    else:
        geom = C.plain_geom(H, W_, I, O, kh, tag="modconv")
        wmat = C.plain_wmat(w)
    d = None
    if demodulate:
        q = (w * w).sum(dim=(0, 1))                                         # [I,O]
        d = torch.rsqrt((s * s) @ q + 1e-8)                                 # :80-82  [B,O]
    if fused_epilogue:
        from . import kernels as K

        xs = K.modulate(x.contiguous(), s.contiguous())                     # :96
        epi = dict(col_scale=d.contiguous() if d is not None else None,
                   noise=noise.contiguous() if noise is not None else None,
                   noise_strength=noise_strength.reshape(1).contiguous() if noise is not None else None,
                   bias=bias.contiguous() if bias is not None else None,
                   act=1 if act else 0, act_gain=SQRT2 if act else 1.0)
        with torch.no_grad():
            return C.conv(xs, wmat, geom, epi)
    # twice-differentiable path (path-length regulariser): every node is a member of the closed primitive set of
    # second_order.py, so the first AND the second backward pass stay on the kernels
    from . import second_order as SO

    xs = SO.modulate(x.to(ACT_DTYPE), s)                                    # :96
    if SPEC_SECOND_ORDER and kh == 3:
        y = SO.lin_conv(xs, w_raw, "upT" if up else "plain", kh, True, "modconv")
    else:
        y = C.conv(xs, wmat, geom)
    if d is not None:
        y = SO.modulate(y, d)                                               # :121
    if noise is None and bias is None and not act:
        return y
    return SO.bias_act(y, noise, noise_strength if noise is not None else None, bias, 1 if act else 0,
                       SQRT2 if act else 1.0)                               # noise.py:21, bias_act.py:25-34


def modulated_conv2d_rgb(x: torch.Tensor, P: Params, prefix: str, rgb_prefix: str, *, noise: torch.Tensor,
                         noise_strength: torch.Tensor, bias: torch.Tensor, s_conv: torch.Tensor, s_rgb: torch.Tensor,
                         y_prev: Optional[torch.Tensor], mask_words: Optional[torch.Tensor], nchw: bool):
    """Second convolution of a synthesis block + its ToRGB + the skip sum (synthesis_block.py:143-152) as one autograd
    node (fused.ModConvActRGB).  Returns (x_out bf16 NHWC, y)."""
    from .fused import ModConvActRGB

    w_raw = P[prefix + "/w"]
    kh, _, I, O = w_raw.shape
    _, H, W_, _ = x.shape
    spec = C.weight_spec("plain", H, W_, I, O, kh, True, "modconv")
    w_rgb_raw = P[rgb_prefix + "/conv/w"]                                   # [1,1,C,3]
    w_rgb = runtime_coef(w_rgb_raw.shape) * w_rgb_raw[0, 0]
    ws = s_rgb[:, :, None] * w_rgb[None]                                    # [B,C,3]  (to_rgb.py:28-33: no demodulation)
    return ModConvActRGB.apply(x, s_conv, w_raw, noise, noise_strength, bias, spec, SQRT2, ws, P[rgb_prefix + "/bias/b"],
                               y_prev, mask_words, nchw)


def all_style_scales(style: torch.Tensor, P: Params, prefixes, idxs):
    """s_l for every modulated convolution ``prefixes[l]`` fed by style row ``idxs[l]``, in one launch
    (fused.StyleScales); style [B, n_style, S]."""
    from .fused import StyleScales

    wb = []
    for pf in prefixes:
        wb += [P[pf + "/mod_dense/w"], P[pf + "/mod_bias/b"]]
    coef = runtime_coef(P[prefixes[0] + "/mod_dense/w"].shape)
    return StyleScales.apply(style, tuple(int(i) for i in idxs), coef, *wb)


def to_rgb(x: torch.Tensor, style: torch.Tensor, P: Params, prefix: str,
           s_pre: Optional[torch.Tensor] = None, skip_args=None) -> torch.Tensor:
    """1x1 modulated conv without demodulation + bias, fp32 [B,H,W,3] (to_rgb.py:28-33).  ``skip_args`` =
    (y_prev | None, mask_words | None, nchw): fused with the upsampled skip sum of synthesis_block.py:152 and, on the last
    block, mask_text_box and the NCHW layout of the image (fused.ToRGBSkip)."""
    w_raw = P[prefix + "/conv/w"]                                           # [1,1,C,3]
    w = runtime_coef(w_raw.shape) * w_raw[0, 0]
    s = s_pre if s_pre is not None else style_scale(style, P, prefix + "/conv")
    ws = s[:, :, None] * w[None]                                            # [B,C,3]
    if use_fused() and skip_args is not None:
        from .fused import ToRGBSkip

        y_prev, mask_words, nchw = skip_args
        return ToRGBSkip.apply(x, ws, P[prefix + "/bias/b"], y_prev, mask_words, nchw)
    if use_fused():
        from .fused import ToRGB

        return ToRGB.apply(x, ws, P[prefix + "/bias/b"])
    from . import second_order as SO

    return SO.to_rgb(x.to(ACT_DTYPE), ws) + P[prefix + "/bias/b"]


def upsample_rgb(y: torch.Tensor) -> torch.Tensor:
    """upsample_2d(y) with k = outer([1,3,3,1])/64*4, pad (2,1) (synthesis_block.py:97-99,152) on
    the native upfirdn2d op; y is fp32 [B,H,W,3]."""
    return U.upsample_2d_nhwc(y.contiguous())


def minibatch_std(x: torch.Tensor, group_size: int = 4, n_calls: int = 1) -> torch.Tensor:
    """mini_batch_std.py:10-35 on NHWC: returns the [B,1] per-sample statistic (fp32).  ``n_calls`` > 1:
    ``x`` is the concatenation of that many independent discriminator calls; the statistic is taken
    inside each call's own batch, exactly as if the calls were separate."""
    Bt = x.shape[0]
    B = Bt // n_calls
    g = min(group_size, B)
    y = x.float().reshape(n_calls, g, B // g, -1)
    y = y - y.mean(dim=1, keepdim=True)
    y = torch.sqrt((y * y).mean(dim=1) + 1e-8)                              # [n_calls, B/g, F]
    y = y.mean(dim=2, keepdim=True)                                         # [n_calls, B/g, 1]
    return y[:, None].expand(n_calls, g, B // g, 1).reshape(Bt, 1)
