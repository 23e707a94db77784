"""upfirdn_2d and its wrappers (mirror of models/custom_stylegan2/layers/upfirdn/upfirdn_2d_v2.py)
over the ``tbg_upfirdn2d`` kernel.

Like ``upfirdn_2d_cuda`` (upfirdn_2d_v2.py:186-246) the gradient of the op is another call of the
same op with up/down swapped, the flipped kernel and pads ``gpad0 = k - pad0 - 1``,
``gpad1 = in*up - out*down + pad0 - up + 1`` — so gradients of any order are available.
"""
from __future__ import annotations

from functools import lru_cache

import torch

from . import kernels as K


# Resampling constants of the four sites the reference evaluates its padding rule at (SURVEY.md Appendix A.4;
# FIR k = outer([1,3,3,1]) / 64 times the gain).  (gain, pad0, pad1):
RESAMPLE_SITES = {
    "modconv_up": (4.0, 1, 1),     # upsample_conv_2d: convT(3x3, s2, VALID) -> pad -> 4x4 FIR   (modulated_conv2d.py:47-49)
    "rgb_up": (4.0, 2, 1),         # upsample_2d of the RGB skip: zero-insert x2 -> pad -> FIR     (synthesis_block.py:97-99)
    "conv_down_3x3": (1.0, 2, 3),  # conv_downsample_2d in front of a 3x3 stride-(2|1,2) conv      (conv.py:37-39)
    "conv_down_1x1": (1.0, 1, 2),  # ... in front of the 1x1 skip conv
}
FIR_TAPS = (1.0, 3.0, 3.0, 1.0)


class _UpFirDn2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, upx, upy, downx, downy, padx0, padx1, pady0, pady1):
        ctx.cfg = (upx, upy, downx, downy, padx0, padx1, pady0, pady1)
        ctx.in_hw = (x.shape[1], x.shape[2])
        ctx.save_for_backward(k)
        return K.upfirdn2d(x.contiguous(), k, upx=upx, upy=upy, downx=downx, downy=downy, padx0=padx0, padx1=padx1,
                           pady0=pady0, pady1=pady1)

    @staticmethod
    def backward(ctx, dy):
        (k,) = ctx.saved_tensors
        upx, upy, downx, downy, padx0, padx1, pady0, pady1 = ctx.cfg
        inH, inW = ctx.in_hw
        kH, kW = k.shape
        outH, outW = dy.shape[1], dy.shape[2]
        gpadx0 = kW - padx0 - 1                                    # upfirdn_2d_v2.py:205-209
        gpady0 = kH - pady0 - 1
        gpadx1 = inW * upx - outW * downx + padx0 - upx + 1
        gpady1 = inH * upy - outH * downy + pady0 - upy + 1
        gk = torch.flip(k, dims=(0, 1)).contiguous()
        dx = _UpFirDn2D.apply(dy, gk, downx, downy, upx, upy, gpadx0, gpadx1, gpady0, gpady1)
        return (dx,) + (None,) * 9


def upfirdn_2d(x: torch.Tensor, k: torch.Tensor, upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0,
               pady1=0) -> torch.Tensor:
    """upfirdn_2d_v2.py:116-163 — x: [majorDim, inH, inW, minorDim]."""
    assert x.dim() == 4 and k.dim() == 2
    return _UpFirDn2D.apply(x, k, upx, upy, downx, downy, padx0, padx1, pady0, pady1)


@lru_cache(maxsize=None)
def _fir_kernel(device_str: str, gain: float) -> torch.Tensor:
    t = torch.tensor(FIR_TAPS, dtype=torch.float64)
    k = torch.outer(t, t)
    return (k / k.sum() * gain).to(dtype=torch.float32, device=device_str)


def upsample_2d_nhwc(y: torch.Tensor) -> torch.Tensor:
    """upsample_2d (upfirdn_2d_v2.py:58-62) for an NHWC tensor [B,H,W,C]: zero-insert x2, pad 2 / 1,
    k = outer([1,3,3,1]) / 64 * 4."""
    gain, pad0, pad1 = RESAMPLE_SITES["rgb_up"]
    k = _fir_kernel(str(y.device), gain)
    return upfirdn_2d(y, k, upx=2, upy=2, padx0=pad0, padx1=pad1, pady0=pad0, pady1=pad1)


def upsample_2d_nhwc_adjoint(g: torch.Tensor) -> torch.Tensor:
    """Gradient of :func:`upsample_2d_nhwc` with respect to its input: the same op with up and down exchanged
    (upfirdn_2d_v2.py:205-246); g [B,2H,2W,C] -> [B,H,W,C]."""
    gain, pad0, pad1 = RESAMPLE_SITES["rgb_up"]
    k = _fir_kernel(str(g.device), gain)                       # symmetric: flipping it is the identity
    inH, inW = g.shape[1] // 2, g.shape[2] // 2
    gp0 = 4 - pad0 - 1
    gpx1 = inW * 2 - g.shape[2] + pad0 - 2 + 1
    gpy1 = inH * 2 - g.shape[1] + pad0 - 2 + 1
    return K.upfirdn2d(g.contiguous(), k, downx=2, downy=2, padx0=gp0, padx1=gpx1, pady0=gp0, pady1=gpy1)
