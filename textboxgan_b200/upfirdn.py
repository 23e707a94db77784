"""upfirdn_2d and its wrappers (mirror of models/custom_stylegan2/layers/upfirdn/upfirdn_2d_v2.py)
over the ``tbg_upfirdn2d`` kernel.

Like ``upfirdn_2d_cuda`` (upfirdn_2d_v2.py:186-246) the gradient of the op is another call of the
same op with up/down swapped, the flipped kernel and pads ``gpad0 = k - pad0 - 1``,
``gpad1 = in*up - out*down + pad0 - up + 1`` — so gradients of any order are available.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch

from . import kernels as K


def _setup_kernel(k) -> np.ndarray:
    """upfirdn_2d_v2.py:18-25"""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k


def compute_paddings(resample_kernel, up, down, is_conv, convW=3, factor=2, gain=1):
    """upfirdn_2d_v2.py:28-55"""
    assert not (up and down)
    k = [1] * factor if resample_kernel is None else resample_kernel
    if up:
        k = _setup_kernel(k) * (gain * (factor ** 2))
        if is_conv:
            p = (k.shape[0] - factor) - (convW - 1)
            pad0 = (p + 1) // 2 + factor - 1
            pad1 = p // 2 + 1
        else:
            p = k.shape[0] - factor
            pad0 = (p + 1) // 2 + factor - 1
            pad1 = p // 2
    elif down:
        k = _setup_kernel(k) * gain
        if is_conv:
            p = (k.shape[0] - factor) + (convW - 1)
            pad0 = (p + 1) // 2
            pad1 = p // 2 + 1
        else:
            p = k.shape[0] - factor
            pad0 = (p + 1) // 2
            pad1 = p // 2
    else:
        k = resample_kernel
        pad0, pad1 = 0, 0
    return k, pad0, pad1


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
    k = _setup_kernel([1, 3, 3, 1]) * gain
    return torch.as_tensor(k, dtype=torch.float32, device=device_str)


def upsample_2d_nhwc(y: torch.Tensor, factor: int = 2) -> torch.Tensor:
    """upsample_2d (upfirdn_2d_v2.py:58-62) for an NHWC tensor [B,H,W,C]: k = [1,3,3,1]^2/64 * 4,
    pad0 = 2, pad1 = 1 (compute_paddings(up=True, is_conv=False))."""
    _, pad0, pad1 = compute_paddings([1, 3, 3, 1], up=True, down=False, is_conv=False)
    k = _fir_kernel(str(y.device), float(factor ** 2))
    return upfirdn_2d(y, k, upx=factor, upy=factor, padx0=pad0, padx1=pad1, pady0=pad0, pady1=pad1)
