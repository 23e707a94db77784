"""Thin Python callers of the C ABI: torch tensors in (device memory + current stream only),
raw pointers out.  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.TbgError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.TbgError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.TbgError(f"{name}: expected a contiguous tensor")


def conv2d_igemm(
    x: torch.Tensor,            # bf16 [B, H, W, Cin]
    w: torch.Tensor,            # bf16 [n_total, taps_h*taps_w*Cin]
    *,
    Ho: int,
    Wo: int,
    taps: tuple[int, int],
    pad: tuple[int, int],
    stride: tuple[int, int] = (1, 1),
    up: bool = False,
    col_scale: Optional[torch.Tensor] = None,   # fp32 [B, cout]
    bias: Optional[torch.Tensor] = None,        # fp32 [cout]
    noise: Optional[torch.Tensor] = None,       # fp32 [B, out_H, out_W]
    noise_strength: Optional[torch.Tensor] = None,  # fp32 scalar tensor
    residual: Optional[torch.Tensor] = None,    # bf16, shape of out
    res_scale: float = 1.0,
    act: int = 0,
    act_gain: float = 1.0,
    out_fp32: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    _require(x, torch.bfloat16, "x")
    _require(w, torch.bfloat16, "w")
    B, H, W_, Cin = x.shape
    n_total = w.shape[0]
    cout = n_total // 4 if up else n_total
    oH, oW = (2 * Ho, 2 * Wo) if up else (Ho, Wo)
    if w.shape[1] != taps[0] * taps[1] * Cin:
        raise _lib.TbgError(f"w: expected K={taps[0] * taps[1] * Cin}, got {w.shape[1]}")
    if out is None:
        out = torch.empty((B, oH, oW, cout), device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    for t, n in ((col_scale, "col_scale"), (bias, "bias"), (noise, "noise"), (noise_strength, "noise_strength")):
        if t is not None:
            _require(t, torch.float32, n)
    if residual is not None:
        _require(residual, torch.bfloat16, "residual")
    a = _lib.ConvArgs(
        x=_ptr(x), w=_ptr(w), out=_ptr(out),
        B=B, H=H, W=W_, Cin=Cin, Ho=Ho, Wo=Wo, n_total=n_total, cout=cout,
        taps_h=taps[0], taps_w=taps[1], pad_h=pad[0], pad_w=pad[1],
        stride_h=stride[0], stride_w=stride[1], up=int(up),
        col_scale=_ptr(col_scale), bias=_ptr(bias), noise=_ptr(noise), noise_strength=_ptr(noise_strength),
        residual=_ptr(residual), res_scale=res_scale, act=act, act_gain=act_gain, out_fp32=int(out_fp32),
    )
    _lib.check(_lib.load().tbg_conv2d_igemm(C.byref(a), _stream()), "tbg_conv2d_igemm")
    return out


def conv2d_wgrad(
    x: torch.Tensor,            # bf16 [B, H, W, Cin]
    gy: torch.Tensor,           # bf16 [B, gy_H, gy_W, cout]
    *,
    Ho: int,
    Wo: int,
    taps: tuple[int, int],
    pad: tuple[int, int],
    stride: tuple[int, int] = (1, 1),
    up: bool = False,
    gw: Optional[torch.Tensor] = None,          # fp32 [n_total, taps*Cin], accumulated into
) -> torch.Tensor:
    _require(x, torch.bfloat16, "x")
    _require(gy, torch.bfloat16, "gy")
    B, H, W_, Cin = x.shape
    cout = gy.shape[3]
    n_total = 4 * cout if up else cout
    if gw is None:
        gw = torch.zeros((n_total, taps[0] * taps[1] * Cin), device=x.device, dtype=torch.float32)
    _require(gw, torch.float32, "gw")
    a = _lib.WgradArgs(
        x=_ptr(x), gy=_ptr(gy), gw=_ptr(gw),
        B=B, H=H, W=W_, Cin=Cin, Ho=Ho, Wo=Wo, n_total=n_total, cout=cout,
        taps_h=taps[0], taps_w=taps[1], pad_h=pad[0], pad_w=pad[1],
        stride_h=stride[0], stride_w=stride[1], up=int(up),
    )
    _lib.check(_lib.load().tbg_conv2d_wgrad(C.byref(a), _stream()), "tbg_conv2d_wgrad")
    return gw
