"""Host-side convolution algebra on the CPU (kernels replaced by their documented semantics, fp64): every
formulation the training step uses is an exact linear map with the adjoint / equivalence it claims.
  * <conv_g(x, W), y> == <x, conv_{g^T}(y, relayout(W))>  for s1 / s2 / up geometries (conv.py::ConvGeom.adjoint),
  * unfolded upsample_conv_2d (transposed conv with tap masks + FIR) == folded 4-phase GEMM == the literal
    reference sequence (oracle: conv2d_transpose of the flipped kernel, then upfirdn), forward and adjoint,
  * unfolded conv_downsample_2d (FIR + strided conv) == folded (k+3)x(k+3) conv, and wprep/wfold are transposes."""
import math

import pytest
import torch

import emu
from oracle import stylegan as OS
from textboxgan_b200 import conv as C


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


@pytest.mark.parametrize("name,g", [
    ("s1", C.plain_geom(6, 10, 8, 16, 3)),
    ("s1-1x1", C.plain_geom(4, 6, 8, 8, 1)),
    ("up", C.up_geom(3, 5, 8, 8)),
    ("down3", C.down_geom(8, 12, 8, 16, 3, True)),
    ("down1", C.down_geom(8, 12, 8, 8, 1, True)),
    ("down3-w", C.down_geom(6, 8, 8, 8, 3, False)),
])
def test_adjoint_identity_of_every_geometry(name, g):
    x = _rand(2, g.H, g.W, g.cin, seed=1)
    w = _rand(g.n_total, g.k_total, seed=2)
    y = emu.emu_conv2d_igemm(x, w, out_fp32=True, **g.kernel_kwargs()).double()
    gy = _rand(*y.shape, seed=3)
    a = g.adjoint()
    gx = emu.emu_conv2d_igemm(gy, C.relayout_for_adjoint(w, g), out_fp32=True, **a.kernel_kwargs()).double()
    assert gx.shape == x.shape
    lhs, rhs = (y * gy).sum(), (x * gx).sum()
    assert abs(lhs - rhs) <= 1e-6 * (abs(lhs) + 1)       # the emulation returns fp32 tensors
    # weight gradient is the third face of the same trilinear form
    gw = emu.emu_conv2d_wgrad(x, gy, **g.kernel_kwargs()).double()
    assert abs((gw * w).sum() - lhs) <= 1e-6 * (abs(lhs) + 1)


@pytest.mark.parametrize("h,w,I,O", [(3, 5, 64, 64), (4, 8, 64, 128)])
def test_unfolded_upsample_conv_equals_folded_and_reference_sequence(h, w, I, O):
    B = 2
    x = _rand(B, h, w, I, seed=4)
    wr = _rand(3, 3, I, O, seed=5)
    up, upT = C.weight_spec("up", h, w, I, O, 3, True, "t"), C.weight_spec("upT", h, w, I, O, 3, True, "t")
    import textboxgan_b200.layers as L
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float64
    try:
        f_fold, _, _ = emu.emu_wprep(wr, up, want_adj=False)
        f_unf, a_unf, _ = emu.emu_wprep(wr, upT, want_adj=True)
    finally:
        L.ACT_DTYPE = old
    y_fold = emu.emu_conv2d_igemm(x, f_fold.double(), out_fp32=True, **up.geom.kernel_kwargs()).double()
    T = emu.emu_conv2d_igemm(x, f_unf.double(), out_fp32=True, **upT.fwd_kwargs).double()
    y_unf = emu.emu_fir4(T, upT.out_hw, (-1, -1), 1.0 / 16.0).double()
    assert (y_fold - y_unf).abs().max() < 1e-5 * y_fold.abs().max()
    # the literal reference sequence (upfirdn_2d_v2.py:65-103) on NCHW
    kk, p0, p1 = OS.compute_paddings([1, 3, 3, 1], up=True, down=False, is_conv=True)
    ref = OS.upsample_conv_2d(x.permute(0, 3, 1, 2).float(), w, h, (wr * up.coef).float(), p0, p1, kk)
    ref = ref.permute(0, 2, 3, 1).double()
    assert (y_unf - ref).abs().max() < 1e-4 * ref.abs().max()
    # adjoint of the unfolded pair = FIR adjoint + stride-2 conv with the adjoint-layout matrix
    gy = _rand(*y_unf.shape, seed=6)
    gT = emu.emu_fir4(gy, upT.t_hw, (-2, -2), 1.0 / 16.0).double()
    gx = emu.emu_conv2d_igemm(gT, a_unf.double(), out_fp32=True, **upT.s2_kwargs).double()
    lhs, rhs = (y_unf * gy).sum(), (x * gx).sum()
    assert abs(lhs - rhs) <= 1e-6 * (abs(lhs) + 1)
    # role-swapped weight gradient folded back onto the master weight == d<y, gy>/dw
    gadj = emu.emu_conv2d_wgrad(gT, x, **upT.s2_kwargs).double()
    gw = emu.emu_wfold_adj(gadj, upT, flip=True).double()
    eps_dir = _rand(3, 3, I, O, seed=7)
    L.ACT_DTYPE = torch.float64
    try:
        f_dir, _, _ = emu.emu_wprep(eps_dir, upT, want_adj=False)
    finally:
        L.ACT_DTYPE = old
    dT = emu.emu_conv2d_igemm(x, f_dir.double(), out_fp32=True, **upT.fwd_kwargs).double()
    dy = emu.emu_fir4(dT, upT.out_hw, (-1, -1), 1.0 / 16.0).double()
    assert abs((dy * gy).sum() - (gw * eps_dir).sum()) <= 1e-6 * (abs((dy * gy).sum()) + 1)


@pytest.mark.parametrize("k,rh", [(3, True), (1, True), (3, False), (1, False)])
def test_unfolded_downsample_conv_equals_folded(k, rh):
    B, H, W, I, O = 2, 8, 12, 64, 64
    x = _rand(B, H, W, I, seed=8)
    wr = _rand(k, k, I, O, seed=9)
    fo, un = C.weight_spec("down", H, W, I, O, k, rh, "t"), C.weight_spec("downU", H, W, I, O, k, rh, "t")
    import textboxgan_b200.layers as L
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float64
    try:
        f_fold, a_fold, _ = emu.emu_wprep(wr, fo, want_adj=True)
        f_unf, a_unf, _ = emu.emu_wprep(wr, un, want_adj=True)
    finally:
        L.ACT_DTYPE = old
    y_fold = emu.emu_conv2d_igemm(x, f_fold.double(), out_fp32=True, **fo.geom.kernel_kwargs()).double()
    xb = emu.emu_fir4(x, un.fir["out_hw"], un.fir["off"], un.fir["scale"]).double()
    y_unf = emu.emu_conv2d_igemm(xb, f_unf.double(), out_fp32=True, **un.fwd_kwargs).double()
    assert y_fold.shape == y_unf.shape and (y_fold - y_unf).abs().max() < 1e-5 * y_fold.abs().max()
    assert (a_fold.double() - a_unf.double()).abs().max() == 0.0          # the input gradient stays folded
    # literal reference sequence
    kk, p0, p1 = OS.compute_paddings([1, 3, 3, 1], up=False, down=True, is_conv=True, convW=k)
    ref = OS.conv_downsample_2d(x.permute(0, 3, 1, 2).float(), H, W, (wr * fo.coef).float(), p0, p1, kk, rh)
    ref = ref.permute(0, 2, 3, 1).double()
    assert (y_unf - ref).abs().max() < 1e-4 * ref.abs().max()
    # weight gradient on the filtered tensor folded with the plain tables == d<y, gy>/dw
    gy = _rand(*y_unf.shape, seed=10)
    gw = emu.emu_wfold(emu.emu_conv2d_wgrad(xb, gy, **un.fwd_kwargs).double(), un).double()
    d = _rand(k, k, I, O, seed=11)
    L.ACT_DTYPE = torch.float64
    try:
        f_d, _, _ = emu.emu_wprep(d, fo, want_adj=False)
    finally:
        L.ACT_DTYPE = old
    dy = emu.emu_conv2d_igemm(x, f_d.double(), out_fp32=True, **fo.geom.kernel_kwargs()).double()
    assert abs((dy * gy).sum() - (gw * d).sum()) <= 1e-6 * (abs((dy * gy).sum()) + 1)
