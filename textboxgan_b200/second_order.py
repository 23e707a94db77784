"""Twice-differentiable layer primitives for the regularisers (path length, training_step.py:300-347; R1, :349-373).

The plain training step runs every layer group as a first-order fused Function (:mod:`textboxgan_b200.fused`).  The two
regularisers differentiate a gradient (``tape.gradient`` inside the outer tape, training_step.py:323-333, 363-368), so
their forward pass needs nodes whose backward is again made of differentiable nodes.  The set below is CLOSED under
differentiation — every backward is expressed with members of the set — so gradients of any order stay on the CUDA
kernels and no fp32 element-wise torch pass over an activation remains (the round-1 path spent ~22 ms of a 56 ms
path-length step in such passes, profiles/r02f_timeline_c2_pl.txt):

* ``modulate(x, s)``           x * s[b,c]                (tbg_modulate)     d/dx = modulate(g, s), d/ds = rowdot(g, x)
* ``rowdot(a, b)``             sum_p a*b -> [B,C]        (tbg_rowdot)       d/da = modulate(b, g), d/db = modulate(a, g)
* ``bias_act(t, nz, ns, b)``   act(t + nz*ns + b)*gain   (tbg_bias_act_fwd) d/dt = mask_mul(g, out)
* ``mask_mul(g, out)``         g * gain * act'(out)      (tbg_bias_act_bwd) linear in g with a piecewise-constant mask:
                                                                            d/dg = mask_mul(gg, out)
* ``to_rgb(x, ws)`` / its pair of adjoints                (tbg_torgb_fwd / tbg_torgb_bwd)
* convolutions and weight gradients: :func:`textboxgan_b200.conv.conv` / ``conv_wgrad`` (already closed), or —
  ``layers.SPEC_SECOND_ORDER`` — the triple ``lin_conv`` F(x, w) / A(g, w) / G(x, g) below, which works on the MASTER weight
  and reuses the launch recipes of the first-order path (``conv.weight_spec``): prepared matrices come from the step's
  grouped weight preparation instead of torch re-layouts, the weight gradients of the resampling layers run unfolded
  (9 taps on the filtered tensor instead of 36), and nothing is re-indexed with flip / pad / copy kernels.
* ``fir4_down(x)`` / its adjoint (the skip branch of a discriminator block): linear, each the other's derivative.

The only non-linear pieces of a modulated layer are the demodulation coefficient (parameter-sized, plain torch) and the
leaky-ReLU mask (zero second derivative almost everywhere) — SURVEY.md Appendix F.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K


class _Modulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        x = x.contiguous()
        s = s.contiguous().float()
        ctx.save_for_backward(x, s)
        return K.modulate(x, s)

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        gx = modulate(g, s) if ctx.needs_input_grad[0] else None
        gs = rowdot(g, x) if ctx.needs_input_grad[1] else None
        return gx, gs


class _RowDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return K.rowdot(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = modulate(b, g) if ctx.needs_input_grad[0] else None
        gb = modulate(a, g) if ctx.needs_input_grad[1] else None
        return ga, gb


def modulate(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """x bf16 [B,...,C] * s fp32 [B,C] -> bf16."""
    return _Modulate.apply(x, s)


def rowdot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sum over the pixels of a * b: bf16 [B,...,C] x 2 -> fp32 [B,C]."""
    return _RowDot.apply(a, b)


class _MaskMul(torch.autograd.Function):
    """gt = g * gain * act'(out) (+ the per-(sample, channel) sums of gt and gt * noise for the bias / noise-strength
    gradients, as non-differentiable side outputs)."""

    @staticmethod
    def forward(ctx, g, out, noise, act: int, gain: float, want_sums: bool):
        g, out = g.contiguous(), out.contiguous()
        ctx.save_for_backward(out)
        ctx.cfg = (act, gain)
        if want_sums:
            gt, S1, _, Snz = K.bias_act_bwd(g, out, noise=noise, act=act, gain=gain, want_sums=True)
        else:
            gt, _, _, _ = K.bias_act_bwd(g, out, act=act, gain=gain, want_sums=False)
            S1 = Snz = g.new_zeros((), dtype=torch.float32)
        ctx.mark_non_differentiable(S1, Snz)
        return gt, S1, Snz

    @staticmethod
    def backward(ctx, ggt, _g1, _g2):
        (out,) = ctx.saved_tensors
        act, gain = ctx.cfg
        gg = mask_mul(ggt, out, act, gain) if ctx.needs_input_grad[0] else None
        return gg, None, None, None, None, None


def mask_mul(g: torch.Tensor, out: torch.Tensor, act: int, gain: float) -> torch.Tensor:
    return _MaskMul.apply(g, out, None, act, gain, False)[0]


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, noise, ns, bias, act: int, gain: float):
        t = t.contiguous()
        noise = noise.contiguous() if noise is not None else None
        out = K.bias_act_fwd(t, noise=noise, noise_strength=ns.reshape(1).contiguous() if noise is not None else None,
                             bias=bias.contiguous() if bias is not None else None, act=act, gain=gain)
        ctx.save_for_backward(out, noise)
        ctx.cfg = (act, gain, ns is not None and noise is not None, bias is not None, tuple(ns.shape) if ns is not None else ())
        return out

    @staticmethod
    def backward(ctx, g):
        out, noise = ctx.saved_tensors
        act, gain, has_noise, has_bias, ns_shape = ctx.cfg
        want_ns = has_noise and ctx.needs_input_grad[2]
        want_b = has_bias and ctx.needs_input_grad[3]
        gt, S1, Snz = _MaskMul.apply(g, out, noise if want_ns else None, act, gain, want_ns or want_b)
        g_ns = Snz.sum().reshape(ns_shape) if want_ns else None
        g_b = S1.sum(0) if want_b else None
        return (gt if ctx.needs_input_grad[0] else None), None, g_ns, g_b, None, None


def bias_act(t: torch.Tensor, noise: Optional[torch.Tensor], ns: Optional[torch.Tensor], bias: Optional[torch.Tensor],
             act: int, gain: float) -> torch.Tensor:
    """act(t + noise[b,y,x]*ns + bias[c]) * gain on bf16 NHWC (Noise.call noise.py:12-22 + BiasAct.call bias_act.py:25-34)."""
    return _BiasAct.apply(t, noise, ns, bias, act, gain)


# ----------------------------------------------------------------------------------------------
# ToRGB (to_rgb.py:28-33: 1x1 modulated convolution to 3 channels, no demodulation)
# ----------------------------------------------------------------------------------------------
class _ToRGB(torch.autograd.Function):
    """y[b,p,:] = x[b,p,:] @ ws[b] — x bf16 [B,H,W,C], ws fp32 [B,C,3] -> fp32 [B,H,W,3]."""

    @staticmethod
    def forward(ctx, x, ws):
        x, ws = x.contiguous(), ws.contiguous().float()
        ctx.save_for_backward(x, ws)
        return K.torgb_fwd(x, ws, None)

    @staticmethod
    def backward(ctx, gy):
        x, ws = ctx.saved_tensors
        gx, gws = _ToRGBAdjoint.apply(x, ws, gy)
        return (gx if ctx.needs_input_grad[0] else None), (gws if ctx.needs_input_grad[1] else None)


class _ToRGBAdjoint(torch.autograd.Function):
    """(gx, gws) = (gy @ ws^T, x^T @ gy) in one pass over x (tbg_torgb_bwd)."""

    @staticmethod
    def forward(ctx, x, ws, gy):
        x, ws = x.contiguous(), ws.contiguous().float()
        gy = gy.contiguous().float()
        ctx.save_for_backward(x, ws, gy)
        return K.torgb_bwd(x, ws, gy)

    @staticmethod
    def backward(ctx, ggx, ggws):
        x, ws, gy = ctx.saved_tensors
        g_x = g_ws = g_gy = None
        if ggws is not None:
            # gws = x^T gy:   d/dx = gy @ ggws^T,  d/dgy = x @ ggws
            if ctx.needs_input_grad[0]:
                g_x = _ToRGBAdjoint.apply(x, ggws, gy)[0]
            if ctx.needs_input_grad[2]:
                g_gy = to_rgb(x, ggws)
        if ggx is not None:
            # gx = gy @ ws^T: d/dws = ggx^T gy,    d/dgy = ggx @ ws
            if ctx.needs_input_grad[1]:
                g_ws = _ToRGBAdjoint.apply(ggx, ws, gy)[1]
            if ctx.needs_input_grad[2]:
                t = to_rgb(ggx, ws)
                g_gy = t if g_gy is None else g_gy + t
        return g_x, g_ws, g_gy


def to_rgb(x: torch.Tensor, ws: torch.Tensor) -> torch.Tensor:
    return _ToRGB.apply(x, ws)


# ----------------------------------------------------------------------------------------------
# Convolution layers as a closed triple of bilinear maps on the master weight w [KH,KW,I,O] (fp32):
#   F(x, w)   the layer's convolution (incl. the FIR of upsample_conv_2d / conv_downsample_2d, upfirdn_2d_v2.py:65-113)
#   A(g, w)   its adjoint in x        (= dF/dx applied to g)
#   G(x, g)   its adjoint in w        (= dF/dw applied to g, returned in the master layout, equalised-LR coefficient included)
# d/dx F = A(., w), d/dw F = G(x, .);   d/dg A = F(., w), d/dw A = G(., g);   d/dx G = A(g, .), d/dg G = F(x, .)
# so gradients of any order stay inside the set.  The launch recipes are those of the first-order Functions of fused.py:
#   plain   : 3x3 / 1x1 SAME convolution
#   downU   : F = FIR pre-pass + strided VALID convolution, A = folded 4-phase convolution straight to the input gradient,
#             G on the filtered tensor (9 taps)
#   upT     : F = FIR folded into a 4-phase 3x3 convolution, A = FIR adjoint + stride-2 VALID convolution, G = that
#             convolution's weight gradient with the roles exchanged (9 taps)
# ----------------------------------------------------------------------------------------------
class LinConv:
    def __init__(self, kind: str, H: int, W: int, I: int, O: int, k: int, reduce_height: bool = True, tag: str = ""):
        from . import conv as C

        assert kind in ("plain", "downU", "upT"), kind
        self.kind = kind
        self.spec = C.weight_spec(kind, H, W, I, O, k, reduce_height, tag)
        self.fspec = C.weight_spec("up", H, W, I, O, k, True, tag) if kind == "upT" else None

    @staticmethod
    def _mats(w, spec, want_adj: bool, is_param: bool):
        if is_param:                       # a model weight: prepared once per iteration (fused.StepWeights)
            from .fused import _prepared

            return _prepared(w, spec, want_adj, False)
        return K.wprep(w.contiguous().float(), spec, want_adj=want_adj, want_q=False)    # a cotangent in weight layout

    def F(self, x, w, is_param: bool):
        x = _bf16(x)
        if self.kind == "upT":
            wf = self._mats(w, self.fspec, False, is_param)[0]
            K.PROFILE_TAG = (self.spec.geom.tag, self.fspec.geom.algo_frac)
            return K.conv2d_igemm(x, wf, **self.fspec.geom.kernel_kwargs())
        spec = self.spec
        wm = self._mats(w, spec, False, is_param)[0]
        if spec.fir is not None:
            x = K.fir4(x, spec.fir["out_hw"], spec.fir["off"], spec.fir["scale"])
        K.PROFILE_TAG = (spec.geom.tag, spec.geom.algo_frac)
        return K.conv2d_igemm(x, wm, **spec.fwd_kwargs)

    def A(self, g, w, is_param: bool):
        g = _bf16(g)
        spec = self.spec
        wa = self._mats(w, spec, True, is_param)[1]
        if self.kind == "upT":
            g = K.fir4(g, spec.t_hw, (-2, -2), 1.0 / 16.0)
            K.PROFILE_TAG = (spec.geom.tag, 1.0)
            return K.conv2d_igemm(g, wa, **spec.s2_kwargs)
        K.PROFILE_TAG = (spec.geom.tag, spec.adj_frac)
        return K.conv2d_igemm(g, wa, **spec.adj_kwargs)

    def G(self, x, g):
        x, g = _bf16(x), _bf16(g)
        spec = self.spec
        if self.kind == "upT":
            gT = K.fir4(g, spec.t_hw, (-2, -2), 1.0 / 16.0)
            K.PROFILE_TAG = (spec.geom.tag, 1.0)
            return K.wfold_adj(K.conv2d_wgrad(gT, x, **spec.s2_kwargs), spec, flip=True)
        if spec.fir is not None:
            x = K.fir4(x, spec.fir["out_hw"], spec.fir["off"], spec.fir["scale"])
        K.PROFILE_TAG = (spec.geom.tag, spec.geom.algo_frac)
        return K.wfold(K.conv2d_wgrad(x, g, **spec.fwd_kwargs), spec)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    from . import conv as C

    return C._as_bf16(t)              # (the CPU emulation of the tests keeps its own activation dtype through this hook)


class _LinF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, lc: LinConv, is_param: bool):
        ctx.save_for_backward(x, w)
        ctx.lc, ctx.is_param = lc, is_param
        return lc.F(x, w, is_param)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = _LinA.apply(gy, w, ctx.lc, ctx.is_param) if ctx.needs_input_grad[0] else None
        gw = _LinG.apply(x, gy, ctx.lc) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None


class _LinA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, w, lc: LinConv, is_param: bool):
        ctx.save_for_backward(g, w)
        ctx.lc, ctx.is_param = lc, is_param
        return lc.A(g, w, is_param)

    @staticmethod
    def backward(ctx, ggx):
        g, w = ctx.saved_tensors
        gg = _LinF.apply(ggx, w, ctx.lc, ctx.is_param) if ctx.needs_input_grad[0] else None
        gw = _LinG.apply(ggx, g, ctx.lc) if ctx.needs_input_grad[1] else None
        return gg, gw, None, None


class _LinG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, lc: LinConv):
        ctx.save_for_backward(x, g)
        ctx.lc = lc
        return lc.G(x, g)

    @staticmethod
    def backward(ctx, cw):
        x, g = ctx.saved_tensors
        gx = _LinA.apply(g, cw, ctx.lc, False) if ctx.needs_input_grad[0] else None
        gg = _LinF.apply(x, cw, ctx.lc, False) if ctx.needs_input_grad[1] else None
        return gx, gg, None


_LIN_CACHE: dict = {}


def lin_conv(x: torch.Tensor, w_raw: torch.Tensor, kind: str, k: int, reduce_height: bool = True, tag: str = "") -> torch.Tensor:
    """The convolution of a layer (no bias / activation) on the master weight ``w_raw`` [k,k,I,O]; x bf16 [B,H,W,I]."""
    _, H, W_, _ = x.shape
    I, O = w_raw.shape[2], w_raw.shape[3]
    key = (kind, H, W_, I, O, k, bool(reduce_height), tag)
    lc = _LIN_CACHE.get(key)
    if lc is None:
        lc = _LIN_CACHE[key] = LinConv(kind, H, W_, I, O, k, reduce_height, tag)
    return _LinF.apply(x, w_raw, lc, True)


# ----------------------------------------------------------------------------------------------
# FIR of the discriminator's skip branch at the pixels its strided 1x1 convolution reads (tbg_fir4_down) and its transpose
# ----------------------------------------------------------------------------------------------
class _Fir4Down(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sy: int):
        x = _bf16(x)
        ctx.sy, ctx.hw = sy, (x.shape[1], x.shape[2])
        return K.fir4_down(x, (x.shape[1] // sy, x.shape[2] // 2), sy, (-1, -1), 1.0 / 64.0)

    @staticmethod
    def backward(ctx, g):
        return _Fir4DownAdj.apply(g, ctx.sy, ctx.hw), None


class _Fir4DownAdj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, sy: int, hw):
        ctx.sy = sy
        return K.fir4_down_adjoint(_bf16(g), hw, sy, (-1, -1), 1.0 / 64.0)

    @staticmethod
    def backward(ctx, gg):
        return _Fir4Down.apply(gg, ctx.sy), None, None


def fir4_down(x: torch.Tensor, sy: int) -> torch.Tensor:
    return _Fir4Down.apply(x, sy)
