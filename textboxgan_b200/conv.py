"""Convolution geometry algebra + autograd bindings of the tcgen05 implicit-GEMM kernels.

Every convolution on the TextBoxGAN path is one *geometry*: per spatial axis either

* ``s1``  out[p]      = sum_u in[p + u - pad]      * W[u]        (stride 1),
* ``s2``  out[p]      = sum_u in[2p + u - pad]     * W[u]        (stride 2),
* ``up``  out[2q+phi] = sum_t in[q + t - pad]      * W[phi, t]   (2-phase transposed conv).

The family is closed under transposition (s1 <-> s1 with flipped taps, s2 <-> up), so forward,
input-gradient and their second-order terms all run on the same ``tbg_conv2d_igemm`` kernel, and
weight gradients on ``tbg_conv2d_wgrad``.  Reference sites: ModulatedConv2D.call
(modulated_conv2d.py:66-122), upsample_conv_2d (upfirdn_2d_v2.py:65-103: transposed 3x3 conv +
4x4 FIR folded into a 4-phase 3x3 GEMM), conv_downsample_2d (upfirdn_2d_v2.py:106-113: 4x4 FIR +
strided conv folded into a 6x6 / 4x4 stride-2 conv), Conv2D.call (conv.py:51-73).

The three autograd Functions (conv, conv-of-cotangent, weight-gradient) call each other in their
backward passes, so gradients of any order are available — required by the path-length and R1
regularisers (training_step.py:323-333, 363-368).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional, Tuple

import torch

from . import kernels as K


@dataclass(frozen=True)
class Axis:
    kind: str   # 's1' | 's2' | 'up'
    k: int      # taps (per phase for 'up')
    pad: int

    @property
    def phases(self) -> int:
        return 2 if self.kind == "up" else 1

    def out_size(self, n_in: int) -> int:
        if self.kind == "up":
            return 2 * n_in
        if self.kind == "s2":
            assert n_in % 2 == 0
            return n_in // 2
        return n_in

    def adjoint(self) -> "Axis":
        if self.kind == "s1":
            return Axis("s1", self.k, self.k - 1 - self.pad)
        if self.kind == "s2":
            # gx[2q+phi] = sum_t gy[q + t - 1] * W[u], u = phi + pad + 2 - 2t  (zero if u outside [0,k))
            assert self.k <= 6 and 2 * self.pad + 2 >= self.k - 1, "s2 adjoint needs 3 taps/phase"
            return Axis("up", 3, 1)
        # up(T,P) -> s2(2T, 2(T-1-P)),  W'[u] = W[phi=u%2, t=T-1-u//2]
        return Axis("s2", 2 * self.k, 2 * (self.k - 1 - self.pad))


@dataclass(frozen=True)
class ConvGeom:
    """One convolution: input grid (H, W), channel counts and the two axis specs."""
    H: int
    W: int
    cin: int
    cout: int
    ah: Axis
    aw: Axis
    tag: str = ""            # profiling label ("modconv", "dconv", "aster", ...)
    algo_frac: float = 1.0   # algorithmic FLOPs / executed GEMM FLOPs (SURVEY.md §8d accounting)

    @property
    def out_hw(self) -> Tuple[int, int]:
        return self.ah.out_size(self.H), self.aw.out_size(self.W)

    @property
    def n_total(self) -> int:
        return self.cout * self.ah.phases * self.aw.phases

    @property
    def k_total(self) -> int:
        return self.ah.k * self.aw.k * self.cin

    @property
    def grid(self) -> Tuple[int, int]:
        """GEMM row grid: output pixels, or input pixels along an 'up' axis."""
        oh, ow = self.out_hw
        return (self.H if self.ah.kind == "up" else oh, self.W if self.aw.kind == "up" else ow)

    def adjoint(self) -> "ConvGeom":
        oh, ow = self.out_hw
        return ConvGeom(oh, ow, self.cout, self.cin, self.ah.adjoint(), self.aw.adjoint(), self.tag, self.algo_frac)

    def kernel_kwargs(self) -> dict:
        gh, gw = self.grid
        return dict(
            Ho=gh, Wo=gw, taps=(self.ah.k, self.aw.k), pad=(self.ah.pad, self.aw.pad),
            stride=(2 if self.ah.kind == "s2" else 1, 2 if self.aw.kind == "s2" else 1),
            up=(int(self.ah.kind == "up"), int(self.aw.kind == "up")),
        )

    def flops(self, batch: int) -> float:
        gh, gw = self.grid
        return 2.0 * batch * gh * gw * self.n_total * self.k_total


# ----------------------------------------------------------------------------------------------
# weight re-layout for the adjoint geometry (differentiable torch indexing on parameter-sized
# tensors; activations never pass through here)
# ----------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _axis_adjoint_map(kind: str, k: int, pad: int):
    """Index/mask tables mapping the flattened (phase, tap) axis of W to that of the adjoint."""
    src = Axis(kind, k, pad)
    dst = src.adjoint()
    idx = torch.zeros(dst.phases, dst.k, dtype=torch.long)
    mask = torch.zeros(dst.phases, dst.k)
    if kind == "s1":
        for u in range(k):
            idx[0, u] = k - 1 - u
            mask[0, u] = 1.0
    elif kind == "s2":
        for phi in range(2):
            for t in range(3):
                u = phi + pad + 2 - 2 * t
                if 0 <= u < k:
                    idx[phi, t] = u
                    mask[phi, t] = 1.0
    else:  # up(T=k, P=pad): W'[u] = W[phi=u%2, t=T-1-u//2]; source flat index = phi*T + t
        for u in range(2 * k):
            idx[0, u] = (u % 2) * k + (k - 1 - u // 2)
            mask[0, u] = 1.0
    return idx.flatten(), mask.flatten()


_DEVICE_CACHE: dict = {}


def _on_device(key, device, make):
    """Constant tables are uploaded once per device (never during a CUDA-graph capture)."""
    k = (key, str(device))
    t = _DEVICE_CACHE.get(k)
    if t is None:
        t = make()
        t = tuple(x.to(device) for x in t) if isinstance(t, tuple) else t.to(device)
        _DEVICE_CACHE[k] = t
    return t


def _tables_on(device, kind, k, pad):
    return _on_device(("adj", kind, k, pad), device, lambda: _axis_adjoint_map(kind, k, pad))


def _axis_relayout(w: torch.Tensor, dim: int, ax: Axis) -> torch.Tensor:
    """Re-index the combined (phase, tap) axis ``dim`` of ``w`` for the adjoint of ``ax`` using only
    pad / reshape / flip / transpose (cheap, differentiable; equals the index tables of
    ``_axis_adjoint_map``)."""
    shp = list(w.shape)
    if ax.kind == "s1":                                   # W'[u] = W[k-1-u]
        return w.flip(dim)
    if ax.kind == "up":                                   # [phi, t] -> u = 2(T-1-t) + phi
        T = ax.k
        v = w.reshape(shp[:dim] + [2, T] + shp[dim + 1:]).flip(dim + 1).transpose(dim, dim + 1)
        return v.reshape(shp[:dim] + [2 * T] + shp[dim + 1:])
    # s2(k, pad) -> up(3, 1): u' = u + (2 - pad) = phi + 4 - 2t on a zero-padded 6-tap kernel
    front = 2 - ax.pad
    back = 6 - ax.k - front
    assert front >= 0 and back >= 0, "s2 adjoint needs k <= 6 and pad <= 2"
    if front or back:
        pads = [0, 0] * (w.dim() - 1 - dim) + [front, back]
        w = torch.nn.functional.pad(w, pads)
    v = w.reshape(shp[:dim] + [3, 2] + shp[dim + 1:]).flip(dim).transpose(dim, dim + 1)   # [phi, t]
    return v.reshape(shp[:dim] + [6] + shp[dim + 1:])


def relayout_for_adjoint(wmat: torch.Tensor, g: ConvGeom) -> torch.Tensor:
    """[n_total, K] weights of ``g`` -> [n_total', K'] weights of ``g.adjoint()``."""
    ph, pw = g.ah.phases, g.aw.phases
    w6 = wmat.reshape(ph, pw, g.cout, g.ah.k, g.aw.k, g.cin)
    # bring (phase_h, tap_h) and (phase_w, tap_w) together: [O, I, ph*kh, pw*kw]
    w4 = w6.permute(2, 5, 0, 3, 1, 4).reshape(g.cout, g.cin, ph * g.ah.k, pw * g.aw.k)
    w4 = _axis_relayout(w4, 2, g.ah)
    w4 = _axis_relayout(w4, 3, g.aw)
    a = g.adjoint()
    # -> [ph', pw', I(as out), kh', kw', O(as in)]
    w6a = w4.reshape(g.cout, g.cin, a.ah.phases, a.ah.k, a.aw.phases, a.aw.k).permute(2, 4, 1, 3, 5, 0)
    return w6a.reshape(a.n_total, a.k_total)


# ----------------------------------------------------------------------------------------------
# autograd Functions
# ----------------------------------------------------------------------------------------------
def _as_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.bfloat16 else t.to(torch.bfloat16).contiguous()


class _ConvFn(torch.autograd.Function):
    """y = conv_g(x, W).  x bf16 NHWC, W fp32 [n_total, K] (rounded to bf16 for the tensor cores)."""

    @staticmethod
    def forward(ctx, x, wmat, geom: ConvGeom, epi: Optional[dict]):
        ctx.geom = geom
        ctx.save_for_backward(x, wmat)
        K.PROFILE_TAG = (geom.tag, geom.algo_frac)
        return K.conv2d_igemm(_as_bf16(x), _as_bf16(wmat), **geom.kernel_kwargs(), **(epi or {}))

    @staticmethod
    def backward(ctx, gy):
        x, wmat = ctx.saved_tensors
        g = ctx.geom
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = conv(gy, relayout_for_adjoint(wmat, g), g.adjoint())
        if ctx.needs_input_grad[1]:
            gw = conv_wgrad(x, gy, g)
        return gx, gw, None, None


class _WgradFn(torch.autograd.Function):
    """gW = wgrad_g(x, gy): fp32 [n_total, K]."""

    @staticmethod
    def forward(ctx, x, gy, geom: ConvGeom):
        ctx.geom = geom
        ctx.save_for_backward(x, gy)
        K.PROFILE_TAG = (geom.tag, geom.algo_frac)
        return K.conv2d_wgrad(_as_bf16(x), _as_bf16(gy), **geom.kernel_kwargs())

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        g = ctx.geom
        gx = ggy = None
        if ctx.needs_input_grad[0]:
            gx = conv(gy, relayout_for_adjoint(ggw, g), g.adjoint())
        if ctx.needs_input_grad[1]:
            ggy = conv(x, ggw, g)
        return gx, ggy, None


def conv(x: torch.Tensor, wmat: torch.Tensor, geom: ConvGeom, epi: Optional[dict] = None) -> torch.Tensor:
    """Differentiable (any order) convolution; ``epi`` (fused epilogue kwargs of
    kernels.conv2d_igemm) may only be used where no gradient flows through the epilogue terms."""
    assert x.shape[1] == geom.H and x.shape[2] == geom.W and x.shape[3] == geom.cin, (tuple(x.shape), geom)
    assert wmat.shape == (geom.n_total, geom.k_total), (tuple(wmat.shape), geom)
    return _ConvFn.apply(x, wmat, geom, epi)


def conv_wgrad(x: torch.Tensor, gy: torch.Tensor, geom: ConvGeom) -> torch.Tensor:
    return _WgradFn.apply(x, gy, geom)


# ----------------------------------------------------------------------------------------------
# weight preparation: reference HWIO weights -> GEMM matrices of each geometry
# ----------------------------------------------------------------------------------------------
FIR_1D = (1.0, 3.0, 3.0, 1.0)   # resample_kernel [1,3,3,1] (synthesis_block.py:36, discriminator.py:44)


def plain_geom(H: int, W: int, cin: int, cout: int, k: int, tag: str = "", algo_frac: float = 1.0) -> ConvGeom:
    """SAME stride-1 k x k convolution (modulated_conv2d.py:110-112, conv.py:69-71)."""
    return ConvGeom(H, W, cin, cout, Axis("s1", k, k // 2), Axis("s1", k, k // 2), tag, algo_frac)


def plain_wmat(w_hwio: torch.Tensor) -> torch.Tensor:
    kh, kw, I, O = w_hwio.shape
    return w_hwio.permute(3, 0, 1, 2).reshape(O, kh * kw * I)


def up_geom(h: int, w: int, cin: int, cout: int, tag: str = "") -> ConvGeom:
    """upsample_conv_2d (upfirdn_2d_v2.py:65-103) as a 4-phase 3x3 GEMM over the input grid.
    Algorithmic work is the transposed conv (9*I*O MACs per input pixel); the 4-phase GEMM with
    the FIR folded in executes 4x that."""
    return ConvGeom(h, w, cin, cout, Axis("up", 3, 1), Axis("up", 3, 1), tag, 0.25)


@lru_cache(maxsize=None)
def _up_coef() -> torch.Tensor:
    """C[phi, t, kh] = k1[1 - kh - phi + 2t] with k1 = [1,3,3,1]/4 (FIR gain 4 = 2 per axis,
    pad0 = pad1 = 1: compute_paddings(up, is_conv), upfirdn_2d_v2.py:36-39)."""
    k1 = [v / 4.0 for v in FIR_1D]
    c = torch.zeros(2, 3, 3)
    for phi in range(2):
        for t in range(3):
            for kh in range(3):
                j = 1 - kh - phi + 2 * t
                if 0 <= j < 4:
                    c[phi, t, kh] = k1[j]
    return c


def up_wmat(w_hwio: torch.Tensor) -> torch.Tensor:
    """Weff[(py,px,o), (ty,tx,i)] = sum_{kh,kw} C[py,ty,kh] C[px,tx,kw] w[kh,kw,i,o]."""
    c = _on_device(("upcoef",), w_hwio.device, _up_coef).to(w_hwio.dtype)
    weff = torch.einsum("pak,qbl,klio->pqoabi", c, c, w_hwio)
    O, I = w_hwio.shape[3], w_hwio.shape[2]
    return weff.reshape(4 * O, 9 * I)


def down_geom(H: int, W: int, cin: int, cout: int, k: int, reduce_height: bool, tag: str = "") -> ConvGeom:
    """conv_downsample_2d (upfirdn_2d_v2.py:106-113): 4x4 FIR (pad0 = (k+1)//2 + ..., A.4) followed
    by a VALID k x k conv of stride (2|1, 2), folded into one (k+3) x (k+3) convolution."""
    kk = k + 3
    pad0 = (2 + (k - 1) + 1) // 2            # compute_paddings(down, is_conv): p=(4-2)+(k-1); pad0=(p+1)//2
    aw = Axis("s2", kk, pad0)
    ah = Axis("s2", kk, pad0) if reduce_height else Axis("s1", kk, pad0)
    return ConvGeom(H, W, cin, cout, ah, aw, tag, float(k * k) / float(kk * kk))


@lru_cache(maxsize=None)
def _fold_table(k: int) -> torch.Tensor:
    """S[u, t] = kf[u - t] with kf = [1,3,3,1]/8 (FIR gain 1)."""
    kf = [v / 8.0 for v in FIR_1D]
    s = torch.zeros(k + 3, k)
    for u in range(k + 3):
        for t in range(k):
            if 0 <= u - t < 4:
                s[u, t] = kf[u - t]
    return s


def down_wmat(w_hwio: torch.Tensor) -> torch.Tensor:
    """G[uy,ux] = sum_{ty,tx} kf[uy-ty] kf[ux-tx] w[ty,tx]  ->  [O, (k+3)^2 * I]."""
    k = w_hwio.shape[0]
    s = _on_device(("fold", k), w_hwio.device, lambda: _fold_table(k)).to(w_hwio.dtype)
    g = torch.einsum("ut,vs,tsio->ouvi", s, s, w_hwio)
    O, I = w_hwio.shape[3], w_hwio.shape[2]
    return g.reshape(O, (k + 3) * (k + 3) * I)


# ----------------------------------------------------------------------------------------------
# WeightSpec: the linear map master weight -> GEMM matrices of a geometry, as per-axis tables for
# the tbg_wprep / tbg_wfold kernels
# ----------------------------------------------------------------------------------------------
def _axis_fwd_table(kind: str, k_master: int) -> torch.Tensor:
    """F[phase, tap, master_tap] of one axis."""
    if kind == "plain":
        return torch.eye(k_master).reshape(1, k_master, k_master)
    if kind == "up":
        assert k_master == 3
        return _up_coef().clone()
    if kind == "down":
        return _fold_table(k_master).reshape(1, k_master + 3, k_master).clone()
    raise ValueError(kind)


def _axis_adj_table(F: torch.Tensor, ax: Axis) -> torch.Tensor:
    idx, mask = _axis_adjoint_map(ax.kind, ax.k, ax.pad)
    a = ax.adjoint()
    flat = F.reshape(F.shape[0] * F.shape[1], F.shape[2])
    return (flat[idx] * mask[:, None]).reshape(a.phases, a.k, F.shape[2])


class WeightSpec:
    """Everything tbg_wprep / tbg_wfold need for one (geometry, master-weight shape) pair."""

    def __init__(self, geom: ConvGeom, kind_h: str, kind_w: str, KH: int, KW: int, I: int, O: int, coef: float,
                 tables=None, adj_shape=None):
        import ctypes

        self.geom, self.KH, self.KW, self.I, self.O, self.coef = geom, KH, KW, I, O, float(coef)
        self.Ipad, self.Opad = geom.cin, geom.cout
        if tables is not None:
            fy, fx, ay, ax = tables
        else:
            fy, fx = _axis_fwd_table(kind_h, KH), _axis_fwd_table(kind_w, KW)
            ay, ax = _axis_adj_table(fy, geom.ah), _axis_adj_table(fx, geom.aw)
        assert fy.shape[:2] == (geom.ah.phases, geom.ah.k) and fx.shape[:2] == (geom.aw.phases, geom.aw.k)
        self.tables = (fy, fx, ay, ax)

        def pack(ts):
            buf = []
            for t in ts:
                blk = [float(t.shape[0]), float(t.shape[1]), float(t.shape[2])] + [float(v) for v in t.flatten()]
                buf += blk + [0.0] * (39 - len(blk))
            return (ctypes.c_float * len(buf))(*buf)

        zero = torch.zeros(0, 1, max(KH, 1))
        self.ctable = pack([fy, fx, ay, ax])
        self.ctable_noadj = pack([fy, fx, torch.zeros(0, 0, KH), torch.zeros(0, 0, KW)])
        self.fir = None                      # optional FIR pre-pass of the forward input (tbg_fir4 arguments)
        self.fwd_kwargs = geom.kernel_kwargs()
        self.adj_frac = geom.algo_frac
        try:
            self.adj_kwargs = geom.adjoint().kernel_kwargs() if adj_shape is None else None
        except AssertionError:
            self.adj_kwargs = None
        self.fwd_rows, self.fwd_cols = geom.n_total, geom.k_total
        if adj_shape is not None:
            self.adj_rows, self.adj_cols = adj_shape
        else:
            a = geom.adjoint()
            self.adj_rows, self.adj_cols = a.n_total, a.k_total


@lru_cache(maxsize=None)
def weight_spec(kind: str, H: int, W: int, I: int, O: int, k: int, reduce_height: bool = True, tag: str = "",
                scale: float = 1.0) -> WeightSpec:
    """kind: 'plain' | 'up' | 'down'.  Channel counts are padded to multiples of 64 (K blocks of the
    tensor-core kernels); coef = equalised-LR runtime coefficient 1/sqrt(k*k*I) (commons.py:4-12) x scale."""
    Ip, Op = (I + 63) // 64 * 64, (O + 63) // 64 * 64
    coef = scale / math.sqrt(k * k * I)
    if kind == "plain":
        g = plain_geom(H, W, Ip, Op, k, tag=tag, algo_frac=(I * O) / float(Ip * Op))
        return WeightSpec(g, "plain", "plain", k, k, I, O, coef)
    if kind == "up":
        g = up_geom(H, W, Ip, Op, tag=tag)
        return WeightSpec(g, "up", "up", k, k, I, O, coef)
    if kind == "down":
        g = down_geom(H, W, Ip, Op, k, reduce_height, tag=tag)
        return WeightSpec(g, "down" if reduce_height else "down", "down", k, k, I, O, coef)
    if kind == "upT":
        return _upT_spec(H, W, I, O, Ip, Op, k, coef, tag)
    if kind == "downU":
        return _downU_spec(H, W, I, O, Ip, Op, k, reduce_height, coef, scale, tag)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# conv_downsample_2d with the FIR unfolded in the FORWARD and WEIGHT-GRADIENT directions
# (upfirdn_2d_v2.py:106-113 in the reference's order): xb = FIR(x) once (tbg_fir4, saved for backward), then the
# VALID k x k stride-(2|1, 2) convolution and its weight gradient run at their algorithmic cost (k*k taps
# instead of (k+3)^2).  The INPUT gradient keeps the folded form (one 4-phase GEMM straight from the output
# gradient to the input gradient, no FIR-adjoint pass).
#   k = 3: xb[j] = sum_m kf[m] x[j+m-2], (H+2) x (W+2);  y[p] = sum_t xb[s*p + t] w[t]
#   k = 1: xb[j] = sum_m kf[m] x[j+m-1],  H x W;          y[p] = xb[s*p] w
# ----------------------------------------------------------------------------------------------
def _downU_spec(H: int, W: int, I: int, O: int, Ip: int, Op: int, k: int, reduce_height: bool, coef: float,
                scale: float, tag: str) -> "WeightSpec":
    folded = weight_spec("down", H, W, I, O, k, reduce_height, tag, scale)
    ext = 2 if k == 3 else 0
    sh = 2 if reduce_height else 1
    g = ConvGeom(H + ext, W + ext, Ip, Op, Axis("s2" if reduce_height else "s1", k, 0), Axis("s2", k, 0), tag,
                 (I * O) / float(Ip * Op))
    ident = torch.eye(k).reshape(1, k, k)
    a = folded.geom.adjoint()
    spec = WeightSpec(g, "plain", "plain", k, k, I, O, coef, tables=(ident, ident.clone(), folded.tables[2], folded.tables[3]),
                      adj_shape=(a.n_total, a.k_total))
    spec.fwd_kwargs = dict(Ho=H // sh, Wo=W // 2, taps=(k, k), pad=(0, 0), stride=(sh, 2), up=(0, 0))
    spec.adj_kwargs = a.kernel_kwargs()
    spec.adj_frac = folded.geom.algo_frac
    spec.fir = dict(out_hw=(H + ext, W + ext), off=(-2, -2) if k == 3 else (-1, -1), scale=1.0 / 64.0)
    return spec


# ----------------------------------------------------------------------------------------------
# upsample_conv_2d WITHOUT folding the FIR (upfirdn_2d_v2.py:65-103 as the reference orders it): the
# stride-2 transposed 3x3 convolution runs at its algorithmic cost (9 tap blocks), the 4x4 FIR is a
# separate bandwidth-bound pass (tbg_fir4) that also applies the layer epilogue.
#   T[2a+py, 2c+px] = sum_{ty,tx} x[a+ty-1, c+tx-1] * Wt[(py,px), (ty,tx)]      a in [0,h], c in [0,w]
# conv2d_transpose runs on the spatially flipped kernel (upfirdn_2d_v2.py:80), T[Y] = sum_i x[i] w[2-(Y-2i)], so
# per axis  phase 0 (even rows): taps t=0 -> w[0] (input a-1), t=1 -> w[2] (input a);
#           phase 1 (odd rows):  tap  t=1 -> w[1] (input a); t=0 is structurally zero (masked).
# Its adjoint (input gradient) is the plain stride-2 VALID 3x3 convolution of the (2h+2) x (2w+2)
# tensor, and its weight gradient is that convolution's weight gradient with the roles of input and
# output gradient exchanged — both run on the power-of-two h x w grid with exactly 9 taps.
# ----------------------------------------------------------------------------------------------
UPT_TAP_MASK = (0b1111, 0b1010, 0b1100, 0b1000)   # phases (py,px) = (0,0), (0,1), (1,0), (1,1); bit = ty*2+tx


def _upT_table() -> torch.Tensor:
    f = torch.zeros(2, 2, 3)
    f[0, 0, 0] = 1.0
    f[0, 1, 2] = 1.0
    f[1, 1, 1] = 1.0
    return f


def _upT_spec(h: int, w: int, I: int, O: int, Ip: int, Op: int, k: int, coef: float, tag: str) -> "WeightSpec":
    assert k == 3
    algo = 9.0 * h * w / (16.0 * (h + 1) * (w + 1))
    g = ConvGeom(h, w, Ip, Op, Axis("up", 2, 1), Axis("up", 2, 1), tag, algo)
    f, ident = _upT_table(), torch.eye(3).flip(0).reshape(1, 3, 3)     # adjoint: stride-2 conv with the flipped kernel
    spec = WeightSpec(g, "upT", "upT", 3, 3, I, O, coef, tables=(f, f.clone(), ident, ident.clone()),
                      adj_shape=(Ip, 9 * Op))
    # kernel launch descriptions of the three convolutions
    spec.fwd_kwargs = dict(Ho=h + 1, Wo=w + 1, taps=(2, 2), pad=(1, 1), stride=(1, 1), up=(1, 1), tap_mask=UPT_TAP_MASK)
    spec.s2_kwargs = dict(Ho=h, Wo=w, taps=(3, 3), pad=(0, 0), stride=(2, 2), up=(0, 0))
    spec.t_hw = (2 * h + 2, 2 * w + 2)
    spec.out_hw = (2 * h, 2 * w)
    return spec
