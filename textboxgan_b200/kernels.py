"""Thin Python callers of the C ABI: torch tensors in (device memory + current stream only),
raw pointers out.  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _lib


# Optional per-launch device timing (bench.py's roofline pass): when PROFILE is a list, every
# tensor-core launch appends (kernel_name, tag, algorithmic_flops, start_event, end_event) recorded on the
# launching stream.
PROFILE: Optional[list] = None
PROFILE_TAG: str = ""


class _Timed:
    def __init__(self, name: str, flops: float):
        self.name, self.flops = name, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, PROFILE_TAG, self.flops, self.e0, self.e1))
        return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _aligned(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Epilogue vectors are read with 128-bit loads; parameters living at odd offsets of a flat
    buffer are copied to an aligned scratch tensor first (a few hundred bytes)."""
    if t is None or t.data_ptr() % 16 == 0:
        return t
    return t.clone()


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
    up: tuple[int, int] | bool = (0, 0),
    col_scale: Optional[torch.Tensor] = None,   # fp32 [B, cout]
    bias: Optional[torch.Tensor] = None,        # fp32 [cout]
    noise: Optional[torch.Tensor] = None,       # fp32 [B, out_H, out_W]
    noise_strength: Optional[torch.Tensor] = None,  # fp32 scalar tensor
    residual: Optional[torch.Tensor] = None,    # bf16, shape of out
    res_scale: float = 1.0,
    res_first: bool = False,
    act: int = 0,
    act_gain: float = 1.0,
    out_fp32: bool = False,
    out: Optional[torch.Tensor] = None,
    tap_mask: Optional[tuple] = None,           # per output phase: bit (th*taps_w+tw) = tap computed (None: all)
    relu_mask: Optional[torch.Tensor] = None,   # bf16, shape of out: result zeroed where relu_mask <= 0
) -> torch.Tensor:
    _require(x, torch.bfloat16, "x")
    _require(w, torch.bfloat16, "w")
    if isinstance(up, bool):
        up = (int(up), int(up))
    B, H, W_, Cin = x.shape
    n_total = w.shape[0]
    cout = n_total // ((1 + up[0]) * (1 + up[1]))
    oH, oW = Ho * (1 + up[0]), Wo * (1 + up[1])
    if w.shape[1] != taps[0] * taps[1] * Cin:
        raise _lib.TbgError(f"w: expected K={taps[0] * taps[1] * Cin}, got {w.shape[1]}")
    if out is None:
        out = torch.empty((B, oH, oW, cout), device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    for t, n in ((col_scale, "col_scale"), (bias, "bias"), (noise, "noise"), (noise_strength, "noise_strength")):
        if t is not None:
            _require(t, torch.float32, n)
    if residual is not None:
        _require(residual, torch.bfloat16, "residual")
    if relu_mask is not None:
        _require(relu_mask, torch.bfloat16, "relu_mask")
    col_scale, bias = _aligned(col_scale), _aligned(bias)
    a = _lib.ConvArgs(
        x=_ptr(x), w=_ptr(w), out=_ptr(out),
        B=B, H=H, W=W_, Cin=Cin, Ho=Ho, Wo=Wo, n_total=n_total, cout=cout,
        taps_h=taps[0], taps_w=taps[1], pad_h=pad[0], pad_w=pad[1],
        stride_h=stride[0], stride_w=stride[1], up_h=up[0], up_w=up[1],
        col_scale=_ptr(col_scale), bias=_ptr(bias), noise=_ptr(noise), noise_strength=_ptr(noise_strength),
        residual=_ptr(residual), res_scale=res_scale, res_first=int(res_first), act=act, act_gain=act_gain, out_fp32=int(out_fp32),
        relu_mask=_ptr(relu_mask),
    )
    if tap_mask is not None:
        for i, m in enumerate(tap_mask):
            a.tap_mask[i] = int(m)
    with _Timed("conv_igemm", 2.0 * B * Ho * Wo * n_total * taps[0] * taps[1] * Cin):
        _lib.check(_lib.load().tbg_conv2d_igemm(C.byref(a), _stream()), "tbg_conv2d_igemm")
    if PROFILE is not None:
        # launch recipe for bench.py's isolated re-timing of this configuration
        PROFILE[-1] = PROFILE[-1] + (dict(
            x_shape=tuple(x.shape), w_shape=tuple(w.shape), Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=up,
            has_scale=col_scale is not None, has_bias=bias is not None, has_noise=noise is not None,
            has_res=residual is not None, res_scale=res_scale, res_first=res_first, act=act, act_gain=act_gain,
            out_fp32=out_fp32, tap_mask=tap_mask),)
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
    up: tuple[int, int] | bool = (0, 0),
    gw: Optional[torch.Tensor] = None,          # fp32 [n_total, taps*Cin], accumulated into
) -> torch.Tensor:
    _require(x, torch.bfloat16, "x")
    _require(gy, torch.bfloat16, "gy")
    if isinstance(up, bool):
        up = (int(up), int(up))
    B, H, W_, Cin = x.shape
    cout = gy.shape[3]
    n_total = cout * (1 + up[0]) * (1 + up[1])
    if gw is None:
        gw = torch.zeros((n_total, taps[0] * taps[1] * Cin), device=x.device, dtype=torch.float32)
    _require(gw, torch.float32, "gw")
    a = _lib.WgradArgs(
        x=_ptr(x), gy=_ptr(gy), gw=_ptr(gw),
        B=B, H=H, W=W_, Cin=Cin, Ho=Ho, Wo=Wo, n_total=n_total, cout=cout,
        taps_h=taps[0], taps_w=taps[1], pad_h=pad[0], pad_w=pad[1],
        stride_h=stride[0], stride_w=stride[1], up_h=up[0], up_w=up[1],
    )
    with _Timed("conv_wgrad", 2.0 * B * Ho * Wo * n_total * taps[0] * taps[1] * Cin):
        _lib.check(_lib.load().tbg_conv2d_wgrad(C.byref(a), _stream()), "tbg_conv2d_wgrad")
    if PROFILE is not None:
        PROFILE[-1] = PROFILE[-1] + (dict(x_shape=tuple(x.shape), w_shape=tuple(gy.shape), stride=stride, up=up),)
    return gw


def upfirdn2d(x: torch.Tensor, k: torch.Tensor, *, upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0,
              pady1=0) -> torch.Tensor:
    """x: [major, inH, inW, minor] fp32 or bf16 (contiguous); k: fp32 [kH, kW]."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise _lib.TbgError(f"upfirdn2d: unsupported dtype {x.dtype}")
    _require(x, x.dtype, "x")
    _require(k, torch.float32, "k")
    major, inH, inW, minor = x.shape
    kH, kW = k.shape
    outW = (inW * upx + padx0 + padx1 - kW + downx) // downx
    outH = (inH * upy + pady0 + pady1 - kH + downy) // downy
    y = torch.empty((major, max(outH, 0), max(outW, 0), minor), device=x.device, dtype=x.dtype)
    st = _lib.load().tbg_upfirdn2d(_ptr(x), _ptr(k), _ptr(y), int(x.dtype == torch.bfloat16), major, inH, inW, minor,
                                   kH, kW, upx, upy, downx, downy, padx0, padx1, pady0, pady1, _stream())
    _lib.check(st, "tbg_upfirdn2d")
    return y


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr_t, beta1: float,
              beta2: float, eps: float) -> None:
    """``lr_t`` is a Python float or a 1-element fp32 device tensor (CUDA-graph friendly)."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _require(t, torch.float32, n)
    lr_dev = None
    if torch.is_tensor(lr_t):
        _require(lr_t, torch.float32, "lr_t")
        lr_dev, lr_t = lr_t, 0.0
    st = _lib.load().tbg_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr_t, _ptr(lr_dev), beta1, beta2,
                                   eps, _stream())
    _lib.check(st, "tbg_adam_step")


def ema_step(dst: torch.Tensor, src: torch.Tensor, beta: float) -> None:
    _require(dst, torch.float32, "dst")
    _require(src, torch.float32, "src")
    st = _lib.load().tbg_ema_step(_ptr(dst), _ptr(src), dst.numel(), beta, _stream())
    _lib.check(st, "tbg_ema_step")


def lstm_seq_fwd(xp: torch.Tensor, w_packed: torch.Tensor):
    """xp f32 [D,B,T,4H]; w_packed bf16 [D,H,H,4] -> (h [D,B,T,H], gates [D,B,T,4H], c [D,B,T,H])."""
    _require(xp, torch.float32, "xp")
    _require(w_packed, torch.bfloat16, "w_packed")
    D, B, T, H4 = xp.shape
    H = H4 // 4
    h = torch.empty((D, B, T, H), device=xp.device, dtype=torch.float32)
    gates = torch.empty((D, B, T, H4), device=xp.device, dtype=torch.float32)
    c = torch.empty((D, B, T, H), device=xp.device, dtype=torch.float32)
    st = _lib.load().tbg_lstm_seq_fwd(_ptr(xp), _ptr(w_packed), _ptr(h), _ptr(gates), _ptr(c), D, B, T, H, _stream())
    _lib.check(st, "tbg_lstm_seq_fwd")
    return h, gates, c


def lstm_seq_bwd(g_h: torch.Tensor, gates: torch.Tensor, c: torch.Tensor, wT_packed: torch.Tensor) -> torch.Tensor:
    _require(g_h, torch.float32, "g_h")
    _require(wT_packed, torch.bfloat16, "wT_packed")
    D, B, T, H = g_h.shape
    g_xp = torch.empty((D, B, T, 4 * H), device=g_h.device, dtype=torch.float32)
    st = _lib.load().tbg_lstm_seq_bwd(_ptr(g_h), _ptr(gates), _ptr(c), _ptr(wT_packed), _ptr(g_xp), D, B, T, H, _stream())
    _lib.check(st, "tbg_lstm_seq_bwd")
    return g_xp


def _bhwc(t: torch.Tensor):
    B, C = t.shape[0], t.shape[-1]
    return B, t.numel() // (B * C), C


def modulate(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """xs = x * s[b, c];  x bf16 [B,...,C], s fp32 [B,C]."""
    _require(x, torch.bfloat16, "x")
    _require(s, torch.float32, "s")
    B, HW, C = _bhwc(x)
    xs = torch.empty_like(x)
    _lib.check(_lib.load().tbg_modulate(_ptr(x), _ptr(s), _ptr(xs), B, HW, C, _stream()), "tbg_modulate")
    return xs


def modulate_bwd(gxs: torch.Tensor, x: torch.Tensor, s: torch.Tensor, gs_init: Optional[torch.Tensor] = None):
    """gx = gxs*s; gs = (gs_init or 0) + sum_hw gxs*x (gs_init is accumulated into, in place)."""
    _require(gxs, torch.bfloat16, "gxs")
    _require(x, torch.bfloat16, "x")
    _require(s, torch.float32, "s")
    B, HW, C = _bhwc(x)
    gx = torch.empty_like(x)
    if gs_init is not None:
        _require(gs_init, torch.float32, "gs_init")
        gs = gs_init
    else:
        gs = torch.zeros_like(s)
    st = _lib.load().tbg_modulate_bwd(_ptr(gxs), _ptr(x), _ptr(s), _ptr(gx), _ptr(gs), B, HW, C, _stream())
    _lib.check(st, "tbg_modulate_bwd")
    return gx, gs


def bias_act_bwd(g_out: torch.Tensor, out: torch.Tensor, *, residual=None, noise=None, d=None, act=True,
                 gain: float = 1.0, want_sums: bool = True, bias_grad_only: bool = False):
    """Returns (gy0 bf16, S1, Spre, Snz) — see include/tbg.h.  ``bias_grad_only``: S1 is the [C]
    bias gradient (summed over the batch too), Spre / Snz are None."""
    _require(g_out, torch.bfloat16, "g_out")
    _require(out, torch.bfloat16, "out")
    B, HW, C = _bhwc(out)
    gy0 = torch.empty_like(out)
    S1 = Spre = Snz = None
    if bias_grad_only:
        S1 = torch.zeros((C,), device=out.device, dtype=torch.float32)
    elif want_sums:
        sums = torch.zeros((3, B, C), device=out.device, dtype=torch.float32)
        S1, Spre, Snz = sums[0], sums[1], sums[2]
    st = _lib.load().tbg_bias_act_bwd(_ptr(g_out), _ptr(out), _ptr(residual), _ptr(noise), _ptr(d), _ptr(gy0),
                                      _ptr(S1), _ptr(Spre), _ptr(Snz),
                                      B, HW, C, int(act), float(gain), int(bias_grad_only), _stream())
    _lib.check(st, "tbg_bias_act_bwd")
    return gy0, S1, Spre, Snz


def bias_act_rgb_bwd(g_out: Optional[torch.Tensor], out: torch.Tensor, g_rgb: torch.Tensor, ws: torch.Tensor, *,
                     noise: Optional[torch.Tensor], d: torch.Tensor, act: int = 1, gain: float = 1.0):
    """bias_act_bwd with the ToRGB gradient formed in the kernel — see include/tbg.h (tbg_bias_act_rgb_bwd).
    Returns (gy0 bf16, S1, Spre, Snz [B,C], gws [B,C,3])."""
    _require(out, torch.bfloat16, "out")
    _require(g_rgb, torch.float32, "g_rgb")
    _require(ws, torch.float32, "ws")
    _require(d, torch.float32, "d")
    if g_out is not None:
        _require(g_out, torch.bfloat16, "g_out")
    B, HW, C_ = _bhwc(out)
    gy0 = torch.empty_like(out)
    sums = torch.zeros((6, B, C_), device=out.device, dtype=torch.float32)       # S1 | Spre | Snz | gws (3 planes)
    gws = sums[3:].view(B, C_, 3)
    st = _lib.load().tbg_bias_act_rgb_bwd(_ptr(g_out), _ptr(out), _ptr(noise), _ptr(d), _ptr(g_rgb), _ptr(ws), _ptr(gy0),
                                          _ptr(sums[0]), _ptr(sums[1]), _ptr(sums[2]), _ptr(gws), B, HW, C_, int(act),
                                          float(gain), _stream())
    _lib.check(st, "tbg_bias_act_rgb_bwd")
    return gy0, sums[0], sums[1], sums[2], gws


def torgb_fwd(x: torch.Tensor, ws: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x bf16 [B,H,W,C], ws fp32 [B,C,3], bias fp32 [3] -> y fp32 [B,H,W,3]."""
    _require(x, torch.bfloat16, "x")
    _require(ws, torch.float32, "ws")
    B, HW, C = _bhwc(x)
    y = torch.empty(x.shape[:-1] + (3,), device=x.device, dtype=torch.float32)
    st = _lib.load().tbg_torgb_fwd(_ptr(x), _ptr(ws), _ptr(bias), _ptr(y), B, HW, C, _stream())
    _lib.check(st, "tbg_torgb_fwd")
    return y


def torgb_bwd(x: torch.Tensor, ws: torch.Tensor, gy: torch.Tensor):
    _require(x, torch.bfloat16, "x")
    _require(ws, torch.float32, "ws")
    _require(gy, torch.float32, "gy")
    B, HW, C = _bhwc(x)
    gx = torch.empty_like(x)
    gws = torch.zeros_like(ws)
    st = _lib.load().tbg_torgb_bwd(_ptr(x), _ptr(ws), _ptr(gy), _ptr(gx), _ptr(gws), B, HW, C, _stream())
    _lib.check(st, "tbg_torgb_bwd")
    return gx, gws


def wprep(w_raw: torch.Tensor, spec, *, want_adj: bool = True, want_q: bool = False, act_dtype=torch.bfloat16):
    """Master weight fp32 [KH,KW,I,O] -> (fwd bf16 [n_total, K], adj bf16 | None, q fp32 [I,O] | None)."""
    _require(w_raw, torch.float32, "w_raw")
    fwd = torch.empty((spec.fwd_rows, spec.fwd_cols), device=w_raw.device, dtype=torch.bfloat16)
    adj = torch.empty((spec.adj_rows, spec.adj_cols), device=w_raw.device, dtype=torch.bfloat16) if want_adj else None
    q = torch.empty((spec.I, spec.O), device=w_raw.device, dtype=torch.float32) if want_q else None
    tables = spec.ctable if want_adj else spec.ctable_noadj
    st = _lib.load().tbg_wprep(_ptr(w_raw), tables, spec.coef, spec.KH, spec.KW, spec.I, spec.O, spec.Ipad, spec.Opad,
                               _ptr(fwd), _ptr(adj), _ptr(q), _stream())
    _lib.check(st, "tbg_wprep")
    return fwd, adj, q


class WPrepPlan:
    """All weight preparations of an iteration as one launch (tbg_wprep_group).  ``entries``: [(w_raw, spec, want_adj,
    want_q)].  Output matrices are allocated once and overwritten by every :meth:`run`; the packed job table lives in
    device memory, so a run is a single kernel launch (CUDA-graph capturable)."""

    def __init__(self, entries):
        h = _lib.load()
        self.entries = list(entries)
        nbytes = h.tbg_wprep_job_bytes()
        host = (C.c_uint8 * (nbytes * len(self.entries)))()
        self.outputs = []
        begin = 0
        for n, (w_raw, spec, want_adj, want_q) in enumerate(self.entries):
            _require(w_raw, torch.float32, "w_raw")
            dev = w_raw.device
            fwd = torch.empty((spec.fwd_rows, spec.fwd_cols), device=dev, dtype=torch.bfloat16)
            adj = torch.empty((spec.adj_rows, spec.adj_cols), device=dev, dtype=torch.bfloat16) if want_adj else None
            q = torch.empty((spec.I, spec.O), device=dev, dtype=torch.float32) if want_q else None
            tables = spec.ctable if want_adj else spec.ctable_noadj
            blocks = h.tbg_wprep_make_job(C.byref(host, n * nbytes), begin, _ptr(w_raw), tables, spec.coef, spec.KH,
                                          spec.KW, spec.I, spec.O, spec.Ipad, spec.Opad, _ptr(fwd), _ptr(adj), _ptr(q))
            if blocks <= 0:
                _lib.check(blocks if blocks < 0 else -1, "tbg_wprep_make_job")
            begin += blocks
            self.outputs.append((fwd, adj, q))
        self.total_blocks = begin
        self.jobs = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(self.entries[0][0].device)

    def run(self):
        st = _lib.load().tbg_wprep_group(_ptr(self.jobs), len(self.entries), self.total_blocks, _stream())
        _lib.check(st, "tbg_wprep_group")
        return self.outputs


def wfold(gfwd: torch.Tensor, spec, *, gq: Optional[torch.Tensor] = None, w_raw: Optional[torch.Tensor] = None,
          out: Optional[torch.Tensor] = None, s: Optional[torch.Tensor] = None,
          t: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Accumulate the master-weight gradient from the fp32 gradient of the fwd GEMM matrix; the
    demodulation term comes from ``gq`` [I,O] or from (``s`` [B,I], ``t`` [B,O]) of demod_bwd."""
    _require(gfwd, torch.float32, "gfwd")
    accumulate = out is not None
    if out is None:
        out = torch.empty((spec.KH, spec.KW, spec.I, spec.O), device=gfwd.device, dtype=torch.float32)
    nb = 0
    if s is not None:
        _require(s, torch.float32, "s")
        _require(t, torch.float32, "t")
        nb = s.shape[0]
    st = _lib.load().tbg_wfold(_ptr(gfwd), _ptr(gq), _ptr(w_raw), spec.ctable, spec.coef, spec.KH, spec.KW, spec.I,
                               spec.O, spec.Ipad, spec.Opad, _ptr(out), _ptr(s), _ptr(t), nb, int(accumulate), _stream())
    _lib.check(st, "tbg_wfold")
    return out


def demod_coef(s: torch.Tensor, q: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """d[b,o] = rsqrt(sum_i s[b,i]^2 q[i,o] + eps)   (modulated_conv2d.py:80-82)."""
    _require(s, torch.float32, "s")
    _require(q, torch.float32, "q")
    B, I = s.shape
    O = q.shape[1]
    d = torch.empty((B, O), device=s.device, dtype=torch.float32)
    _lib.check(_lib.load().tbg_demod_coef(_ptr(s), _ptr(q), _ptr(d), B, I, O, float(eps), _stream()), "tbg_demod_coef")
    return d


def demod_bwd(S1, Spre, Snz, d, ns, bias, s, q):
    """-> (t [B,O], gbias [O], gns [1], gs_demod [B,I]) — see include/tbg.h."""
    for tn, n in ((S1, "S1"), (Spre, "Spre"), (d, "d"), (s, "s"), (q, "q")):
        _require(tn, torch.float32, n)
    B, I = s.shape
    O = q.shape[1]
    dev = s.device
    t = torch.empty((B, O), device=dev, dtype=torch.float32)
    gbias = torch.empty((O,), device=dev, dtype=torch.float32)
    gns = torch.empty((1,), device=dev, dtype=torch.float32)
    gs = torch.empty((B, I), device=dev, dtype=torch.float32)
    st = _lib.load().tbg_demod_bwd(_ptr(S1), _ptr(Spre), _ptr(Snz), _ptr(d), _ptr(ns), _ptr(bias), _ptr(s), _ptr(q),
                                   _ptr(t), _ptr(gbias), _ptr(gns), _ptr(gs), B, I, O, _stream())
    _lib.check(st, "tbg_demod_bwd")
    return t, gbias, gns, gs


def style_dense_fwd(style: torch.Tensor, ws, bs, idxs, coef: float):
    """style fp32 [B,n,S]; per layer w [S,I_l], b [I_l], style row idx_l -> list of s_l [B,I_l]
    (= coef * style[:,idx_l] @ w_l + b_l + 1), one launch."""
    _require(style, torch.float32, "style")
    B, n, S = style.shape
    outs = [torch.empty((B, w.shape[1]), device=style.device, dtype=torch.float32) for w in ws]
    arr = (_lib.StyleLayer * len(ws))()
    for a, w, b, o, i in zip(arr, ws, bs, outs, idxs):
        _require(w, torch.float32, "w")
        _require(b, torch.float32, "b")
        a.w, a.b, a.s, a.I, a.idx = _ptr(w), _ptr(b), _ptr(o), w.shape[1], int(i)
    st = _lib.load().tbg_style_dense_fwd(arr, len(ws), _ptr(style), B, n, S, float(coef), _stream())
    _lib.check(st, "tbg_style_dense_fwd")
    return outs


def style_dense_bwd(style: torch.Tensor, ws, gss, idxs, coef: float):
    """-> (gstyle [B,n,S], [gw_l [S,I_l]], [gb_l [I_l]]), two launches."""
    _require(style, torch.float32, "style")
    B, n, S = style.shape
    dev = style.device
    gstyle = torch.empty_like(style)
    gws = [torch.empty_like(w) for w in ws]
    gbs = [torch.empty((w.shape[1],), device=dev, dtype=torch.float32) for w in ws]
    arr = (_lib.StyleLayer * len(ws))()
    dummy = gbs[0]
    for a, w, g, gw, gb, i in zip(arr, ws, gss, gws, gbs, idxs):
        _require(g, torch.float32, "gs")
        a.w, a.b, a.gs, a.gw, a.gb, a.I, a.idx = _ptr(w), _ptr(dummy), _ptr(g), _ptr(gw), _ptr(gb), w.shape[1], int(i)
    st = _lib.load().tbg_style_dense_bwd(arr, len(ws), _ptr(style), _ptr(gstyle), B, n, S, float(coef), _stream())
    _lib.check(st, "tbg_style_dense_bwd")
    return gstyle, gws, gbs


def crop_resize_fwd(img: torch.Tensor, labels: torch.Tensor, blank: int, char_width, out_hw) -> torch.Tensor:
    """convert_inputs forward — see include/tbg.h (tbg_crop_resize_fwd)."""
    from fractions import Fraction

    _require(img, torch.float32, "img")
    _require(labels, torch.int32, "labels")
    B, _, H, W_ = img.shape
    cw = Fraction(char_width)
    out = torch.empty((B, out_hw[0], out_hw[1], 3), device=img.device, dtype=torch.float32)
    st = _lib.load().tbg_crop_resize_fwd(_ptr(img), _ptr(labels), _ptr(out), B, H, W_, out_hw[0], out_hw[1], labels.shape[1],
                                         int(blank), cw.numerator, cw.denominator, _stream())
    _lib.check(st, "tbg_crop_resize_fwd")
    return out


def crop_resize_bwd(g: torch.Tensor, labels: torch.Tensor, blank: int, char_width, img_hw) -> torch.Tensor:
    from fractions import Fraction

    _require(g, torch.float32, "g")
    _require(labels, torch.int32, "labels")
    B, oh, ow, _ = g.shape
    cw = Fraction(char_width)
    gimg = torch.zeros((B, 3, img_hw[0], img_hw[1]), device=g.device, dtype=torch.float32)
    st = _lib.load().tbg_crop_resize_bwd(_ptr(g), _ptr(labels), _ptr(gimg), B, img_hw[0], img_hw[1], oh, ow, labels.shape[1],
                                         int(blank), cw.numerator, cw.denominator, _stream())
    _lib.check(st, "tbg_crop_resize_bwd")
    return gimg


def fromrgb_fwd(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, coef: float, gain: float) -> torch.Tensor:
    """img fp32 NCHW [B,3,H,W], w fp32 [3,C], bias [C] -> lrelu(coef*img.w + bias)*gain, bf16 NHWC [B,H,W,C]."""
    _require(img, torch.float32, "img")
    _require(w, torch.float32, "w")
    _require(bias, torch.float32, "bias")
    B, _, H, W_ = img.shape
    C_ = w.shape[1]
    out = torch.empty((B, H, W_, C_), device=img.device, dtype=torch.bfloat16)
    st = _lib.load().tbg_fromrgb_fwd(_ptr(img), _ptr(w), _ptr(bias), _ptr(out), B, H * W_, C_, float(coef), float(gain),
                                     _stream())
    _lib.check(st, "tbg_fromrgb_fwd")
    return out


def fromrgb_bwd(img: torch.Tensor, w: torch.Tensor, g_out: torch.Tensor, out: torch.Tensor, coef: float, gain: float,
                *, want_img: bool = True, want_w: bool = True):
    """-> (gimg fp32 NCHW | None, gw fp32 [3,C] | None, gb fp32 [C] | None)."""
    _require(img, torch.float32, "img")
    _require(g_out, torch.bfloat16, "g_out")
    _require(out, torch.bfloat16, "out")
    B, _, H, W_ = img.shape
    C_ = w.shape[1]
    gimg = torch.empty_like(img) if want_img else None
    gwb = torch.zeros((4, C_), device=img.device, dtype=torch.float32) if want_w else None
    st = _lib.load().tbg_fromrgb_bwd(_ptr(img), _ptr(w), _ptr(g_out), _ptr(out), _ptr(gimg),
                                     _ptr(gwb[:3]) if want_w else None, _ptr(gwb[3]) if want_w else None, B, H * W_, C_,
                                     float(coef), float(gain), _stream())
    _lib.check(st, "tbg_fromrgb_bwd")
    return gimg, (gwb[:3] if want_w else None), (gwb[3] if want_w else None)


def fir4(x: torch.Tensor, out_hw: tuple[int, int], off: tuple[int, int], scale: float, *, d=None, noise=None,
         noise_strength=None, bias=None, act: int = 0, gain: float = 1.0) -> torch.Tensor:
    """4x4 FIR [1,3,3,1]^2 on NHWC bf16 with the optional layer epilogue — see include/tbg.h (tbg_fir4)."""
    _require(x, torch.bfloat16, "x")
    B, IH, IW, C_ = x.shape
    out = torch.empty((B, out_hw[0], out_hw[1], C_), device=x.device, dtype=torch.bfloat16)
    for t, n in ((d, "d"), (noise, "noise"), (noise_strength, "noise_strength"), (bias, "bias")):
        if t is not None:
            _require(t, torch.float32, n)
    st = _lib.load().tbg_fir4(_ptr(x), _ptr(out), B, IH, IW, out_hw[0], out_hw[1], C_, int(off[0]), int(off[1]),
                              float(scale), _ptr(d), _ptr(noise), _ptr(noise_strength), _ptr(bias), int(act),
                              float(gain), _stream())
    _lib.check(st, "tbg_fir4")
    return out


def fir4_down(x: torch.Tensor, out_hw: tuple[int, int], sy: int, off: tuple[int, int], scale: float) -> torch.Tensor:
    """The 4x4 FIR at the pixels a stride-(sy, 2) 1x1 convolution reads — see include/tbg.h (tbg_fir4_down)."""
    _require(x, torch.bfloat16, "x")
    B, IH, IW, C_ = x.shape
    out = torch.empty((B, out_hw[0], out_hw[1], C_), device=x.device, dtype=torch.bfloat16)
    st = _lib.load().tbg_fir4_down(_ptr(x), _ptr(out), B, IH, IW, out_hw[0], out_hw[1], C_, int(sy), int(off[0]),
                                   int(off[1]), float(scale), _stream())
    _lib.check(st, "tbg_fir4_down")
    return out


def fir4_down_adjoint(g: torch.Tensor, in_hw: tuple[int, int], sy: int, off: tuple[int, int], scale: float,
                      add: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Transpose of :func:`fir4_down` (+ ``add``) — see include/tbg.h (tbg_fir4_down_adjoint)."""
    _require(g, torch.bfloat16, "g")
    B, OH, OW, C_ = g.shape
    if out is None:
        out = torch.empty((B, in_hw[0], in_hw[1], C_), device=g.device, dtype=torch.bfloat16)
    else:
        _require(out, torch.bfloat16, "out")
    if add is not None:
        _require(add, torch.bfloat16, "add")
        if tuple(add.shape) != tuple(out.shape):
            raise _lib.TbgError(f"add: expected shape {tuple(out.shape)}, got {tuple(add.shape)}")
    st = _lib.load().tbg_fir4_down_adjoint(_ptr(g), _ptr(add), _ptr(out), B, in_hw[0], in_hw[1], OH, OW, C_, int(sy),
                                           int(off[0]), int(off[1]), float(scale), _stream())
    _lib.check(st, "tbg_fir4_down_adjoint")
    return out


def wfold_adj(gadj: torch.Tensor, spec, *, w_raw: Optional[torch.Tensor] = None, s: Optional[torch.Tensor] = None,
              t: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, flip: bool = False) -> torch.Tensor:
    """Master-weight gradient from a gradient in the adjoint-matrix layout [Ipad, taps*Opad] (identity tables,
    or the spatially flipped kernel when ``flip``)."""
    _require(gadj, torch.float32, "gadj")
    accumulate = out is not None
    if out is None:
        out = torch.empty((spec.KH, spec.KW, spec.I, spec.O), device=gadj.device, dtype=torch.float32)
    nb = s.shape[0] if s is not None else 0
    st = _lib.load().tbg_wfold_adj(_ptr(gadj), _ptr(w_raw), spec.coef, spec.KH, spec.KW, spec.I, spec.O, spec.Opad,
                                   _ptr(out), _ptr(s), _ptr(t), nb, int(flip), int(accumulate), _stream())
    _lib.check(st, "tbg_wfold_adj")
    return out


def _dec_struct(w: dict):
    return _lib.DecWeights(**{k: _ptr(v) for k, v in w.items()})


def attn_decoder_fwd(mem: torch.Tensor, keys: torch.Tensor, w: dict, steps: int):
    """mem f32 [B,T,512], keys f32 [B,T,256], w: dict of packed decoder weights -> (logits, saved)."""
    _require(mem, torch.float32, "mem")
    _require(keys, torch.float32, "keys")
    B, T, _ = mem.shape
    dev = mem.device
    logits = torch.empty((B, steps, 96), device=dev)
    sv = dict(a=torch.empty((B, steps, T), device=dev), ctx=torch.empty((B, steps, 512), device=dev),
              gates=torch.empty((B, steps, 1024), device=dev), c=torch.empty((B, steps, 256), device=dev),
              h=torch.empty((B, steps, 256), device=dev), prev=torch.empty((B, steps), device=dev, dtype=torch.int32))
    ws = _dec_struct(w)
    st = _lib.load().tbg_attn_decoder_fwd(_ptr(mem), _ptr(keys), C.byref(ws), _ptr(logits), _ptr(sv["a"]),
                                          _ptr(sv["ctx"]), _ptr(sv["gates"]), _ptr(sv["c"]), _ptr(sv["h"]),
                                          _ptr(sv["prev"]), B, T, steps, _stream())
    _lib.check(st, "tbg_attn_decoder_fwd")
    return logits, sv


def attn_decoder_bwd(mem: torch.Tensor, keys: torch.Tensor, w: dict, g_logits: torch.Tensor, sv: dict):
    _require(g_logits, torch.float32, "g_logits")
    B, T, _ = mem.shape
    steps = g_logits.shape[1]
    g_mem = torch.empty_like(mem)
    g_keys = torch.empty_like(keys)
    ws = _dec_struct(w)
    st = _lib.load().tbg_attn_decoder_bwd(_ptr(mem), _ptr(keys), C.byref(ws), _ptr(g_logits), _ptr(sv["a"]),
                                          _ptr(sv["ctx"]), _ptr(sv["gates"]), _ptr(sv["c"]), _ptr(sv["h"]),
                                          _ptr(g_mem), _ptr(g_keys), B, T, steps, _stream())
    _lib.check(st, "tbg_attn_decoder_bwd")
    return g_mem, g_keys


# ----------------------------------------------------------------------------------------------
# small fp32 dense layers, word encoder, batch statistics, RGB branch (csrc/dense.cu, stats.cu, rgb.cu)
# ----------------------------------------------------------------------------------------------
def dense_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, coef: float = 1.0, bias_coef: float = 1.0,
              act: int = 0, gain: float = 1.0) -> torch.Tensor:
    """y = act((x @ w) * coef + bias * bias_coef) * gain — x fp32 [M,K], w fp32 [K,N] (tbg_dense_fwd)."""
    _require(x, torch.float32, "x")
    _require(w, torch.float32, "w")
    if bias is not None:
        _require(bias, torch.float32, "bias")
    M, Kd = x.shape
    N = w.shape[1]
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    st = _lib.load().tbg_dense_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), M, Kd, N, float(coef), float(bias_coef), int(act),
                                   float(gain), _stream())
    _lib.check(st, "tbg_dense_fwd")
    return y


def dense_bwd(x: torch.Tensor, w: torch.Tensor, y: Optional[torch.Tensor], gy: torch.Tensor, *, coef: float = 1.0,
              bias_coef: float = 1.0, act: int = 0, gain: float = 1.0, want_gx: bool = True, want_gw: bool = True,
              want_gb: bool = True):
    """-> (gx [M,K] | None, gw [K,N] | None, gb [N] | None) — see include/tbg.h (tbg_dense_bwd)."""
    _require(gy, torch.float32, "gy")
    M, Kd = x.shape
    N = w.shape[1]
    dev = x.device
    gpre = torch.empty((M, N), device=dev, dtype=torch.float32)
    gx = torch.empty((M, Kd), device=dev, dtype=torch.float32) if want_gx else None
    gw = torch.empty((Kd, N), device=dev, dtype=torch.float32) if want_gw else None
    gb = torch.empty((N,), device=dev, dtype=torch.float32) if want_gb else None
    st = _lib.load().tbg_dense_bwd(_ptr(x), _ptr(w), _ptr(y), _ptr(gy), _ptr(gpre), _ptr(gx), _ptr(gw), _ptr(gb), M, Kd, N,
                                   float(coef), float(bias_coef), int(act), float(gain), 0, _stream())
    _lib.check(st, "tbg_dense_bwd")
    return gx, gw, gb


def pixel_norm_fwd(x: torch.Tensor) -> torch.Tensor:
    _require(x, torch.float32, "x")
    y = torch.empty_like(x)
    _lib.check(_lib.load().tbg_pixel_norm_fwd(_ptr(x), _ptr(y), x.shape[0], x.shape[1], _stream()), "tbg_pixel_norm_fwd")
    return y


def pixel_norm_bwd(x: torch.Tensor, gy: torch.Tensor) -> torch.Tensor:
    _require(x, torch.float32, "x")
    _require(gy, torch.float32, "gy")
    gx = torch.empty_like(x)
    st = _lib.load().tbg_pixel_norm_bwd(_ptr(x), _ptr(gy), _ptr(gx), x.shape[0], x.shape[1], _stream())
    _lib.check(st, "tbg_pixel_norm_bwd")
    return gx


def word_encoder_fwd(words: torch.Tensor, w0: torch.Tensor, table: torch.Tensor, mask: Optional[torch.Tensor], keep: float,
                     fc_w: torch.Tensor, fc_b: torch.Tensor, out_hwc: tuple, act_dtype=torch.bfloat16):
    """-> (out NHWC bf16 [B,h,w,c], emb fp32 [B*mcn,E], act fp32 [B*mcn,D]) — tbg_word_encoder_fwd."""
    _require(words, torch.int32, "words")
    for t, n in ((w0, "w0"), (table, "table"), (fc_w, "fc_w"), (fc_b, "fc_b")):
        _require(t, torch.float32, n)
    if mask is not None:
        _require(mask, torch.float32, "mask")
    B, mcn = words.shape
    E, D = fc_w.shape
    oh, ow, oc = out_hwc
    dev = words.device
    out = torch.empty((B, oh, ow, oc), device=dev, dtype=torch.bfloat16)
    emb = torch.empty((B * mcn, E), device=dev, dtype=torch.float32)
    act = torch.empty((B * mcn, D), device=dev, dtype=torch.float32)
    st = _lib.load().tbg_word_encoder_fwd(_ptr(words), _ptr(w0), _ptr(table), _ptr(mask), float(keep), _ptr(fc_w),
                                          _ptr(fc_b), _ptr(emb), _ptr(act), _ptr(out), B, mcn, E, D, oh, ow, oc, _stream())
    _lib.check(st, "tbg_word_encoder_fwd")
    return out, emb, act


def word_encoder_bwd(words, mask, keep: float, fc_w, emb, act, g_out, table_rows: int, out_hwc: tuple):
    """-> (g_table [V-1,E], g_fc_w [E,D], g_fc_b [D]) — tbg_word_encoder_bwd."""
    _require(g_out, torch.bfloat16, "g_out")
    B, mcn = words.shape
    E, D = fc_w.shape
    oh, ow, oc = out_hwc
    dev = words.device
    gpre = torch.empty((B * mcn, D), device=dev, dtype=torch.float32)
    g_table = torch.zeros((table_rows, E), device=dev, dtype=torch.float32)
    g_fc_w = torch.empty((E, D), device=dev, dtype=torch.float32)
    g_fc_b = torch.empty((D,), device=dev, dtype=torch.float32)
    st = _lib.load().tbg_word_encoder_bwd(_ptr(words), _ptr(mask), float(keep), _ptr(fc_w), _ptr(emb), _ptr(act),
                                          _ptr(g_out), _ptr(gpre), _ptr(g_table), _ptr(g_fc_w), _ptr(g_fc_b), B, mcn, E, D,
                                          oh, ow, oc, _stream())
    _lib.check(st, "tbg_word_encoder_bwd")
    return g_table, g_fc_w, g_fc_b


def minibatch_std_fwd(x: torch.Tensor, n_calls: int, cpad: int, group_size: int = 4):
    """x bf16 [n_calls*B, H, W, C] -> (xcat bf16 [n_calls*B, H, W, cpad] = [x | statistic | 0], stat fp32 [n_calls*B])."""
    _require(x, torch.bfloat16, "x")
    Bt, H, W_, C_ = x.shape
    xcat = torch.empty((Bt, H, W_, cpad), device=x.device, dtype=torch.bfloat16)
    stat = torch.empty((Bt,), device=x.device, dtype=torch.float32)
    st = _lib.load().tbg_minibatch_std_fwd(_ptr(x), _ptr(xcat), _ptr(stat), Bt // n_calls, n_calls, group_size, H * W_, C_,
                                           cpad, _stream())
    _lib.check(st, "tbg_minibatch_std_fwd")
    return xcat, stat


def minibatch_std_bwd(x: torch.Tensor, gxcat: torch.Tensor, n_calls: int, group_size: int = 4) -> torch.Tensor:
    _require(x, torch.bfloat16, "x")
    _require(gxcat, torch.bfloat16, "gxcat")
    Bt, H, W_, C_ = x.shape
    gx = torch.empty_like(x)
    st = _lib.load().tbg_minibatch_std_bwd(_ptr(x), _ptr(gxcat), _ptr(gx), Bt // n_calls, n_calls, group_size, H * W_, C_,
                                           gxcat.shape[3], _stream())
    _lib.check(st, "tbg_minibatch_std_bwd")
    return gx


def r1_sqnorm(g: torch.Tensor) -> torch.Tensor:
    """out[b] = sum of g[b]^2 over everything but the batch axis (fp32)."""
    _require(g, torch.float32, "g")
    B = g.shape[0]
    out = torch.empty((B,), device=g.device, dtype=torch.float32)
    _lib.check(_lib.load().tbg_r1_sqnorm(_ptr(g), _ptr(out), B, g.numel() // B, _stream()), "tbg_r1_sqnorm")
    return out


def r1_sqnorm_bwd(g: torch.Tensor, gout: torch.Tensor) -> torch.Tensor:
    _require(g, torch.float32, "g")
    _require(gout, torch.float32, "gout")
    gg = torch.empty_like(g)
    B = g.shape[0]
    _lib.check(_lib.load().tbg_r1_sqnorm_bwd(_ptr(g), _ptr(gout), _ptr(gg), B, g.numel() // B, _stream()), "tbg_r1_sqnorm_bwd")
    return gg


def torgb_skip_fwd(x: torch.Tensor, ws: torch.Tensor, bias: Optional[torch.Tensor], y_prev: Optional[torch.Tensor],
                   words: Optional[torch.Tensor], nchw: bool) -> torch.Tensor:
    """ToRGB + upsampled skip (+ mask_text_box + NCHW layout) — see include/tbg.h (tbg_torgb_skip_fwd)."""
    _require(x, torch.bfloat16, "x")
    _require(ws, torch.float32, "ws")
    B, H, W_, C_ = x.shape
    if y_prev is not None:
        _require(y_prev, torch.float32, "y_prev")
    mcn = 0
    if words is not None:
        _require(words, torch.int32, "words")
        mcn = words.shape[1]
    out = torch.empty((B, 3, H, W_) if nchw else (B, H, W_, 3), device=x.device, dtype=torch.float32)
    st = _lib.load().tbg_torgb_skip_fwd(_ptr(x), _ptr(ws), _ptr(bias), _ptr(y_prev), _ptr(words), _ptr(out), B, H, W_, C_,
                                        mcn, int(nchw), _stream())
    _lib.check(st, "tbg_torgb_skip_fwd")
    return out


def image_grad_nhwc(g: torch.Tensor, words: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 NCHW [B,3,H,W] -> masked NHWC [B,H,W,3] (adjoint of the mask + layout change of torgb_skip_fwd)."""
    _require(g, torch.float32, "g")
    B, _, H, W_ = g.shape
    mcn = words.shape[1] if words is not None else 0
    out = torch.empty((B, H, W_, 3), device=g.device, dtype=torch.float32)
    _lib.check(_lib.load().tbg_image_grad_nhwc(_ptr(g), _ptr(words), _ptr(out), B, H, W_, mcn, _stream()), "tbg_image_grad_nhwc")
    return out


def bias_act_fwd(t: torch.Tensor, *, noise: Optional[torch.Tensor] = None, noise_strength: Optional[torch.Tensor] = None,
                 bias: Optional[torch.Tensor] = None, act: int = 1, gain: float = 1.0) -> torch.Tensor:
    """out = act(t + noise*noise_strength + bias) * gain — t bf16 [B,...,C] (tbg_bias_act_fwd)."""
    _require(t, torch.bfloat16, "t")
    B, HW, C_ = _bhwc(t)
    out = torch.empty_like(t)
    bias = _aligned(bias)
    st = _lib.load().tbg_bias_act_fwd(_ptr(t), _ptr(noise), _ptr(noise_strength), _ptr(bias), _ptr(out), B, HW, C_, int(act),
                                      float(gain), _stream())
    _lib.check(st, "tbg_bias_act_fwd")
    return out


def rowdot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """out[b,c] = sum over pixels of a*b — a, b bf16 [B,...,C] -> fp32 [B,C] (tbg_rowdot)."""
    _require(a, torch.bfloat16, "a")
    _require(b, torch.bfloat16, "b")
    B, HW, C_ = _bhwc(a)
    out = torch.zeros((B, C_), device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().tbg_rowdot(_ptr(a), _ptr(b), _ptr(out), B, HW, C_, _stream()), "tbg_rowdot")
    return out


def batch_resize_normalize(packed: torch.Tensor, offsets: torch.Tensor, src_h: torch.Tensor, src_w: torch.Tensor,
                           dst_w: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Loader transform of a batch on the device — see include/tbg.h (tbg_batch_resize_normalize).  ``packed`` uint8 device
    buffer holding every HWC BGR image back to back; offsets int64 [B]; src_h / src_w / dst_w int32 [B]."""
    _require(packed, torch.uint8, "packed")
    _require(offsets, torch.int64, "offsets")
    for t, n in ((src_h, "src_h"), (src_w, "src_w"), (dst_w, "dst_w")):
        _require(t, torch.int32, n)
    B = offsets.shape[0]
    out = torch.empty((B, 3, H, W), device=packed.device, dtype=torch.float32)
    st = _lib.load().tbg_batch_resize_normalize(_ptr(packed), _ptr(offsets), _ptr(src_h), _ptr(src_w), _ptr(dst_w), _ptr(out),
                                                B, int(H), int(W), _stream())
    _lib.check(st, "tbg_batch_resize_normalize")
    return out
