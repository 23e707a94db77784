"""First-order fused layer Functions of the plain training step.

Each Function is one reference layer group executed as a handful of kernels, with a hand-written
backward (no autograd tape through the element-wise terms):

* :class:`ModConvAct` — ModulatedConv2D + Noise + BiasAct (modulated_conv2d.py:66-122, noise.py:12-22,
  bias_act.py:25-34): ``modulate`` -> tcgen05 conv with the demod/noise/bias/lrelu epilogue; backward
  = ``bias_act_bwd`` -> input-gradient conv + weight-gradient conv -> ``modulate_bwd``.
* :class:`ConvAct` — Conv2D + BiasAct (+ residual merge) of the discriminator (conv.py:51-73,
  discriminator.py:68-84).
* :class:`ToRGB` — the N = 3 modulated 1x1 convolution (to_rgb.py:28-33).

They are only differentiable once.  The path-length and R1 regularisers (double backward,
training_step.py:323-333, 363-368) run the same layers through the composable primitives of
:mod:`textboxgan_b200.conv` inside ``layers.double_backward()``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .conv import ConvGeom, relayout_for_adjoint


# Prepared GEMM matrices of a master weight, shared by every use of that weight inside ONE training
# step (the discriminator runs twice per step, on fake and on real images, training_step.py:260,288;
# weights only change in the optimiser updates at the end of the step).
_STEP_CACHE: dict = {}
# Requests of the running training iteration: key -> [w_raw, spec, want_adj, want_q] (None outside begin_step/end_step).
_STEP_LOG: Optional[dict] = None
GROUP_WPREP = True


def clear_step_cache() -> None:
    _STEP_CACHE.clear()


class StepWeights:
    """Weight preparation of a training iteration as ONE grouped launch (K.WPrepPlan) instead of one latency-bound launch
    per weight.  The first eager iteration records which (weight, geometry) pairs the step asks for; :meth:`end_step`
    builds the plan from that record (never while a CUDA graph is being captured), and from then on :meth:`begin_step`
    fills the step cache with one launch.  A request that is not in the plan (weights re-bound to new storage, a step
    variant with more layers) falls back to its own launch and makes the next eager iteration rebuild the plan."""

    def __init__(self):
        self.plan = None
        self.keys: dict = {}
        self.retired: list = []       # superseded plans: CUDA graphs captured earlier still replay their launch and buffers

    def begin_step(self) -> None:
        global _STEP_LOG
        _STEP_CACHE.clear()
        _STEP_LOG = {}
        if self.plan is not None and GROUP_WPREP:
            outs = self.plan.run()
            for key, out in zip(self.keys, outs):
                _STEP_CACHE[key] = out

    def end_step(self) -> None:
        global _STEP_LOG
        log, _STEP_LOG = _STEP_LOG, None
        _STEP_CACHE.clear()
        if not GROUP_WPREP or not log:
            return
        covered = all(k in self.keys and self.keys[k][0] >= v[2] and self.keys[k][1] >= v[3] for k, v in log.items())
        if covered or (log[next(iter(log))][0].is_cuda and torch.cuda.is_current_stream_capturing()):
            return
        if self.plan is not None:
            self.retired.append(self.plan)
            # keep what the old plan covered and this iteration did not ask for (another step variant's layers)
            for k, ent in zip(self.keys, self.plan.entries):
                if k not in log:
                    log[k] = list(ent)
        self.keys = {k: (v[2], v[3]) for k, v in log.items()}
        self.plan = K.WPrepPlan([tuple(v) for v in log.values()])


def _prepared(w_raw, spec, want_adj: bool, want_q: bool):
    key = (w_raw.data_ptr(), id(spec))
    if _STEP_LOG is not None:
        rec = _STEP_LOG.get(key)
        if rec is None:
            _STEP_LOG[key] = [w_raw, spec, bool(want_adj), bool(want_q)]
        else:
            rec[2], rec[3] = rec[2] or bool(want_adj), rec[3] or bool(want_q)
    hit = _STEP_CACHE.get(key)
    if hit is not None and (hit[1] is not None or not want_adj) and (hit[2] is not None or not want_q):
        return hit
    out = K.wprep(w_raw, spec, want_adj=want_adj, want_q=want_q)
    _STEP_CACHE[key] = out
    return out


def _aligned_vec(t):
    return t if t.is_contiguous() else t.contiguous()


# Tags ("dconv", ...) whose weight gradients are not wanted by the running backward pass: the
# generator pass of training_step.py:194-199 differentiates through the discriminator but only asks
# for generator variables, which a static ``needs_input_grad`` cannot express.
_SKIP_WGRAD_TAGS: frozenset = frozenset()


class skip_weight_grads:
    def __init__(self, *tags: str):
        self.tags = frozenset(tags)

    def __enter__(self):
        global _SKIP_WGRAD_TAGS
        self.prev = _SKIP_WGRAD_TAGS
        _SKIP_WGRAD_TAGS = self.tags
        return self

    def __exit__(self, *exc):
        global _SKIP_WGRAD_TAGS
        _SKIP_WGRAD_TAGS = self.prev
        return False


# Leading-batch limit per tag for the running backward pass: when the discriminator evaluated fake and
# real images in one concatenated pass, the generator pass only carries a cotangent on the fake half, so
# its input-gradient kernels run on that half and the rest of the gradient is zero by construction.
_BATCH_LIMIT: dict = {}


class backward_batch_limit:
    def __init__(self, tag: str, limit: int):
        self.tag, self.limit = tag, int(limit)

    def __enter__(self):
        self.prev = _BATCH_LIMIT.get(self.tag)
        _BATCH_LIMIT[self.tag] = self.limit
        return self

    def __exit__(self, *exc):
        if self.prev is None:
            _BATCH_LIMIT.pop(self.tag, None)
        else:
            _BATCH_LIMIT[self.tag] = self.prev
        return False


def _limited_grad(shape, like: torch.Tensor) -> torch.Tensor:
    """Gradient buffer of a batch-limited backward kernel: only the first ``lim`` samples are written AND only those are
    ever read — every consumer on the generator pass (ConvAct, SkipSplit, FromRGB, and the per-sample / per-call ops
    between them) slices or ignores the rest, and the concatenation's backward drops it — so the tail is left
    uninitialised instead of being zero-filled (25 memsets, 0.4 ms per iteration at configs[2])."""
    return torch.empty(shape, device=like.device, dtype=like.dtype)


# upsample_conv_2d forward: FIR folded into a 4-phase 3x3 convolution (True) or transposed conv + FIR pass (False);
# the backward pass is the unfolded one either way (FIR adjoint + stride-2 3x3 convolution at algorithmic cost).
FOLD_UP_FORWARD = True


def _act_dtype():
    from . import layers as L

    return L.ACT_DTYPE


class ModConvAct(torch.autograd.Function):
    """x, style scale s, MASTER weight -> lrelu(demod(conv(x*s, w)) + noise*ns + bias)*gain."""

    @staticmethod
    def forward(ctx, x, s, w_raw, noise, ns, bias, spec, gain: float):
        geom = spec.geom
        x = x.contiguous()
        s = s.contiguous()
        wmat, wadj, q = _prepared(w_raw, spec, True, True)                       # one launch per step
        d = K.demod_coef(s, q)                                                   # modulated_conv2d.py:80-82
        xs = K.modulate(x, s)
        K.PROFILE_TAG = (geom.tag, geom.algo_frac)
        out = K.conv2d_igemm(xs, wmat, **geom.kernel_kwargs(), col_scale=d, noise=noise.contiguous(),
                             noise_strength=ns.reshape(1), bias=bias, act=1, act_gain=gain)
        ctx.save_for_backward(x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias)
        ctx.spec, ctx.gain = spec, gain
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias = ctx.saved_tensors
        spec = ctx.spec
        g = spec.geom
        gy0, S1, Spre, Snz = K.bias_act_bwd(g_out.contiguous(), out, noise=noise.contiguous(), d=d, act=True,
                                            gain=ctx.gain)
        # dL/d(d) -> t = dL/d(s^2 @ q), the bias / noise-strength gradients and the demodulation term
        # of dL/ds in one launch (d = rsqrt(s^2 @ q + eps))
        t, gbias, gns, gs = K.demod_bwd(S1, Spre, Snz, d, ns.reshape(1), _aligned_vec(bias), s, q)
        K.PROFILE_TAG = (g.tag, g.algo_frac)
        gxs = K.conv2d_igemm(gy0, wadj, **g.adjoint().kernel_kwargs())
        gwmat = K.conv2d_wgrad(xs, gy0, **g.kernel_kwargs())
        gx, gs = K.modulate_bwd(gxs, x, s, gs_init=gs)
        gw_raw = K.wfold(gwmat, spec, w_raw=w_raw, s=s, t=t)                   # + dL/dq = (s^2)^T t
        return gx, gs, gw_raw, None, gns.reshape(ns.shape), gbias, None, None


class ModConvActRGB(torch.autograd.Function):
    """ModConvAct (second convolution of a synthesis block) together with the ToRGB that reads its output and the skip sum
    (fused.ToRGBSkip) — synthesis_block.py:143-152.  One Function so that the backward pass never materialises the
    C-channel ToRGB input gradient: ``tbg_bias_act_rgb_bwd`` forms g_out + g_rgb (x) ws on the fly, applies the activation
    gradient and also accumulates the ToRGB weight gradient in the same pass over ``out`` (previously torgb_bwd wrote
    that gradient, autograd added it to the next block's, and bias_act_bwd re-read the sum).
    Returns (out bf16 [B,H,W,O], y: the RGB skip sum [B,H,W,3] fp32, or the masked NCHW image on the last block)."""

    @staticmethod
    def forward(ctx, x, s, w_raw, noise, ns, bias, spec, gain: float, ws_rgb, bias_rgb, y_prev, words, nchw: bool):
        geom = spec.geom
        x = x.contiguous()
        s = s.contiguous()
        wmat, wadj, q = _prepared(w_raw, spec, True, True)
        d = K.demod_coef(s, q)
        xs = K.modulate(x, s)
        K.PROFILE_TAG = (geom.tag, geom.algo_frac)
        out = K.conv2d_igemm(xs, wmat, **geom.kernel_kwargs(), col_scale=d, noise=noise.contiguous(),
                             noise_strength=ns.reshape(1), bias=bias, act=1, act_gain=gain)
        ws_rgb = ws_rgb.contiguous()
        if words is not None:
            words = words.to(torch.int32).contiguous()
        y_prev = y_prev.contiguous() if y_prev is not None else None
        y = K.torgb_skip_fwd(out, ws_rgb, bias_rgb, y_prev, words, bool(nchw))
        ctx.save_for_backward(x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias, ws_rgb, words)
        ctx.spec, ctx.gain, ctx.nchw, ctx.has_prev = spec, gain, bool(nchw), y_prev is not None
        ctx.set_materialize_grads(False)
        return out, y

    @staticmethod
    def backward(ctx, g_out, g_y):
        from . import upfirdn as U

        x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias, ws_rgb, words = ctx.saved_tensors
        spec = ctx.spec
        g = spec.geom
        gws = gbias_rgb = g_prev = None
        if g_y is not None:
            gr = g_y.contiguous().float()
            if ctx.nchw:
                gr = K.image_grad_nhwc(gr, words)
            elif words is not None:
                gr = K.image_grad_nhwc(gr.permute(0, 3, 1, 2).contiguous(), words)
            gy0, S1, Spre, Snz, gws = K.bias_act_rgb_bwd(g_out.contiguous() if g_out is not None else None, out, gr, ws_rgb,
                                                         noise=noise.contiguous(), d=d, act=1, gain=ctx.gain)
            gbias_rgb = gr.sum(dim=(0, 1, 2))
            g_prev = U.upsample_2d_nhwc_adjoint(gr) if ctx.has_prev else None
        else:
            gy0, S1, Spre, Snz = K.bias_act_bwd(g_out.contiguous(), out, noise=noise.contiguous(), d=d, act=True,
                                                gain=ctx.gain)
        t, gbias, gns, gs = K.demod_bwd(S1, Spre, Snz, d, ns.reshape(1), _aligned_vec(bias), s, q)
        K.PROFILE_TAG = (g.tag, g.algo_frac)
        gxs = K.conv2d_igemm(gy0, wadj, **g.adjoint().kernel_kwargs())
        gwmat = K.conv2d_wgrad(xs, gy0, **g.kernel_kwargs())
        gx, gs = K.modulate_bwd(gxs, x, s, gs_init=gs)
        gw_raw = K.wfold(gwmat, spec, w_raw=w_raw, s=s, t=t)
        return (gx, gs, gw_raw, None, gns.reshape(ns.shape), gbias, None, None, gws, gbias_rgb, g_prev, None, None)


class ModUpConvAct(torch.autograd.Function):
    """ModulatedConv2D(up=True) + Noise + BiasAct with upsample_conv_2d in the reference's own order
    (upfirdn_2d_v2.py:65-103): transposed stride-2 3x3 convolution on the tensor cores at its algorithmic
    cost, then the 4x4 FIR as one bandwidth-bound pass that also applies demodulation, noise, bias and
    leaky-ReLU.  ``spec`` is weight_spec("upT", ...)."""

    @staticmethod
    def forward(ctx, x, s, w_raw, noise, ns, bias, spec, gain: float):
        g = spec.geom
        x = x.contiguous()
        s = s.contiguous()
        wmat, wadj, q = _prepared(w_raw, spec, True, True)
        d = K.demod_coef(s, q)
        xs = K.modulate(x, s)
        if FOLD_UP_FORWARD:
            # forward with the FIR folded into the weights: ONE 4-phase 3x3 convolution on the input grid (halo-reuse
            # kernel when the grid is a multiple of 16 x 16) with the whole layer epilogue; no [B,2h+2,2w+2,O]
            # intermediate and no FIR pass.  4x the algorithmic FLOPs, but measured faster than transposed conv + FIR
            # (profiles/r02b_*: 64x32x128x128: 313 us + 280 us unfolded vs one launch at ~1.3 PFLOP/s executed).
            from .conv import weight_spec

            fspec = weight_spec("up", g.H, g.W, spec.I, spec.O, 3, True, g.tag)
            wf, _, _ = _prepared(w_raw, fspec, False, False)
            K.PROFILE_TAG = (g.tag, fspec.geom.algo_frac)
            out = K.conv2d_igemm(xs, wf, **fspec.geom.kernel_kwargs(), col_scale=d, noise=noise.contiguous(),
                                 noise_strength=ns.reshape(1), bias=_aligned_vec(bias), act=1, act_gain=gain)
        else:
            K.PROFILE_TAG = (g.tag, g.algo_frac)
            T = K.conv2d_igemm(xs, wmat, **spec.fwd_kwargs)                       # [B, 2h+2, 2w+2, O]
            out = K.fir4(T, spec.out_hw, (-1, -1), 1.0 / 16.0, d=d, noise=noise.contiguous(),
                         noise_strength=ns.reshape(1), bias=_aligned_vec(bias), act=1, gain=gain)
        ctx.save_for_backward(x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias)
        ctx.spec, ctx.gain = spec, gain
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, xs, out, s, d, wadj, q, w_raw, noise, ns, bias = ctx.saved_tensors
        spec = ctx.spec
        g = spec.geom
        gy0, S1, Spre, Snz = K.bias_act_bwd(g_out.contiguous(), out, noise=noise.contiguous(), d=d, act=True,
                                            gain=ctx.gain)
        t, gbias, gns, gs = K.demod_bwd(S1, Spre, Snz, d, ns.reshape(1), _aligned_vec(bias), s, q)
        gT = K.fir4(gy0, spec.t_hw, (-2, -2), 1.0 / 16.0)                        # adjoint of the FIR pass
        K.PROFILE_TAG = (g.tag, 1.0)
        gxs = K.conv2d_igemm(gT, wadj, **spec.s2_kwargs)                          # stride-2 3x3 VALID conv
        gwadj = K.conv2d_wgrad(gT, xs, **spec.s2_kwargs)                          # roles exchanged: [I, 9*O]
        gx, gs = K.modulate_bwd(gxs, x, s, gs_init=gs)
        gw_raw = K.wfold_adj(gwadj, spec, w_raw=w_raw, s=s, t=t, flip=True)
        return gx, gs, gw_raw, None, gns.reshape(ns.shape), gbias, None, None


class ConvAct(torch.autograd.Function):
    """out = lrelu(conv(x, w) + bias)*gain (+ residual)   |   out = conv(x, w) when bias is None; ``w`` is
    the fp32 HWIO master weight (equalised-LR coefficient and FIR folding happen in tbg_wprep).  A spec
    with a ``fir`` pre-pass (conv_downsample_2d, weight_spec("downU")) filters x once, convolves and takes
    the weight gradient on the filtered tensor, and keeps the folded form for the input gradient."""

    @staticmethod
    def forward(ctx, x, w_raw, bias, residual, spec, gain: float):
        geom = spec.geom
        x = x.contiguous()
        has_act = bias is not None
        need_gx = ctx.needs_input_grad[0]
        wmat, wadj, _ = _prepared(w_raw, spec, need_gx, False)
        if spec.fir is not None:
            x = K.fir4(x, spec.fir["out_hw"], spec.fir["off"], spec.fir["scale"])
        K.PROFILE_TAG = (geom.tag, geom.algo_frac)
        out = K.conv2d_igemm(x, wmat, **spec.fwd_kwargs, bias=bias, act=1 if has_act else 0,
                             act_gain=gain if has_act else 1.0,
                             residual=residual.contiguous() if residual is not None else None, res_scale=1.0)
        ctx.save_for_backward(x, wadj, out if has_act else None, residual if has_act else None)
        ctx.spec, ctx.gain, ctx.has_act, ctx.has_res = spec, gain, has_act, residual is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, wadj, out, residual = ctx.saved_tensors          # x: the (filtered) tensor the convolution read
        spec = ctx.spec
        g = spec.geom
        g_out = g_out.contiguous()
        want_w = ctx.needs_input_grad[1] and g.tag not in _SKIP_WGRAD_TAGS
        gbias = None
        nb = g_out.shape[0]
        gx_shape = (nb, g.H, g.W, g.cin) if spec.fir is None else \
            (nb, spec.fir["out_hw"][0] - (2 if spec.KH == 3 else 0), spec.fir["out_hw"][1] - (2 if spec.KH == 3 else 0), g.cin)
        lim = _BATCH_LIMIT.get(g.tag)
        if lim is not None and nb > lim and not want_w:
            # cotangent is zero beyond the first ``lim`` samples: run on that slice only
            gy0 = g_out[:lim]
            if ctx.has_act:
                gy0, _, _, _ = K.bias_act_bwd(gy0, out[:lim], residual=residual[:lim] if residual is not None else None,
                                              act=True, gain=ctx.gain, want_sums=False)
            gx = None
            if ctx.needs_input_grad[0]:
                gx = _limited_grad(gx_shape, g_out)
                K.PROFILE_TAG = (g.tag, spec.adj_frac)
                K.conv2d_igemm(gy0, wadj, **spec.adj_kwargs, out=gx[:lim])
            return gx, None, None, (g_out if ctx.has_res else None), None, None
        if ctx.has_act:
            gy0, gbias, _, _ = K.bias_act_bwd(g_out, out, residual=residual, act=True, gain=ctx.gain,
                                              want_sums=False, bias_grad_only=want_w)
        else:
            gy0 = g_out
        K.PROFILE_TAG = (g.tag, spec.adj_frac)
        gx = K.conv2d_igemm(gy0, wadj, **spec.adj_kwargs) if ctx.needs_input_grad[0] else None
        gw_raw = None
        if want_w:
            K.PROFILE_TAG = (g.tag, g.algo_frac)
            gw_raw = K.wfold(K.conv2d_wgrad(x, gy0, **spec.fwd_kwargs), spec)
        return gx, gw_raw, gbias, (g_out if ctx.has_res else None), None, None


class SkipSplit(torch.autograd.Function):
    """Entry of a discriminator residual block (discriminator.py:126-131): returns the block input for the main branch
    and, for the skip branch, the 4x4 FIR of conv_downsample_2d (upfirdn_2d_v2.py:106-113) evaluated only at the pixels
    the stride-(sy, 2) 1x1 convolution reads — xd[p,q] = sum k[m] k[n] x[sy*p+m-1, 2*q+n-1] / 64 — so that the 1x1
    convolution, its input gradient and its weight gradient are plain GEMMs on the small grid.  backward adds the
    transposed filter of the skip gradient to the main-branch gradient in one pass."""

    @staticmethod
    def forward(ctx, x, sy: int):
        x = x.contiguous()
        B, H, W_, _ = x.shape
        ctx.sy, ctx.hw = sy, (H, W_)
        xd = K.fir4_down(x, (H // sy, W_ // 2), sy, (-1, -1), 1.0 / 64.0)
        return x.view_as(x), xd

    @staticmethod
    def backward(ctx, g_main, g_xd):
        if g_xd is None:
            return g_main, None
        g_xd = g_xd.contiguous()
        add = g_main.contiguous() if g_main is not None else None
        nb = g_xd.shape[0]
        lim = _BATCH_LIMIT.get("dconv")
        if lim is not None and nb > lim and "dconv" in _SKIP_WGRAD_TAGS:
            # generator pass through a concatenated (fake, real) evaluation: the cotangent is zero beyond ``lim`` samples
            gx = _limited_grad((nb,) + tuple(ctx.hw) + (g_xd.shape[3],), g_xd)
            K.fir4_down_adjoint(g_xd[:lim], ctx.hw, ctx.sy, (-1, -1), 1.0 / 64.0, add=add[:lim] if add is not None else None,
                                out=gx[:lim])
            return gx, None
        return K.fir4_down_adjoint(g_xd, ctx.hw, ctx.sy, (-1, -1), 1.0 / 64.0, add=add), None


class FromRGB(torch.autograd.Function):
    """FromRGB.call (from_rgb.py:26-29) + BiasAct on the NCHW fp32 image -> NHWC bf16, one launch each way."""

    @staticmethod
    def forward(ctx, img, w_raw, bias, coef: float, gain: float):
        img = img.contiguous()
        w2 = w_raw.reshape(3, -1).contiguous()
        out = K.fromrgb_fwd(img, w2, _aligned_vec(bias), coef, gain)
        ctx.save_for_backward(img, w2, out)
        ctx.coef, ctx.gain, ctx.wshape = coef, gain, w_raw.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, w2, out = ctx.saved_tensors
        g_out = g_out.contiguous()
        want_w = ctx.needs_input_grad[1] and "dconv" not in _SKIP_WGRAD_TAGS
        lim = _BATCH_LIMIT.get("dconv")
        if lim is not None and g_out.shape[0] > lim and not want_w:
            gimg = None
            if ctx.needs_input_grad[0]:
                gimg = _limited_grad(img.shape, img)
                gi, _, _ = K.fromrgb_bwd(img[:lim], w2, g_out[:lim], out[:lim], ctx.coef, ctx.gain, want_img=True, want_w=False)
                gimg[:lim].copy_(gi)
            return gimg, None, None, None, None
        gimg, gw, gb = K.fromrgb_bwd(img, w2, g_out, out, ctx.coef, ctx.gain, want_img=ctx.needs_input_grad[0],
                                     want_w=want_w)
        return gimg, (gw.reshape(ctx.wshape) if gw is not None else None), gb, None, None


class StyleScales(torch.autograd.Function):
    """All style projections of the synthesis network at once: for layer l,
    s_l = mod_bias(mod_dense(style[:, idx_l])) + 1 (modulated_conv2d.py:75-76).  Inputs after the
    layout arguments are w_0, b_0, w_1, b_1, ...; outputs one [B, I_l] tensor per layer."""

    @staticmethod
    def forward(ctx, style, idxs, coef, *wb):
        style = style.contiguous()
        ws, bs = [w.contiguous() for w in wb[0::2]], [b.contiguous() for b in wb[1::2]]
        outs = K.style_dense_fwd(style, ws, bs, idxs, coef)
        ctx.save_for_backward(style, *ws)
        ctx.idxs, ctx.coef = idxs, coef
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gss):
        style, *ws = ctx.saved_tensors
        gss = [g.contiguous() if g is not None else torch.zeros((style.shape[0], w.shape[1]), device=style.device)
               for g, w in zip(gss, ws)]
        gstyle, gws, gbs = K.style_dense_bwd(style, ws, gss, ctx.idxs, ctx.coef)
        out = [gstyle, None, None]
        for gw, gb in zip(gws, gbs):
            out += [gw, gb]
        return tuple(out)


class ToRGB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ws, bias):
        x = x.contiguous()
        ws = ws.contiguous()
        ctx.save_for_backward(x, ws)
        return K.torgb_fwd(x, ws, bias)

    @staticmethod
    def backward(ctx, gy):
        x, ws = ctx.saved_tensors
        gy = gy.contiguous()
        gx, gws = K.torgb_bwd(x, ws, gy)
        return gx, gws, gy.sum(dim=(0, 1, 2))


class DenseAct(torch.autograd.Function):
    """Dense.call (dense.py:23-29) [+ bias * bias_coef + activation * gain] on the fp32 ``tbg_dense_*`` kernels:
    y = act((x @ w) * coef + bias * bias_coef) * gain.  ``tag``: weight gradients are skipped while that tag is listed in
    :class:`skip_weight_grads` (discriminator layers during the generator's backward pass)."""

    @staticmethod
    def forward(ctx, x, w, bias, coef: float, bias_coef: float, act: int, gain: float, tag: str = ""):
        x = x.contiguous().float()
        y = K.dense_fwd(x, w.contiguous(), bias, coef=coef, bias_coef=bias_coef, act=act, gain=gain)
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (coef, bias_coef, act, gain, tag, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        coef, bias_coef, act, gain, tag, has_bias = ctx.cfg
        want_w = tag not in _SKIP_WGRAD_TAGS
        gx, gw, gb = K.dense_bwd(x, w.contiguous(), y, gy.contiguous().float(), coef=coef, bias_coef=bias_coef, act=act,
                                 gain=gain, want_gx=ctx.needs_input_grad[0], want_gw=ctx.needs_input_grad[1] and want_w,
                                 want_gb=has_bias and ctx.needs_input_grad[2] and want_w)
        return gx, gw, gb, None, None, None, None, None


class PixelNorm(torch.autograd.Function):
    """x * rsqrt(mean(x^2) + 1e-8) per row (mapping_block.py:15-18)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        ctx.save_for_backward(x)
        return K.pixel_norm_fwd(x)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return K.pixel_norm_bwd(x, gy.contiguous())


class WordEncoderFn(torch.autograd.Function):
    """WordEncoder.call (word_encoder.py:39-63) in one launch each way; returns the base feature map NHWC bf16."""

    @staticmethod
    def forward(ctx, words, w0, table, mask, fc_w, fc_b, keep: float, out_hwc):
        words = words.to(torch.int32).contiguous()
        mask = mask.contiguous().float() if mask is not None else None
        out, emb, act = K.word_encoder_fwd(words, w0.contiguous(), table.contiguous(), mask, keep, fc_w.contiguous(),
                                           fc_b.contiguous(), out_hwc)
        ctx.save_for_backward(words, mask, fc_w, emb, act)
        ctx.cfg = (keep, tuple(out_hwc), table.shape[0])
        return out

    @staticmethod
    def backward(ctx, g_out):
        words, mask, fc_w, emb, act = ctx.saved_tensors
        keep, out_hwc, rows = ctx.cfg
        g_table, g_fc_w, g_fc_b = K.word_encoder_bwd(words, mask, keep, fc_w.contiguous(), emb, act, g_out.contiguous(), rows,
                                                     out_hwc)
        return None, None, g_table, None, g_fc_w, g_fc_b, None, None


class MinibatchStdCat(torch.autograd.Function):
    """MinibatchStd.call (mini_batch_std.py:10-35): x bf16 [n_calls*B,H,W,C] -> [x | statistic | zero padding to cpad]."""

    @staticmethod
    def forward(ctx, x, n_calls: int, cpad: int):
        x = x.contiguous()
        xcat, _ = K.minibatch_std_fwd(x, n_calls, cpad)
        ctx.save_for_backward(x)
        ctx.n_calls = n_calls
        return xcat

    @staticmethod
    def backward(ctx, gxcat):
        (x,) = ctx.saved_tensors
        return K.minibatch_std_bwd(x, gxcat.contiguous(), ctx.n_calls), None, None


class R1SqNorm(torch.autograd.Function):
    """Per-sample squared norm of the image gradient (training_step.py:369)."""

    @staticmethod
    def forward(ctx, g):
        g = g.contiguous().float()
        ctx.save_for_backward(g)
        return K.r1_sqnorm(g)

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return K.r1_sqnorm_bwd(g, gout.contiguous().float())


class ToRGBSkip(torch.autograd.Function):
    """ToRGB.call + the upsampled skip sum of SynthesisBlock.call (to_rgb.py:28-33, synthesis_block.py:152) and, for the
    last block, mask_text_box (utils/utils.py:11-45) + NHWC -> NCHW, in one launch (``tbg_torgb_skip_fwd``)."""

    @staticmethod
    def forward(ctx, x, ws, bias, y_prev, words, nchw: bool):
        x = x.contiguous()
        ws = ws.contiguous()
        if words is not None:
            words = words.to(torch.int32).contiguous()
        y_prev = y_prev.contiguous() if y_prev is not None else None
        ctx.save_for_backward(x, ws, words)
        ctx.nchw, ctx.has_prev = bool(nchw), y_prev is not None
        return K.torgb_skip_fwd(x, ws, bias, y_prev, words, bool(nchw))

    @staticmethod
    def backward(ctx, g):
        from . import upfirdn as U

        x, ws, words = ctx.saved_tensors
        g = g.contiguous().float()
        if ctx.nchw:
            g = K.image_grad_nhwc(g, words)
        elif words is not None:
            g = K.image_grad_nhwc(g.permute(0, 3, 1, 2).contiguous(), words)
        gx, gws = K.torgb_bwd(x, ws, g)
        g_prev = U.upsample_2d_nhwc_adjoint(g) if ctx.has_prev else None
        return gx, gws, g.sum(dim=(0, 1, 2)), g_prev, None, None
