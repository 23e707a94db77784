"""Discriminator (mirror of models/custom_stylegan2/discriminator.py:11-217) on the B200 kernels.

ResNet discriminator on ``[B,3,H,W]`` images: FromRGB -> blocks {conv3x3 + bias-lrelu;
blur + conv3x3 stride (2|1,2) + bias-lrelu; skip = blur + conv1x1 same stride; (x+skip)/sqrt2}
-> minibatch-std -> conv3x3 -> dense -> dense(1).  The blur + strided convolutions are folded
into single 6x6 / 4x4 stride-2 convolutions (conv.down_geom) whose input gradients are 4-phase
3x3 GEMMs, so no blurred intermediate is written to HBM.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import conv as C
from . import layers as L
from .config import Config
from .model_base import Model, fresh_seed

INV_SQRT2 = 1.0 / math.sqrt(2.0)


class Discriminator(Model):
    def __init__(self, cfg: Config, device="cuda", seed: Optional[int] = None):
        super().__init__("discriminator")
        self.cfg = cfg
        self.device = torch.device(device)
        self.resolutions = cfg.discrim_resolutions
        self.feat_maps = cfg.discrim_feat_maps
        self._build(seed)

    def _build(self, seed: Optional[int]) -> None:
        g = torch.Generator().manual_seed(seed if seed is not None else fresh_seed())

        def randn(*shape):
            return torch.randn(*shape, generator=g)

        res, fm = self.resolutions, self.feat_maps
        r0 = res[0]
        self.add_weight(f"{r0[0]}x{r0[1]}/FromRGB/conv/w", randn(1, 1, 3, fm[0]))
        self.add_weight(f"{r0[0]}x{r0[1]}/FromRGB/bias/b", torch.zeros(fm[0]))
        for (h, w), f0, f1 in zip(res[:-1], fm[:-1], fm[1:]):
            pb = f"{h}x{w}"
            self.add_weight(pb + "/conv_0/w", randn(3, 3, f0, f0))
            self.add_weight(pb + "/bias_0/b", torch.zeros(f0))
            self.add_weight(pb + "/conv_1/w", randn(3, 3, f0, f1))
            self.add_weight(pb + "/bias_1/b", torch.zeros(f1))
            self.add_weight(pb + "/skip/w", randn(1, 1, f0, f1))
        rf = res[-1]
        n_f0, n_f1 = fm[-2], fm[-1]                                          # discriminator.py:193
        pl = f"{rf[0]}x{rf[1]}/last"
        self.add_weight(pl + "/conv_0/w", randn(3, 3, n_f0 + 1, n_f0))
        self.add_weight(pl + "/bias_0/b", torch.zeros(n_f0))
        self.add_weight(pl + "/dense_1/w", randn(n_f0 * rf[0] * rf[1], n_f1))
        self.add_weight(pl + "/bias_1/b", torch.zeros(n_f1))
        self.add_weight("last_dense/w", randn(n_f1, 1))
        self.add_weight("last_bias/b", torch.zeros(1))
        self.to(self.device)

    # ------------------------------------------------------------------------------------------
    def _conv(self, x, name: str, *, bias: Optional[str], down: bool, reduce_height: bool = False,
              residual: Optional[torch.Tensor] = None, scale: float = 1.0):
        """Conv2D.call (conv.py:51-73) + BiasAct (bias_act.py:25-34) [+ residual merge :82]."""
        P = self.params
        w_raw = P[name]
        k, _, I, O = w_raw.shape
        B, H, W_, _ = x.shape
        if L.use_fused():
            from .fused import ConvAct

            kind = ("downU" if L.UNFOLD_DOWN else "down") if down else "plain"
            spec = C.weight_spec(kind, H, W_, I, O, k, reduce_height, "dconv", scale)
            if residual is not None:
                # (lrelu(v)*sqrt2 + skip)/sqrt2 == lrelu(v) + skip/sqrt2: the caller pre-scales the
                # (linear) skip branch, so the merge is a plain residual add in the epilogue
                return ConvAct.apply(x, w_raw, P[bias], residual, spec, 1.0)
            return ConvAct.apply(x, w_raw, P[bias] if bias else None, None, spec, L.SQRT2)
        if torch.is_grad_enabled() and L.SPEC_SECOND_ORDER and scale == 1.0 and (not down or (H % 2 == 0 and W_ % 2 == 0)):
            # regulariser pass (R1, training_step.py:363-368): closed bilinear triple on the master weight
            from . import second_order as SO

            if down and k == 1:
                y = SO.lin_conv(SO.fir4_down(x, 2 if reduce_height else 1), w_raw, "plain", 1, True, "dconv")
            else:
                y = SO.lin_conv(x, w_raw, "downU" if down else "plain", k, reduce_height, "dconv")
            if bias is not None:
                y = SO.bias_act(y, None, None, P[bias], 1, L.SQRT2)
            if residual is not None:
                y = (y + residual) * INV_SQRT2
            return y
        w = (L.runtime_coef(w_raw.shape) * scale) * w_raw
        if down:
            geom = C.down_geom(H, W_, I, O, k, reduce_height, tag="dconv")
            wmat = C.down_wmat(w)
        else:
            geom = C.plain_geom(H, W_, I, O, k, tag="dconv")
            wmat = C.plain_wmat(w)
        if not torch.is_grad_enabled():
            epi = dict(bias=P[bias].contiguous() if bias else None, act=1 if bias else 0,
                       act_gain=L.SQRT2 if bias else 1.0,
                       residual=residual.contiguous() if residual is not None else None,
                       res_scale=INV_SQRT2 if residual is not None else 1.0)
            return C.conv(x, wmat, geom, epi)
        y = C.conv(x, wmat, geom)
        if bias is not None:
            from . import second_order as SO

            y = SO.bias_act(y, None, None, P[bias], 1, L.SQRT2)            # twice differentiable (R1, training_step.py:363-368)
        if residual is not None:
            y = (y + residual) * INV_SQRT2
        return y

    def __call__(self, images: torch.Tensor, n_calls: int = 1) -> torch.Tensor:
        """discriminator.py:202-214: [B,3,H,W] fp32 -> [B,1] fp32.  ``n_calls`` > 1 evaluates that many
        independent calls concatenated on the batch axis in one pass (every layer is per-sample except
        the minibatch statistic, which is then taken per call)."""
        P = self.params
        res = self.resolutions
        r0 = res[0]
        w0 = P[f"{r0[0]}x{r0[1]}/FromRGB/conv/w"]
        if L.use_fused():
            from .fused import FromRGB

            # FromRGB (from_rgb.py:26-29): K = 3 -> bandwidth-bound, one fused launch (NCHW fp32 -> NHWC bf16)
            x = FromRGB.apply(images.float(), w0, P[f"{r0[0]}x{r0[1]}/FromRGB/bias/b"], L.runtime_coef(w0.shape), L.SQRT2)
        else:
            x = images.permute(0, 2, 3, 1).float()                           # NHWC
            x = x @ (L.runtime_coef(w0.shape) * w0[0, 0]) + P[f"{r0[0]}x{r0[1]}/FromRGB/bias/b"]
            x = L.lrelu(x).to(L.ACT_DTYPE)
        for (h, w), (nh, nw) in zip(res[:-1], res[1:]):
            pb = f"{h}x{w}"
            rh = h != nh
            if L.use_fused() and L.SKIP_SPLIT and h % (2 if rh else 1) == 0 and w % 2 == 0:
                from .fused import SkipSplit

                # FIR evaluated at the strided pixels only, then a plain 1x1 convolution on the small grid
                x, xd = SkipSplit.apply(x, 2 if rh else 1)
                skip = self._conv(xd, pb + "/skip/w", bias=None, down=False, scale=INV_SQRT2)
            else:
                skip = self._conv(x, pb + "/skip/w", bias=None, down=True, reduce_height=rh,
                                  scale=INV_SQRT2 if L.use_fused() else 1.0)
            x = self._conv(x, pb + "/conv_0/w", bias=pb + "/bias_0/b", down=False)
            x = self._conv(x, pb + "/conv_1/w", bias=pb + "/bias_1/b", down=True, reduce_height=rh, residual=skip)
        rf = res[-1]
        pl = f"{rf[0]}x{rf[1]}/last"
        # minibatch-std feature (mini_batch_std.py): one extra constant channel per sample; padded
        # with zero channels up to a multiple of 64 so that K stays TMA/UMMA aligned
        B, H, W_, Cc = x.shape
        cpad = (Cc + 1 + 63) // 64 * 64
        if L.use_fused():
            from .fused import MinibatchStdCat

            xcat = MinibatchStdCat.apply(x, n_calls, cpad)                   # [x | statistic | zero padding], one launch
        else:
            std = L.minibatch_std(x, n_calls=n_calls)                        # [B,1]
            xcat = torch.cat([x, std.to(x.dtype)[:, None, None, :].expand(B, H, W_, 1),
                              x.new_zeros(B, H, W_, cpad - Cc - 1)], dim=3).contiguous()
        w_raw = P[pl + "/conv_0/w"]                                          # [3,3,C+1,C]
        if L.use_fused():
            from .fused import ConvAct

            spec = C.weight_spec("plain", H, W_, Cc + 1, w_raw.shape[3], 3, True, "dconv")   # K padded 513 -> 576
            y = ConvAct.apply(xcat, w_raw, P[pl + "/bias_0/b"], None, spec, L.SQRT2).float()
        else:
            if torch.is_grad_enabled() and L.SPEC_SECOND_ORDER:
                from . import second_order as SO

                y = SO.lin_conv(xcat, w_raw, "plain", 3, True, "dconv").float()            # K padded 513 -> 576 by the spec
            else:
                w = L.runtime_coef(w_raw.shape) * w_raw
                wpad = torch.cat([w, w.new_zeros(3, 3, cpad - Cc - 1, w.shape[3])], dim=2)
                y = C.conv(xcat, C.plain_wmat(wpad), C.plain_geom(H, W_, cpad, w.shape[3], 3, tag="dconv",
                                                                  algo_frac=(Cc + 1) / cpad)).float()
            y = L.lrelu(y + P[pl + "/bias_0/b"])
        # flatten in the reference's NCHW order (dense.py:26-27 on an NCHW tensor)
        y = y.permute(0, 3, 1, 2).reshape(B, -1)
        if L.use_fused():
            from .fused import DenseAct

            w1, w2 = P[pl + "/dense_1/w"], P["last_dense/w"]
            y = DenseAct.apply(y, w1, P[pl + "/bias_1/b"], L.runtime_coef(w1.shape), 1.0, 1, L.SQRT2, "dconv")
            return DenseAct.apply(y, w2, P["last_bias/b"], L.runtime_coef(w2.shape), 1.0, 0, 1.0, "dconv")
        y = L.lrelu(L.dense(y, P[pl + "/dense_1/w"]) + P[pl + "/bias_1/b"])
        y = L.dense(y, P["last_dense/w"]) + P["last_bias/b"]
        return y
