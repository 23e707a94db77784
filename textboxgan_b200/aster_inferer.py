"""AsterInferer — the frozen OCR head used as a loss (mirror of
aster_ocr_utils/aster_inferer.py:7-190).

**Parity unpinned**: the reference loads ASTER from an external SavedModel that is not in the
repository (aster_inferer.py:24-26; aster_weights/ holds only ``.keep``).  This class keeps the
reference's wrapper behaviour (``convert_inputs`` :153-190, batched ``call`` :28-37,
``_postprocess_simple`` :116-151) and runs the published ASTER recognition topology (45-layer
ResNet encoder, 2 x BiLSTM-256, Bahdanau-attention LSTM decoder, 96 classes) with frozen weights —
seeded synthetic ones unless a weight dictionary is supplied through ``weights=`` (the import hook
for converted ASTER checkpoints).  See oracle/aster.py for the documented simplifications.

The head is frozen but differentiable w.r.t. its input: the OCR loss is back-propagated through
it into the generator (training_step.py:201-206), so every layer has forward + input-gradient and
no weight-gradient.  Convolutions run on the tcgen05 kernel with fused bias + ReLU (+ residual)
epilogues; their input gradients are the same kernel on pre-computed adjoint weights.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import conv as C
from . import kernels as K
from . import layers as L
from .config import Config

NUM_CLASSES = 96
ENC_BLOCKS = [(32, 3, (2, 2)), (64, 4, (2, 2)), (128, 6, (2, 1)), (256, 6, (2, 1)), (512, 3, (2, 1))]
LSTM_HIDDEN = 256
ATT_UNITS = 256
EMB_DIM = 256


def init_aster_params(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Seeded synthetic frozen weights; identical construction (names, shapes, RNG order) to
    oracle/aster.py::init_aster_params so that both sides can be built from one seed."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}

    def conv(name, k, cin, cout, gain=2.0):
        std = math.sqrt(gain / (k * k * cin))
        P[name + "/w"] = (torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * std).float()
        P[name + "/b"] = (torch.randn(cout, generator=g, dtype=torch.float64) * 0.01).float()

    conv("enc/stem", 3, 3, 32)
    cin = 32
    for bi, (ch, units, _stride) in enumerate(ENC_BLOCKS):
        for ui in range(units):
            pre = f"enc/b{bi}/u{ui}"
            conv(pre + "/c1", 1, cin, ch)
            conv(pre + "/c2", 3, ch, ch, gain=1.0)
            if ui == 0:
                conv(pre + "/sc", 1, cin, ch, gain=1.0)
            cin = ch

    def lstm(name, in_dim, hid):
        s = 1.0 / math.sqrt(hid)
        P[name + "/w_ih"] = ((torch.rand(in_dim, 4 * hid, generator=g, dtype=torch.float64) * 2 - 1) * s).float()
        P[name + "/w_hh"] = ((torch.rand(hid, 4 * hid, generator=g, dtype=torch.float64) * 2 - 1) * s).float()
        P[name + "/b"] = torch.zeros(4 * hid)

    for li in range(2):
        for d in ("fw", "bw"):
            lstm(f"rnn/l{li}/{d}", 512, LSTM_HIDDEN)

    def lin(name, i, o, bias=True):
        s = 1.0 / math.sqrt(i)
        P[name + "/w"] = ((torch.rand(i, o, generator=g, dtype=torch.float64) * 2 - 1) * s).float()
        if bias:
            P[name + "/b"] = torch.zeros(o)

    lin("dec/memory_layer", 512, ATT_UNITS, bias=False)
    lin("dec/query_layer", LSTM_HIDDEN, ATT_UNITS, bias=False)
    P["dec/attention_v"] = ((torch.rand(ATT_UNITS, generator=g, dtype=torch.float64) * 2 - 1)
                            / math.sqrt(ATT_UNITS)).float()
    P["dec/embedding"] = (torch.randn(NUM_CLASSES, EMB_DIM, generator=g, dtype=torch.float64) * 0.1).float()
    lstm("dec/lstm_cell", EMB_DIM + 512, LSTM_HIDDEN)
    lin("dec/dense", LSTM_HIDDEN + 512, NUM_CLASSES)
    return P


# the ResNet encoder as one autograd node with ReLU masks / residual sums fused into the input-gradient convs
FUSED_ENCODER = True
# convert_inputs as the tbg_crop_resize kernels (first order) instead of two batched interpolation GEMMs
FUSED_CONVERT_INPUTS = True


def _pad_c(n: int) -> int:
    return (n + 63) // 64 * 64


class _FrozenConv(torch.autograd.Function):
    """relu?(conv(x, W) + b (+ residual)) with frozen W, b: one fused kernel forward; the input
    gradient is the ReLU mask followed by the same kernel on the adjoint weights."""

    @staticmethod
    def forward(ctx, x, layer: "_ConvLayer", residual, relu: bool):
        K.PROFILE_TAG = (layer.geom.tag, layer.geom.algo_frac)
        out = K.conv2d_igemm(x.contiguous(), layer.wmat, **layer.geom.kernel_kwargs(), bias=layer.bias,
                             residual=residual.contiguous() if residual is not None else None, res_scale=1.0,
                             res_first=True, act=2 if relu else 0)
        ctx.layer = layer
        ctx.relu = relu
        ctx.has_res = residual is not None
        ctx.save_for_backward(out if relu else None)
        return out

    @staticmethod
    def backward(ctx, gy):
        layer = ctx.layer
        (out,) = ctx.saved_tensors
        g = gy.contiguous()
        if ctx.relu:
            g = K.bias_act_bwd(g, out, act=2, gain=1.0, want_sums=False)[0]
        K.PROFILE_TAG = (layer.geom.tag, layer.geom.algo_frac)
        gx = K.conv2d_igemm(g, layer.wmat_adj, **layer.geom.adjoint().kernel_kwargs())
        return gx, None, (g if ctx.has_res else None), None


class _EncoderFn(torch.autograd.Function):
    """The whole frozen ResNet encoder as one node: forward = the fused conv+bias+(residual)+ReLU launches;
    backward = one input-gradient launch per convolution with the ReLU masks and the residual-branch sums
    folded into the epilogues (no stand-alone ReLU-backward or add kernels):
        unit:  y = relu(c1(x));  out = relu(c2(y) + sc(x))          (sc = 1x1 conv, or identity)
        gz given (gradient w.r.t. the unit's pre-activation):
            gy  = c2^T(gz) masked by y ;   gx = (c1^T(gy) + sc^T(gz)) masked by x   -> gz of the unit below."""

    @staticmethod
    def forward(ctx, x, enc: "AsterInferer"):
        x = x.contiguous()
        run = lambda layer, t, residual=None, relu=True: K.conv2d_igemm(
            t, layer.wmat, **layer.geom.kernel_kwargs(), bias=layer.bias, residual=residual, res_scale=1.0,
            res_first=True, act=2 if relu else 0)
        K.PROFILE_TAG = ("aster", 1.0)
        saved = [x]
        h = run(enc.stem, x)
        for c1, c2, sc in enc.units:
            y = run(c1, h)
            shortcut = run(sc, h, relu=False) if sc is not None else h
            out = run(c2, y, residual=shortcut)
            saved += [h, y]
            h = out
        saved.append(h)
        ctx.enc = enc
        ctx.save_for_backward(*saved)
        return h

    @staticmethod
    def backward(ctx, g):
        enc = ctx.enc
        saved = ctx.saved_tensors
        out = saved[-1]
        K.PROFILE_TAG = ("aster", 1.0)
        dgrad = lambda layer, t, **kw: K.conv2d_igemm(t, layer.wmat_adj, **layer.geom.adjoint().kernel_kwargs(), **kw)
        gz = K.bias_act_bwd(g.contiguous(), out, act=2, gain=1.0, want_sums=False)[0]     # mask of the last ReLU
        for ui in range(len(enc.units) - 1, -1, -1):
            c1, c2, sc = enc.units[ui]
            h, y = saved[1 + 2 * ui], saved[2 + 2 * ui]
            gy = dgrad(c2, gz, relu_mask=y)
            gsc = dgrad(sc, gz) if sc is not None else gz
            gz = dgrad(c1, gy, residual=gsc, res_scale=1.0, res_first=True, relu_mask=h)
        return dgrad(enc.stem, gz), None


class _CropResize(torch.autograd.Function):
    """convert_inputs as one gather launch forward and one scatter launch backward (first order only)."""

    @staticmethod
    def forward(ctx, images, labels, blank_label, char_width, out_hw):
        images = images.float().contiguous()
        labels = labels.to(torch.int32).contiguous()
        ctx.save_for_backward(labels)
        ctx.args = (int(blank_label), char_width, (images.shape[2], images.shape[3]))
        return K.crop_resize_fwd(images, labels, int(blank_label), char_width, out_hw)

    @staticmethod
    def backward(ctx, g):
        (labels,) = ctx.saved_tensors
        blank, cw, hw = ctx.args
        return K.crop_resize_bwd(g.float().contiguous(), labels, blank, cw, hw), None, None, None, None


class _ConvLayer:
    def __init__(self, w_hwio: torch.Tensor, b: torch.Tensor, H: int, W: int, stride=(1, 1), device="cuda"):
        k, _, cin, cout = w_hwio.shape
        cin_p, cout_p = _pad_c(cin), _pad_c(cout)
        wp = torch.zeros(k, k, cin_p, cout_p)
        wp[:, :, :cin, :cout] = w_hwio
        bp = torch.zeros(cout_p)
        bp[:cout] = b
        kinds = ["s2" if s == 2 else "s1" for s in stride]
        assert k == 1 or stride == (1, 1)
        self.geom = C.ConvGeom(H, W, cin_p, cout_p, C.Axis(kinds[0], k, k // 2), C.Axis(kinds[1], k, k // 2), "aster",
                               (cin * cout) / float(cin_p * cout_p))
        wmat = C.plain_wmat(wp).to(device)
        self.wmat = wmat.to(L.ACT_DTYPE).contiguous()
        self.wmat_adj = C.relayout_for_adjoint(wmat, self.geom).to(L.ACT_DTYPE).contiguous()
        self.bias = bp.to(device).contiguous()

    def __call__(self, x, residual=None, relu=True):
        return _FrozenConv.apply(x, self, residual, relu)


class _LstmLayer:
    """Frozen BiLSTM layer: stacked input projections + recurrent weights packed for the kernels."""

    def __init__(self, P, name: str, device):
        self.w_ih = torch.stack([P[f"{name}/fw/w_ih"], P[f"{name}/bw/w_ih"]]).to(device).float()
        self.b = torch.stack([P[f"{name}/fw/b"], P[f"{name}/bw/b"]])[:, None, None, :].to(device).float()
        w_hh = torch.stack([P[f"{name}/fw/w_hh"], P[f"{name}/bw/w_hh"]]).to(device).float()    # [2,H,4H]
        D, H, _ = w_hh.shape
        w4 = w_hh.reshape(D, H, 4, H)                                              # [d, k, gate, j]
        self.w_packed = w4.permute(0, 1, 3, 2).contiguous().to(L.ACT_DTYPE)        # [d, k, j, gate]
        self.wT_packed = w4.permute(0, 3, 1, 2).contiguous().to(L.ACT_DTYPE)       # [d, j, k, gate]


class _LstmSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, layer: _LstmLayer):
        h, gates, c = K.lstm_seq_fwd(xp.contiguous(), layer.w_packed)
        ctx.layer = layer
        ctx.save_for_backward(gates, c)
        return h

    @staticmethod
    def backward(ctx, g_h):
        gates, c = ctx.saved_tensors
        return K.lstm_seq_bwd(g_h.contiguous(), gates, c, ctx.layer.wT_packed), None


class _AttnDecoder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mem, keys, w: dict, steps: int):
        logits, sv = K.attn_decoder_fwd(mem, keys, w, steps)
        ctx.w, ctx.sv = w, sv
        ctx.save_for_backward(mem, keys)
        return logits

    @staticmethod
    def backward(ctx, g_logits):
        mem, keys = ctx.saved_tensors
        g_mem, g_keys = K.attn_decoder_bwd(mem, keys, ctx.w, g_logits.contiguous().float(), ctx.sv)
        return g_mem, g_keys, None, None


def _pack_decoder(P, device) -> dict:
    """Frozen decoder weights in the layouts the kernels stream (bf16 matrices + their transposes)."""
    act = L.ACT_DTYPE
    m = lambda n: P[n].to(device).float()
    wq, wih, whh, wd = m("dec/query_layer/w"), m("dec/lstm_cell/w_ih"), m("dec/lstm_cell/w_hh"), m("dec/dense/w")
    wg = torch.cat([wih, whh], dim=0)                       # rows [emb | ctx | h]
    return dict(wq=wq.to(act).contiguous(), wqT=wq.t().to(act).contiguous(), v=m("dec/attention_v").contiguous(),
                emb=m("dec/embedding").to(act).contiguous(), wg=wg.to(act).contiguous(),
                wgT=wg.t().to(act).contiguous(), b=m("dec/lstm_cell/b").contiguous(), wd=wd.to(act).contiguous(),
                wdT=wd.t().to(act).contiguous(), bd=m("dec/dense/b").contiguous())


def load_aster_weights(path: str) -> Dict[str, torch.Tensor]:
    """Weight dictionary for :class:`AsterInferer` from ``path``: a ``.npz`` / ``.pt`` file holding the names of
    :func:`init_aster_params`, or a TensorFlow checkpoint prefix / SavedModel ``variables`` directory, read with the
    TensorBundle reader of :mod:`textboxgan_b200.tf_checkpoint` and mapped by name (README.md:60-66)."""
    import os

    import numpy as np

    if path.endswith(".npz"):
        with np.load(path) as z:
            P = {k: torch.from_numpy(np.asarray(z[k])).float() for k in z.files}
    elif path.endswith(".pt") or path.endswith(".pth"):
        P = {k: v.float() for k, v in torch.load(path, map_location="cpu").items()}
    else:
        from .tf_checkpoint import aster_weights_from_checkpoint

        P = aster_weights_from_checkpoint(path)
    want = init_aster_params(0)
    missing = sorted(set(want) - set(P))
    if missing:
        raise KeyError(f"load_aster_weights({path!r}): missing {len(missing)} tensors, e.g. {missing[:4]}")
    for k, v in want.items():
        if tuple(P[k].shape) != tuple(v.shape):
            raise ValueError(f"load_aster_weights: {k} has shape {tuple(P[k].shape)}, expected {tuple(v.shape)}")
    return {k: P[k] for k in want}


class AsterInferer:
    """Reads the word written in a text box (aster_inferer.py:7-37)."""

    def __init__(self, cfg: Config, device="cuda", combine_forward_and_backward: bool = False,
                 weights: Optional[Dict[str, torch.Tensor]] = None, seed: int = 1234,
                 synthetic_weights: Optional[bool] = None):
        """``weights``: a weight dictionary (names of :func:`init_aster_params`); else ``cfg.aster_weights`` (path of a
        converted weight file, see :func:`load_aster_weights`); else seeded synthetic weights — silently only when
        ``synthetic_weights=True`` (tests, bench), with a warning when it is None, an error when it is False (the
        product entry points Trainer / Infer pass ``cfg.aster_synthetic_weights``, default False: the reference
        always loads the pretrained recogniser, aster_inferer.py:24-26)."""
        if combine_forward_and_backward:
            raise NotImplementedError("combine_forward_and_backward=True needs the backward predictor of the "
                                      "external ASTER SavedModel (aster_inferer.py:39-114); the reference default "
                                      "and the training step use False")
        self.cfg = cfg
        self.device = torch.device(device)
        self.combine_forward_and_backward = combine_forward_and_backward
        if weights is None and getattr(cfg, "aster_weights", None):
            weights = load_aster_weights(cfg.aster_weights)
        if weights is None:
            if synthetic_weights is False:
                raise RuntimeError("AsterInferer: no recogniser weights (cfg.aster_weights is None).  The OCR loss would be "
                                   "computed against a random recogniser; set cfg.aster_weights to a converted weight file "
                                   "or opt in explicitly with cfg.aster_synthetic_weights = True")
            if synthetic_weights is None:
                import warnings

                warnings.warn("AsterInferer: cfg.aster_weights is None — using seeded SYNTHETIC recogniser weights; OCR "
                              "losses are not comparable with the reference's pretrained ASTER", RuntimeWarning, stacklevel=2)
            weights = init_aster_params(seed)
        P = weights
        self.P = {k: v.to(self.device).float() for k, v in P.items()}
        self._build_encoder(P)
        self.lstm = {n: _LstmLayer(P, n, self.device) for n in ("rnn/l0", "rnn/l1")}
        self.dec_w = _pack_decoder(P, self.device)

    # ------------------------------------------------------------------------------------------
    def _build_encoder(self, P) -> None:
        dev = self.device
        H, W = self.cfg.aster_image_dims[0] // 2, self.cfg.aster_image_dims[1] // 2   # after the 2x2 pool
        self.stem = _ConvLayer(P["enc/stem/w"], P["enc/stem/b"], H, W, device=dev)
        self.units = []
        for bi, (ch, units, stride) in enumerate(ENC_BLOCKS):
            for ui in range(units):
                pre = f"enc/b{bi}/u{ui}"
                s = stride if ui == 0 else (1, 1)
                c1 = _ConvLayer(P[pre + "/c1/w"], P[pre + "/c1/b"], H, W, stride=s, device=dev)
                sc = _ConvLayer(P[pre + "/sc/w"], P[pre + "/sc/b"], H, W, stride=s, device=dev) if ui == 0 else None
                H, W = H // s[0], W // s[1]
                c2 = _ConvLayer(P[pre + "/c2/w"], P[pre + "/c2/b"], H, W, device=dev)
                self.units.append((c1, c2, sc))
        assert H == 1

    def _encoder(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        """[B,64,256,3] fp32 -> [B,T,512] fp32."""
        B = x_nhwc.shape[0]
        x = F.avg_pool2d(x_nhwc.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)          # rectifier stand-in
        x = F.pad(x, (0, 64 - x.shape[3])).to(L.ACT_DTYPE).contiguous()           # channels 3 -> 64
        if FUSED_ENCODER and torch.is_grad_enabled():
            x = _EncoderFn.apply(x, self)
        else:
            x = self.stem(x)
            for c1, c2, sc in self.units:
                y = c1(x)
                shortcut = sc(x, relu=False) if sc is not None else x
                x = c2(y, residual=shortcut, relu=True)
        return x.reshape(B, x.shape[2], x.shape[3]).float()[:, :, :512]

    # -- recurrent part (plain batched GEMMs + element-wise gates) -------------------------------
    @staticmethod
    def _lstm_cell(x_proj, h, c, w_hh):
        gates = x_proj + h @ w_hh
        i, f, g, o = gates.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        return h, c

    def _bilstm(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """Both directions in one whole-sequence launch (``tbg_lstm_seq_fwd``): the input projections
        for all steps are one batched GEMM, the backward stream is time-reversed around the kernel."""
        lay = self.lstm[name]
        xp = torch.einsum("btk,dkn->dbtn", x, lay.w_ih) + lay.b                     # [2,B,T,1024]
        xp = torch.stack([xp[0], xp[1].flip(1)]).contiguous()
        hs = _LstmSeq.apply(xp, lay)                                                # [2,B,T,256]
        return torch.cat([hs[0], hs[1].flip(1)], dim=2)

    def _decoder(self, mem: torch.Tensor, steps: int) -> torch.Tensor:
        """All decode steps in one launch (``tbg_attn_decoder_fwd``); ``keys = mem @ W_m`` is a plain GEMM."""
        keys = mem @ self.P["dec/memory_layer/w"]
        return _AttnDecoder.apply(mem.contiguous(), keys.contiguous(), self.dec_w, steps)

    # -- reference surface -------------------------------------------------------------------------
    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        """aster_inferer.py:28-37 (batched instead of a batch-1 loop) + _postprocess_simple."""
        mem = self._encoder(inputs)
        mem = self._bilstm(mem, "rnn/l0")
        mem = self._bilstm(mem, "rnn/l1")
        logits = self._decoder(mem, self.cfg.max_char_number)
        return self._postprocess_simple(logits, self.cfg.max_char_number)

    @staticmethod
    def _postprocess_simple(logits: torch.Tensor, max_char_number: int) -> torch.Tensor:
        """aster_inferer.py:116-151"""
        logits = logits[:, :max_char_number]
        padding_len = max_char_number - logits.shape[1]
        if padding_len > 0:
            pad = logits.new_zeros(logits.shape[0], padding_len, logits.shape[2])
            pad[:, :, 1] = 1000.0
            logits = torch.cat([logits, pad], dim=1)
        return logits

    @staticmethod
    def convert_inputs(fake_images: torch.Tensor, labels: torch.Tensor, blank_label: int,
                       cfg: Optional[Config] = None) -> torch.Tensor:
        """aster_inferer.py:153-190 — NCHW -> NHWC, crop each image at the first blank label and
        bilinear-resize (half-pixel centres, no antialias) to cfg.aster_image_dims.  Batched: the
        per-sample crop is folded into a per-sample horizontal interpolation matrix, so the whole
        op is two small batched GEMMs (differentiable to any order)."""
        if cfg is None:
            from .config import cfg as _cfg
            cfg = _cfg
        B, Cc, H, W = fake_images.shape
        oh, ow = cfg.aster_image_dims
        dev = fake_images.device
        if L.use_fused() and FUSED_CONVERT_INPUTS and Cc == 3:
            return _CropResize.apply(fake_images, labels, blank_label, cfg.char_width, (oh, ow))
        is_blank = labels == blank_label
        has_blank = is_blank.any(dim=1)
        first = torch.where(has_blank, is_blank.int().argmax(dim=1), torch.full_like(labels[:, 0], 10 ** 6))
        from .utils import crop_width

        w_crop = torch.clamp(crop_width(first, cfg.char_width), min=1, max=W)       # [B]

        def interp_matrix(out_size: int, in_size: torch.Tensor, in_max: int) -> torch.Tensor:
            # rows: output index; tf.image.resize(bilinear): src = (o + 0.5) * in/out - 0.5
            o = torch.arange(out_size, device=dev, dtype=torch.float32)[None, :]
            scale = in_size.float()[:, None] / out_size
            src = (o + 0.5) * scale - 0.5
            f0 = torch.floor(src)
            lerp = src - f0
            hi = (in_size[:, None] - 1)
            i0 = torch.clamp(f0.long(), min=0)
            i0 = torch.minimum(i0, hi)
            i1 = torch.minimum(torch.clamp(torch.ceil(src).long(), min=0), hi)
            m = torch.zeros(in_size.shape[0], out_size, in_max, device=dev)
            m.scatter_add_(2, i0[..., None], (1.0 - lerp)[..., None])
            m.scatter_add_(2, i1[..., None], lerp[..., None])
            return m

        my = interp_matrix(oh, torch.full((1,), H, device=dev, dtype=torch.long), H)[0]      # [oh, H]
        mx = interp_matrix(ow, w_crop, W)                                                     # [B, ow, W]
        x = torch.einsum("yh,bchw->bcyw", my, fake_images.float())
        x = torch.einsum("bcyw,bxw->byxc", x, mx)
        return x
