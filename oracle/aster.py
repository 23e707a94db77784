"""ASTER recogniser stand-in and the reference's wrapper logic around it.
TEST INFRASTRUCTURE — see oracle/__init__.py.

**PARITY UNPINNED.**  The reference loads ASTER as an external SavedModel
(aster_ocr_utils/aster_inferer.py:24-26) whose weights and graph are not in the repository
(aster_weights/ holds only ``.keep``; README.md:60-66).  What IS restated from the reference:
``convert_inputs`` (aster_inferer.py:153-190), ``_postprocess_simple`` (:116-151), the batching
contract of ``call`` (:28-37) and the OCR losses (models/losses/ocr_losses.py:8-20).

The network itself follows the published ASTER recognition topology (Shi et al., TPAMI 2018;
hints in aster_ocr_utils/weigths_tf1_to_tf2.py:3-19: Bahdanau attention with
query_layer/memory_layer/attention_v, an LSTM-cell decoder and a dense output layer), with
frozen seeded synthetic weights, batch-norm folded into the convolutions, and three documented
simplifications: (1) the TPS rectifier is replaced by its identity sampling grid realised as a
2x2 average pool 64x256 -> 32x128 (keeps every spatial size a power of two; the paper samples
32x100); (2) greedy decoding runs for exactly ``max_char_number`` steps (the reference slices the
logits to that many steps anyway, :131); (3) the previous-symbol feedback is an embedding of the
arg-max (non-differentiable, as in the exported inference graph).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

NUM_CLASSES = 96          # ids 0 (GO) / 1 (EOS == blank/pad, char_tokens.py:16-17) / 2..95 chars
ENC_BLOCKS: List[Tuple[int, int, Tuple[int, int]]] = [
    # (channels, units, stride of the first unit)      ASTER paper Table 1
    (32, 3, (2, 2)),
    (64, 4, (2, 2)),
    (128, 6, (2, 1)),
    (256, 6, (2, 1)),
    (512, 3, (2, 1)),
]
LSTM_HIDDEN = 256
ATT_UNITS = 256
EMB_DIM = 256


def init_aster_params(seed: int = 1234, dtype=torch.float32) -> Params:
    """Seeded synthetic frozen weights (He-scaled so activations stay O(1) through 45 layers)."""
    g = torch.Generator().manual_seed(seed)
    P: Params = {}

    def conv(name, k, cin, cout, gain=2.0):
        std = math.sqrt(gain / (k * k * cin))
        P[name + "/w"] = (torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * std).to(dtype)
        P[name + "/b"] = (torch.randn(cout, generator=g, dtype=torch.float64) * 0.01).to(dtype)

    conv("enc/stem", 3, 3, 32)
    cin = 32
    for bi, (ch, units, stride) in enumerate(ENC_BLOCKS):
        for ui in range(units):
            pre = f"enc/b{bi}/u{ui}"
            conv(pre + "/c1", 1, cin, ch)
            conv(pre + "/c2", 3, ch, ch, gain=1.0)   # residual branch scaled down
            if ui == 0:
                conv(pre + "/sc", 1, cin, ch, gain=1.0)
            cin = ch

    def lstm(name, in_dim, hid):
        s = 1.0 / math.sqrt(hid)
        P[name + "/w_ih"] = ((torch.rand(in_dim, 4 * hid, generator=g, dtype=torch.float64) * 2 - 1) * s).to(dtype)
        P[name + "/w_hh"] = ((torch.rand(hid, 4 * hid, generator=g, dtype=torch.float64) * 2 - 1) * s).to(dtype)
        P[name + "/b"] = torch.zeros(4 * hid, dtype=dtype)

    for li in range(2):
        for d in ("fw", "bw"):
            lstm(f"rnn/l{li}/{d}", 512, LSTM_HIDDEN)

    def lin(name, i, o, bias=True):
        s = 1.0 / math.sqrt(i)
        P[name + "/w"] = ((torch.rand(i, o, generator=g, dtype=torch.float64) * 2 - 1) * s).to(dtype)
        if bias:
            P[name + "/b"] = torch.zeros(o, dtype=dtype)

    lin("dec/memory_layer", 512, ATT_UNITS, bias=False)
    lin("dec/query_layer", LSTM_HIDDEN, ATT_UNITS, bias=False)
    P["dec/attention_v"] = ((torch.rand(ATT_UNITS, generator=g, dtype=torch.float64) * 2 - 1)
                            / math.sqrt(ATT_UNITS)).to(dtype)
    P["dec/embedding"] = (torch.randn(NUM_CLASSES, EMB_DIM, generator=g, dtype=torch.float64) * 0.1).to(dtype)
    lstm("dec/lstm_cell", EMB_DIM + 512, LSTM_HIDDEN)
    lin("dec/dense", LSTM_HIDDEN + 512, NUM_CLASSES)
    return P


def _conv(x, P, name, stride=(1, 1)):
    w = P[name + "/w"]
    k = w.shape[0]
    return F.conv2d(x, w.permute(3, 2, 0, 1), P[name + "/b"], stride=stride, padding=k // 2)


def encoder(x: torch.Tensor, P: Params) -> torch.Tensor:
    """x: [B,3,32,128] -> [B, T=32, 512]."""
    x = torch.relu(_conv(x, P, "enc/stem"))
    for bi, (ch, units, stride) in enumerate(ENC_BLOCKS):
        for ui in range(units):
            pre = f"enc/b{bi}/u{ui}"
            s = stride if ui == 0 else (1, 1)
            xin = x[:, :, :: s[0], :: s[1]]                      # stride on the 1x1 = subsample first
            y = torch.relu(_conv(xin, P, pre + "/c1"))
            y = _conv(y, P, pre + "/c2")
            sc = _conv(xin, P, pre + "/sc") if ui == 0 else x
            x = torch.relu(y + sc)
    assert x.shape[2] == 1
    return x[:, :, 0, :].permute(0, 2, 1)


def lstm_cell(x_proj, h, c, w_hh):
    """gates order i, f, g, o (PyTorch/Keras 'ifgo' convention is immaterial for synthetic weights)."""
    gates = x_proj + h @ w_hh
    i, f, g, o = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def bilstm(x: torch.Tensor, P: Params, name: str) -> torch.Tensor:
    B, T, _ = x.shape
    outs = []
    for d in ("fw", "bw"):
        xp = x @ P[f"{name}/{d}/w_ih"] + P[f"{name}/{d}/b"]
        h = x.new_zeros(B, LSTM_HIDDEN)
        c = x.new_zeros(B, LSTM_HIDDEN)
        hs = [None] * T
        order = range(T) if d == "fw" else range(T - 1, -1, -1)
        for t in order:
            h, c = lstm_cell(xp[:, t], h, c, P[f"{name}/{d}/w_hh"])
            hs[t] = h
        outs.append(torch.stack(hs, dim=1))
    return torch.cat(outs, dim=2)


def attention_decoder(mem: torch.Tensor, P: Params, steps: int) -> torch.Tensor:
    """Greedy Bahdanau-attention LSTM decoder: mem [B,T,512] -> logits [B, steps, C]."""
    B = mem.shape[0]
    keys = mem @ P["dec/memory_layer/w"]
    h = mem.new_zeros(B, LSTM_HIDDEN)
    c = mem.new_zeros(B, LSTM_HIDDEN)
    prev = torch.zeros(B, dtype=torch.long)                       # GO symbol
    logits = []
    for _ in range(steps):
        q = h @ P["dec/query_layer/w"]
        e = torch.tanh(keys + q[:, None, :]) @ P["dec/attention_v"]
        a = torch.softmax(e, dim=1)
        ctx = (a[:, :, None] * mem).sum(dim=1)
        inp = torch.cat([P["dec/embedding"][prev], ctx], dim=1)
        xp = inp @ P["dec/lstm_cell/w_ih"] + P["dec/lstm_cell/b"]
        h, c = lstm_cell(xp, h, c, P["dec/lstm_cell/w_hh"])
        lg = torch.cat([h, ctx], dim=1) @ P["dec/dense/w"] + P["dec/dense/b"]
        logits.append(lg)
        prev = lg.detach().argmax(dim=1)
    return torch.stack(logits, dim=1)


def aster_forward_logits(images_nhwc: torch.Tensor, P: Params, steps: int) -> torch.Tensor:
    """The SavedModel's ``forward_logits`` for a batch: [B,64,256,3] in [-1,1] -> [B, steps, C]."""
    x = images_nhwc.permute(0, 3, 1, 2)
    x = F.avg_pool2d(x, 2)                                        # rectifier stand-in (see header)
    mem = encoder(x, P)
    mem = bilstm(mem, P, "rnn/l0")
    mem = bilstm(mem, P, "rnn/l1")
    return attention_decoder(mem, P, steps)


def postprocess_simple(logits: torch.Tensor, max_char_number: int) -> torch.Tensor:
    """aster_inferer.py:116-151 — keep ``max_char_number`` steps; pad missing steps with
    1000 * onehot(1)."""
    logits = logits[:, :max_char_number]
    padding_len = max_char_number - logits.shape[1]
    if padding_len > 0:
        pad = torch.zeros(logits.shape[0], padding_len, logits.shape[2], dtype=logits.dtype)
        pad[:, :, 1] = 1000.0
        logits = torch.cat([logits, pad], dim=1)
    return logits


def aster_inferer_call(images_nhwc: torch.Tensor, P: Params, cfg) -> torch.Tensor:
    """AsterInferer.call (aster_inferer.py:28-37) with combine_forward_and_backward=False.  The
    reference loops over the batch one image at a time; every op of the stand-in is per-sample, so
    the batched evaluation is identical."""
    return postprocess_simple(aster_forward_logits(images_nhwc, P, cfg.max_char_number), cfg.max_char_number)


def convert_inputs(fake_images: torch.Tensor, labels: torch.Tensor, blank_label: int, cfg) -> torch.Tensor:
    """aster_inferer.py:153-190 — NCHW -> NHWC, crop at the first blank label, bilinear resize
    (tf.image.resize default: half-pixel centres, no antialias) to cfg.aster_image_dims."""
    out = []
    oh, ow = cfg.aster_image_dims
    for b in range(fake_images.shape[0]):
        img = fake_images[b: b + 1]
        idx = (labels[b] == blank_label).nonzero()
        if idx.numel() > 0:
            w_crop = int(int(idx[0, 0]) * cfg.char_width)     # floor for the fractional-width extension (see mask_text_box)
            img = img[:, :, :, :w_crop]
        out.append(F.interpolate(img, size=(oh, ow), mode="bilinear", align_corners=False, antialias=False))
    return torch.cat(out, dim=0).permute(0, 2, 3, 1)


def softmax_cross_entropy_loss(y_pred: torch.Tensor, y_true: torch.Tensor, batch_size: int) -> torch.Tensor:
    """ocr_losses.py:8-11"""
    loss = F.cross_entropy(y_pred.reshape(-1, y_pred.shape[-1]), y_true.reshape(-1).long(), reduction="none")
    return loss.sum() / batch_size


def mean_squared_loss(y_with_noise: torch.Tensor, y_without_noise: torch.Tensor, batch_size: int) -> torch.Tensor:
    """ocr_losses.py:14-20 — keras mse = mean over the last axis."""
    loss = ((y_with_noise - y_without_noise) ** 2).mean(dim=-1)
    return loss.sum() / batch_size
