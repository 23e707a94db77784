"""StyleGAN2 layers, word encoder, generator and discriminator of TextBoxGAN restated in
PyTorch-CPU (NCHW, dtype of the inputs: fp32 for parity, fp64 for gradcheck).
TEST INFRASTRUCTURE — see oracle/__init__.py.  All citations are into /root/reference.

Parameters live in a flat ``dict[str, Tensor]`` whose keys mirror the reference's Keras layer
names (e.g. ``synthesis/4x16/block/conv_0/w``); shapes and initial distributions follow the
reference ``build`` methods.  Every random draw is passed in.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# commons.py:4-12
# ----------------------------------------------------------------------------------------------
def compute_runtime_coef(weight_shape: Sequence[int], gain: float, lrmul: float) -> Tuple[float, float]:
    fan_in = float(np.prod(weight_shape[:-1]))
    he_std = gain / math.sqrt(fan_in)
    init_std = 1.0 / lrmul
    runtime_coef = he_std * lrmul
    return init_std, runtime_coef


# ----------------------------------------------------------------------------------------------
# upfirdn_2d_v2.py
# ----------------------------------------------------------------------------------------------
def _setup_kernel(k) -> np.ndarray:
    """upfirdn_2d_v2.py:18-25"""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2
    assert k.shape[0] == k.shape[1]
    return k


def compute_paddings(resample_kernel, up: bool, down: bool, is_conv: bool, convW: int = 3, factor: int = 2,
                     gain: float = 1):
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


def upfirdn_2d_ref(x: torch.Tensor, k, upx, upy, downx, downy, padx0, padx1, pady0, pady1) -> torch.Tensor:
    """upfirdn_2d_v2.py:249-305, literally: x is [major, inH, inW, minor]."""
    k = np.asarray(k, dtype=np.float32)
    _, inH, inW, minorDim = x.shape
    kernelH, kernelW = k.shape
    assert inW >= 1 and inH >= 1 and kernelW >= 1 and kernelH >= 1
    # Upsample (insert zeros).
    x = x.reshape(-1, inH, 1, inW, 1, minorDim)
    x = F.pad(x, (0, 0, 0, upx - 1, 0, 0, 0, upy - 1))
    x = x.reshape(-1, inH * upy, inW * upx, minorDim)
    # Pad (crop if negative).
    x = F.pad(x, (0, 0, max(padx0, 0), max(padx1, 0), max(pady0, 0), max(pady1, 0)))
    x = x[:, max(-pady0, 0): x.shape[1] - max(-pady1, 0), max(-padx0, 0): x.shape[2] - max(-padx1, 0), :]
    # Convolve with filter.
    x = x.permute(0, 3, 1, 2)
    x = x.reshape(-1, 1, inH * upy + pady0 + pady1, inW * upx + padx0 + padx1)
    w = torch.as_tensor(np.ascontiguousarray(k[::-1, ::-1]), dtype=x.dtype)[None, None]
    x = F.conv2d(x, w)  # tf.nn.conv2d == cross-correlation, VALID
    x = x.reshape(-1, minorDim, inH * upy + pady0 + pady1 - kernelH + 1, inW * upx + padx0 + padx1 - kernelW + 1)
    x = x.permute(0, 2, 3, 1)
    # Downsample (throw away pixels).
    return x[:, ::downy, ::downx, :]


def _simple_upfirdn_2d(x: torch.Tensor, x_res_h: int, x_res_w: int, k, up_x=1, up_y=1, down=1, pad0=0, pad1=0):
    """upfirdn_2d_v2.py:166-183 (x is NCHW)."""
    assert x.dim() == 4
    c = x.shape[1]
    y = x.reshape(-1, x_res_h, x_res_w, 1)
    y = upfirdn_2d_ref(y, k, upx=up_x, upy=up_y, downx=down, downy=down, padx0=pad0, padx1=pad1, pady0=pad0,
                       pady1=pad1)
    return y.reshape(-1, c, y.shape[1], y.shape[2])


def upsample_2d(x, res_h, res_w, pad0, pad1, k, factor=2):
    """upfirdn_2d_v2.py:58-62"""
    return _simple_upfirdn_2d(x, res_h, res_w, k, up_x=factor, up_y=factor, pad0=pad0, pad1=pad1)


def upsample_conv_2d(x, w_res, h_res, w, pad0, pad1, k, factor=2):
    """upfirdn_2d_v2.py:65-103.  ``w`` is [kh, kw, inC, num_groups*outC] (TF HWIO, grouped by
    sample in the fused path), ``x`` is [N, num_groups*inC, h, w]."""
    convH, convW, inC = w.shape[0], w.shape[1], w.shape[2]
    num_groups = x.shape[1] // inC
    # Transpose weights (:78-80).
    w = w.reshape(convH, convW, inC, num_groups, -1)
    w = torch.flip(w, dims=(0, 1)).permute(0, 1, 4, 3, 2)  # [kh, kw, outC, groups, inC]
    outC = w.shape[2]
    # TF conv2d_transpose filter [kh, kw, out_channels, in_channels] with in = groups*inC.
    output_height = (h_res - 1) * factor + convH
    output_width = (w_res - 1) * factor + convW
    # PyTorch conv_transpose2d weight: [in_channels, out_channels/groups, kh, kw]
    wt = w.permute(3, 4, 2, 0, 1).reshape(num_groups * inC, outC, convH, convW)
    # TF's conv2d_transpose is the gradient of a correlation => it scatters x[i]*F[a] to 2i+a,
    # which is exactly what torch.conv_transpose2d does with the same (un-flipped) filter.
    y = F.conv_transpose2d(x, wt, stride=factor, groups=num_groups)
    assert y.shape[2] == output_height and y.shape[3] == output_width
    return _simple_upfirdn_2d(y, output_height, output_width, k, pad0=pad0, pad1=pad1)


def conv_downsample_2d(x, w_res, h_res, w, pad0, pad1, k, reduce_height: bool):
    """upfirdn_2d_v2.py:106-113 (note the swapped argument names at the call sites, SURVEY A.4:
    callers pass (in_h_res, in_w_res) and this forwards them positionally to (x_res_h, x_res_w))."""
    h_stride = 2 if reduce_height else 1
    w_stride = 2
    x = _simple_upfirdn_2d(x, w_res, h_res, k, pad0=pad0, pad1=pad1)
    wt = w.permute(3, 2, 0, 1)  # HWIO -> OIHW
    return F.conv2d(x, wt, stride=(h_stride, w_stride))


# ----------------------------------------------------------------------------------------------
# dense.py / bias_act.py / noise.py
# ----------------------------------------------------------------------------------------------
def dense(x: torch.Tensor, w: torch.Tensor, gain: float = 1.0, lrmul: float = 1.0) -> torch.Tensor:
    """dense.py:13-29"""
    _, coef = compute_runtime_coef(list(w.shape), gain, lrmul)
    return x.reshape(x.shape[0], -1) @ (coef * w)


def bias_act(x: torch.Tensor, b: torch.Tensor, lrmul: float, act: str) -> torch.Tensor:
    """bias_act.py:13-34"""
    assert act in ("linear", "lrelu")
    bb = lrmul * b
    x = x + (bb if x.dim() == 2 else bb.reshape(1, -1, 1, 1))
    if act == "lrelu":
        x = F.leaky_relu(x, 0.2) * math.sqrt(2)
    return x


def apply_noise(x: torch.Tensor, noise: torch.Tensor, strength: torch.Tensor) -> torch.Tensor:
    """noise.py:12-22 — noise is [B,1,H,W]."""
    return x + noise * strength


# ----------------------------------------------------------------------------------------------
# modulated_conv2d.py:66-122
# ----------------------------------------------------------------------------------------------
def modulated_conv2d(x: torch.Tensor, y: torch.Tensor, P: Params, prefix: str, *, up: bool, demodulate: bool,
                     fused: bool, in_h_res: Optional[int] = None, in_w_res: Optional[int] = None,
                     ret_aux: bool = False):
    w_raw = P[prefix + "/w"]  # [k,k,I,O]
    kh, kw, I, O = w_raw.shape
    _, coef = compute_runtime_coef([kh, kw, I, O], 1.0, 1.0)
    k, pad0, pad1 = compute_paddings([1, 3, 3, 1] if up else None, up, False, is_conv=True)
    w = coef * w_raw                                             # :71
    ww = w[None]                                                 # :72
    s = dense(y, P[prefix + "/mod_dense/w"])                     # :75
    s = bias_act(s, P[prefix + "/mod_bias/b"], 1.0, "linear") + 1.0   # :76
    ww = ww * s[:, None, None, :, None]                          # :77
    d = None
    if demodulate:
        d = torch.rsqrt(torch.sum(ww * ww, dim=(1, 2, 3)) + 1e-8)  # :80-82
        ww = ww * d[:, None, None, None, :]                        # :83
    B = x.shape[0]
    if fused:
        xs = x.reshape(1, -1, x.shape[2], x.shape[3])            # :88
        wg = ww.permute(1, 2, 3, 0, 4).reshape(kh, kw, I, -1)    # :89-92  [k,k,I,B*O]
    else:
        xs = x * s[:, :, None, None]                             # :96
        wg = w
    if up:
        out = upsample_conv_2d(xs, in_w_res, in_h_res, wg, pad0, pad1, k)   # :99-108
    else:
        groups = B if fused else 1
        wt = wg.reshape(kh, kw, I, groups, -1).permute(3, 4, 2, 0, 1).reshape(-1, I, kh, kw)
        out = F.conv2d(xs, wt, padding=(kh // 2, kw // 2), groups=groups)   # SAME, stride 1 (:110-112)
    if fused:
        out = out.reshape(-1, O, out.shape[2], out.shape[3])    # :115-118
    elif demodulate:
        out = out * d[:, :, None, None]                          # :121
    if ret_aux:
        return out, s, d
    return out


def to_rgb(x, style, P: Params, prefix: str, fused: bool = True):
    """to_rgb.py:28-33"""
    y = modulated_conv2d(x, style, P, prefix + "/conv", up=False, demodulate=False, fused=fused)
    return bias_act(y, P[prefix + "/bias/b"], 1.0, "linear")


# ----------------------------------------------------------------------------------------------
# synthesis_block.py
# ----------------------------------------------------------------------------------------------
def synthesis_block(x, w0, w1, P: Params, prefix: str, out_h: int, out_w: int, noises, fused: bool = True):
    """synthesis_block.py:62-74"""
    x = modulated_conv2d(x, w0, P, prefix + "/conv_0", up=True, demodulate=True, fused=fused,
                         in_h_res=out_h // 2, in_w_res=out_w // 2)
    x = apply_noise(x, noises[0], P[prefix + "/noise_0/w"])
    x = bias_act(x, P[prefix + "/bias_0/b"], 1.0, "lrelu")
    x = modulated_conv2d(x, w1, P, prefix + "/conv_1", up=False, demodulate=True, fused=fused,
                         in_h_res=out_h, in_w_res=out_w)
    x = apply_noise(x, noises[1], P[prefix + "/noise_1/w"])
    x = bias_act(x, P[prefix + "/bias_1/b"], 1.0, "lrelu")
    return x


def synthesis(x, style, P: Params, cfg, noises: List[torch.Tensor], fused: bool = True, prefix: str = "synthesis"):
    """synthesis_block.py:137-156.  ``noises`` holds 2 tensors per block, [B,1,H_l,W_l]."""
    res = cfg.generator_resolutions
    k, pad0, pad1 = compute_paddings([1, 3, 3, 1], up=True, down=False, is_conv=False)
    y = to_rgb(x, style[:, 0], P, f"{prefix}/{res[0][0]}x{res[0][1]}/ToRGB", fused)          # :140
    for i, (h_res, w_res) in enumerate(res[1:]):
        idx = 3 * i                                                                         # :143
        s0, s1, s2 = style[:, idx], style[:, idx + 1], style[:, idx + 2]
        x = synthesis_block(x, s0, s1, P, f"{prefix}/{h_res}x{w_res}/block", h_res, w_res,
                            noises[2 * i: 2 * i + 2], fused)
        y = upsample_2d(y, h_res // 2, w_res // 2, pad0, pad1, k)                           # :152
        y = y + to_rgb(x, s2, P, f"{prefix}/{h_res}x{w_res}/ToRGB", fused)                  # :153
    return y


# ----------------------------------------------------------------------------------------------
# mapping_block.py / latent_encoder.py
# ----------------------------------------------------------------------------------------------
def mapping(z, P: Params, cfg, prefix: str = "latent_encoder/g_mapping"):
    """mapping_block.py:15-45"""
    x = z * torch.rsqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)
    for i in range(cfg.n_mapping):
        x = dense(x, P[f"{prefix}/dense_{i}/w"], gain=1.0, lrmul=0.01)
        x = bias_act(x, P[f"{prefix}/bias_{i}/b"], 0.01, "lrelu")
    return x


def lerp(a, b, t):
    """custom_stylegan2/utils.py:25-27"""
    return a + (b - a) * t


def latent_encoder(z, P: Params, cfg, n_broadcast: int, *, training: bool, truncation_psi: float = 1.0,
                   draws: Optional[dict] = None, state_out: Optional[dict] = None):
    """latent_encoder.py:80-99.  ``draws``: z2 [B,S], mix_coin (float in [0,1)), mix_cutoff (int in
    [1, n_broadcast)).  The new ``w_avg`` is returned through ``state_out['w_avg']`` (the reference
    assigns the non-trainable variable in place, :39-45)."""
    w = mapping(z, P, cfg)
    wb = w[:, None, :].expand(-1, n_broadcast, -1)                 # :24-26
    if training:
        batch_avg = wb[:, 0].mean(dim=0)                           # :41
        new_avg = lerp(batch_avg, P["latent_encoder/w_avg"], 0.995)  # :44
        if state_out is not None:
            state_out["w_avg"] = new_avg.detach()
        # style mixing :47-71
        w2 = mapping(draws["z2"], P, cfg)
        wb2 = w2[:, None, :].expand(-1, n_broadcast, -1)
        cutoff = int(draws["mix_cutoff"]) if float(draws["mix_coin"]) < 0.9 else n_broadcast
        layer_idx = torch.arange(n_broadcast)[None, :, None]
        wb = torch.where(layer_idx < cutoff, wb, wb2)
    if not training:
        wb = lerp(P["latent_encoder/w_avg"], wb, truncation_psi)   # :73-78
    return wb


# ----------------------------------------------------------------------------------------------
# word_encoder.py:39-63
# ----------------------------------------------------------------------------------------------
def word_encoder(words: torch.Tensor, P: Params, cfg, *, dropout_mask: Optional[torch.Tensor] = None,
                 prefix: str = "word_encoder"):
    """``dropout_mask`` is the Bernoulli(0.7) keep mask [B,mcn,32] (training) or None (inference);
    Keras Dropout scales kept values by 1/0.7."""
    emb_table = torch.cat([P[prefix + "/w0_embedding"], P[prefix + "/w_embedding"]], dim=0)  # :43
    emb = emb_table[words.long()]                                 # :44-46
    if dropout_mask is not None:
        emb = emb * dropout_mask / 0.7                            # :47
    B = words.shape[0]
    x = emb.reshape(B * cfg.max_char_number, cfg.embedding_out_dim)   # :49-51
    x = torch.relu(x @ P[prefix + "/fc/kernel"] + P[prefix + "/fc/bias"])   # :53
    out_h, out_w = cfg.generator_resolutions[0]
    out_c = cfg.generator_feat_maps[0]
    return x.reshape(B, out_w, out_c, out_h).permute(0, 2, 3, 1)  # :55-61


# ----------------------------------------------------------------------------------------------
# generator.py:19-43
# ----------------------------------------------------------------------------------------------
def generator(words, z, P: Params, cfg, *, training: bool, draws: dict, truncation_psi: float = 1.0,
              ret_style: bool = False, fused: bool = True, state_out: Optional[dict] = None):
    """``draws``: noises (list), and when training: dropout_mask, z2, mix_coin, mix_cutoff."""
    x = word_encoder(words, P, cfg, dropout_mask=draws.get("dropout_mask") if training else None)
    n_style = 3 * (len(cfg.generator_resolutions) - 1)            # :16
    style = latent_encoder(z, P, cfg, n_style, training=training, truncation_psi=truncation_psi, draws=draws,
                           state_out=state_out)
    img = synthesis(x, style, P, cfg, draws["noises"], fused)
    return (img, style) if ret_style else img


def synthesis_from_style(words, style, P: Params, cfg, noises, fused: bool = True):
    """generator.call with the latent encoder bypassed (used by the path-length regulariser)."""
    x = word_encoder(words, P, cfg, dropout_mask=None)
    return synthesis(x, style, P, cfg, noises, fused)


# ----------------------------------------------------------------------------------------------
# conv.py / from_rgb.py / mini_batch_std.py / discriminator.py
# ----------------------------------------------------------------------------------------------
def conv2d_layer(x, w_raw, *, down: bool, reduce_height: bool = False, in_h_res=None, in_w_res=None):
    """conv.py:51-73"""
    kh, kw, I, O = w_raw.shape
    _, coef = compute_runtime_coef([kh, kw, I, O], 1.0, 1.0)
    w = coef * w_raw
    if down:
        k, pad0, pad1 = compute_paddings([1, 3, 3, 1], False, True, is_conv=True, convW=kh)
        return conv_downsample_2d(x, in_h_res, in_w_res, w, pad0, pad1, k, reduce_height)
    return F.conv2d(x, w.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2))


def minibatch_std(x, group_size: int = 4, num_new_features: int = 1):
    """mini_batch_std.py:10-35"""
    B, C, H, W = x.shape
    g = min(group_size, B)
    y = x.reshape(g, -1, num_new_features, C // num_new_features, H, W)
    y = y - y.mean(dim=0, keepdim=True)
    y = (y * y).mean(dim=0)
    y = torch.sqrt(y + 1e-8)
    y = y.mean(dim=(2, 3, 4), keepdim=True)
    y = y.mean(dim=2)
    y = y.repeat(g, 1, H, W)
    return torch.cat([x, y], dim=1)


def discriminator(images, P: Params, cfg, prefix: str = ""):
    """discriminator.py:202-214 (+ blocks :68-84, :132-142)."""
    res, fm = cfg.discrim_resolutions, cfg.discrim_feat_maps
    r0 = res[0]
    p0 = f"{prefix}{r0[0]}x{r0[1]}/FromRGB"
    x = conv2d_layer(images, P[p0 + "/conv/w"], down=False)                      # from_rgb.py:26-29
    x = bias_act(x, P[p0 + "/bias/b"], 1.0, "lrelu")
    for (h, w), (nh, nw) in zip(res[:-1], res[1:]):
        pb = f"{prefix}{h}x{w}"
        residual = x
        x = conv2d_layer(x, P[pb + "/conv_0/w"], down=False)
        x = bias_act(x, P[pb + "/bias_0/b"], 1.0, "lrelu")
        x = conv2d_layer(x, P[pb + "/conv_1/w"], down=True, reduce_height=(h != nh), in_h_res=h, in_w_res=w)
        x = bias_act(x, P[pb + "/bias_1/b"], 1.0, "lrelu")
        residual = conv2d_layer(residual, P[pb + "/skip/w"], down=True, reduce_height=(h != nh), in_h_res=h,
                                in_w_res=w)
        x = (x + residual) * (1.0 / math.sqrt(2.0))                              # :82
    rf = res[-1]
    pl = f"{prefix}{rf[0]}x{rf[1]}/last"
    x = minibatch_std(x, 4, 1)
    x = conv2d_layer(x, P[pl + "/conv_0/w"], down=False)
    x = bias_act(x, P[pl + "/bias_0/b"], 1.0, "lrelu")
    x = dense(x, P[pl + "/dense_1/w"])                                           # flatten NCHW
    x = bias_act(x, P[pl + "/bias_1/b"], 1.0, "lrelu")
    x = dense(x, P[prefix + "last_dense/w"])
    x = bias_act(x, P[prefix + "last_bias/b"], 1.0, "linear")
    return x


# ----------------------------------------------------------------------------------------------
# parameter construction (reference build() methods)
# ----------------------------------------------------------------------------------------------
def init_generator_params(cfg, gen: torch.Generator, dtype=torch.float32) -> Params:
    P: Params = {}
    S = cfg.style_dim

    def randn(*shape, std=1.0):
        return (torch.randn(*shape, generator=gen, dtype=torch.float64) * std).to(dtype)

    # word_encoder.py:28-37, Keras Dense(256): glorot-uniform kernel, zero bias
    P["word_encoder/w_embedding"] = randn(len(_MAIN_VOCAB) - 1, cfg.embedding_out_dim)
    P["word_encoder/w0_embedding"] = torch.zeros(1, cfg.embedding_out_dim, dtype=dtype)
    lim = math.sqrt(6.0 / (cfg.embedding_out_dim + cfg.word_encoder_dense_dim))
    P["word_encoder/fc/kernel"] = ((torch.rand(cfg.embedding_out_dim, cfg.word_encoder_dense_dim, generator=gen,
                                               dtype=torch.float64) * 2 - 1) * lim).to(dtype)
    P["word_encoder/fc/bias"] = torch.zeros(cfg.word_encoder_dense_dim, dtype=dtype)
    # mapping (lrmul 0.01 => init std 100), mapping_block.py:13,24-33
    for i in range(cfg.n_mapping):
        in_dim = cfg.z_dim if i == 0 else S
        P[f"latent_encoder/g_mapping/dense_{i}/w"] = randn(in_dim, S, std=1.0 / 0.01)
        P[f"latent_encoder/g_mapping/bias_{i}/b"] = torch.zeros(S, dtype=dtype)
    P["latent_encoder/w_avg"] = torch.zeros(S, dtype=dtype)

    def modconv(prefix, k, I, O):
        P[prefix + "/w"] = randn(k, k, I, O)
        P[prefix + "/mod_dense/w"] = randn(S, I)
        P[prefix + "/mod_bias/b"] = torch.zeros(I, dtype=dtype)

    def torgb(prefix, C):
        modconv(prefix + "/conv", 1, C, 3)
        P[prefix + "/bias/b"] = torch.zeros(3, dtype=dtype)

    res, fm = cfg.generator_resolutions, cfg.generator_feat_maps
    torgb(f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB", fm[0])
    prev = fm[0]
    for (h, w), f in zip(res[1:], fm[1:]):
        pb = f"synthesis/{h}x{w}/block"
        modconv(pb + "/conv_0", 3, prev, f)
        P[pb + "/noise_0/w"] = torch.zeros((), dtype=dtype)
        P[pb + "/bias_0/b"] = torch.zeros(f, dtype=dtype)
        modconv(pb + "/conv_1", 3, f, f)
        P[pb + "/noise_1/w"] = torch.zeros((), dtype=dtype)
        P[pb + "/bias_1/b"] = torch.zeros(f, dtype=dtype)
        torgb(f"synthesis/{h}x{w}/ToRGB", f)
        prev = f
    return P


def init_discriminator_params(cfg, gen: torch.Generator, dtype=torch.float32) -> Params:
    P: Params = {}

    def randn(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float64).to(dtype)

    res, fm = cfg.discrim_resolutions, cfg.discrim_feat_maps
    r0 = res[0]
    P[f"{r0[0]}x{r0[1]}/FromRGB/conv/w"] = randn(1, 1, 3, fm[0])
    P[f"{r0[0]}x{r0[1]}/FromRGB/bias/b"] = torch.zeros(fm[0], dtype=dtype)
    for (h, w), f0, f1 in zip(res[:-1], fm[:-1], fm[1:]):
        pb = f"{h}x{w}"
        P[pb + "/conv_0/w"] = randn(3, 3, f0, f0)
        P[pb + "/bias_0/b"] = torch.zeros(f0, dtype=dtype)
        P[pb + "/conv_1/w"] = randn(3, 3, f0, f1)
        P[pb + "/bias_1/b"] = torch.zeros(f1, dtype=dtype)
        P[pb + "/skip/w"] = randn(1, 1, f0, f1)
    rf = res[-1]
    n_f0, n_f1 = fm[-2], fm[-1]   # discriminator.py:193
    pl = f"{rf[0]}x{rf[1]}/last"
    P[pl + "/conv_0/w"] = randn(3, 3, n_f0 + 1, n_f0)
    P[pl + "/bias_0/b"] = torch.zeros(n_f0, dtype=dtype)
    P[pl + "/dense_1/w"] = randn(n_f0 * rf[0] * rf[1], n_f1)
    P[pl + "/bias_1/b"] = torch.zeros(n_f1, dtype=dtype)
    P["last_dense/w"] = randn(n_f1, 1)
    P["last_bias/b"] = torch.zeros(1, dtype=dtype)
    return P


from .tokens import MAIN_WORD_INDEX as _MAIN_VOCAB  # noqa: E402  (70 entries incl. <OOV>)

NON_TRAINABLE = ("word_encoder/w0_embedding", "latent_encoder/w_avg")


def trainable_names(P: Params, scopes: Sequence[str]) -> List[str]:
    """Names under any of ``scopes`` excluding non-trainable variables, in insertion order
    (model.trainable_variables order is irrelevant to the arithmetic)."""
    return [n for n in P if any(n.startswith(s) for s in scopes) and n not in NON_TRAINABLE]
