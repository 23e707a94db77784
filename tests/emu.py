"""CPU emulation of the two tensor-core entry points (TEST INFRASTRUCTURE).

Lets the host-side logic (geometry algebra, weight re-layouts, autograd wiring, model code) be
checked against the oracle on a machine without a GPU.  It implements the *documented semantics*
of ``tbg_conv2d_igemm`` / ``tbg_conv2d_wgrad`` (include/tbg.h) with dense torch ops; the product
never imports it."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


def _gather_tap(xp, PH, PW, ty, tx, pad, stride, Ho, Wo):
    h0 = PH + ty - pad[0]
    w0 = PW + tx - pad[1]
    return xp[:, h0: h0 + (Ho - 1) * stride[0] + 1: stride[0], w0: w0 + (Wo - 1) * stride[1] + 1: stride[1], :]


def emu_conv2d_igemm(x, w, *, Ho, Wo, taps, pad, stride=(1, 1), up=(0, 0), col_scale=None, bias=None, noise=None,
                     noise_strength=None, residual=None, res_scale=1.0, res_first=False, act=0, act_gain=1.0, out_fp32=False, out=None,
                     tap_mask=None, relu_mask=None):
    if isinstance(up, bool):
        up = (int(up), int(up))
    B, H, W_, Cin = x.shape
    ph, pw = 1 + up[0], 1 + up[1]
    n_total = w.shape[0]
    cout = n_total // (ph * pw)
    w6 = w.reshape(ph, pw, cout, taps[0], taps[1], Cin).to(torch.float64)
    if tap_mask is not None:      # masked taps are not computed (their weight blocks are treated as zero)
        keep = torch.zeros(ph, pw, 1, taps[0], taps[1], 1, dtype=torch.float64)
        for p_ in range(ph):
            for q_ in range(pw):
                m = int(tap_mask[p_ * pw + q_])
                for t_ in range(taps[0] * taps[1]):
                    if m == 0 or (m >> t_) & 1:
                        keep[p_, q_, 0, t_ // taps[1], t_ % taps[1], 0] = 1.0
        w6 = w6 * keep
    PH = taps[0] + pad[0] + stride[0] * Ho + 2
    PW = taps[1] + pad[1] + stride[1] * Wo + 2
    xp = F.pad(x.to(torch.float64), (0, 0, PW, PW, PH, PH))
    acc = torch.zeros(B, Ho, Wo, ph, pw, cout, dtype=torch.float64)
    for ty in range(taps[0]):
        for tx in range(taps[1]):
            xs = _gather_tap(xp, PH, PW, ty, tx, pad, stride, Ho, Wo)
            acc += torch.einsum("bhwc,pqoc->bhwpqo", xs, w6[:, :, :, ty, tx, :])
    y = acc.permute(0, 1, 3, 2, 4, 5).reshape(B, Ho * ph, Wo * pw, cout)
    if col_scale is not None:
        y = y * col_scale[:, None, None, :].double()
    if noise is not None:
        y = y + noise[..., None].double() * noise_strength.double().reshape(())
    if bias is not None:
        y = y + bias.double()
    if residual is not None and res_first:
        y = (y + residual.double()) * res_scale
    if act == 1:
        y = F.leaky_relu(y, 0.2)
    elif act == 2:
        y = torch.relu(y)
    y = y * act_gain
    if residual is not None and not res_first:
        y = (y + residual.double()) * res_scale
    if relu_mask is not None:
        y = torch.where(relu_mask.double() > 0, y, torch.zeros_like(y))
    y = y.to(torch.float32 if out_fp32 else x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def emu_conv2d_wgrad(x, gy, *, Ho, Wo, taps, pad, stride=(1, 1), up=(0, 0), gw=None):
    if isinstance(up, bool):
        up = (int(up), int(up))
    B, H, W_, Cin = x.shape
    ph, pw = 1 + up[0], 1 + up[1]
    cout = gy.shape[3]
    PH = taps[0] + pad[0] + stride[0] * Ho + 2
    PW = taps[1] + pad[1] + stride[1] * Wo + 2
    xp = F.pad(x.to(torch.float64), (0, 0, PW, PW, PH, PH))
    g6 = gy.to(torch.float64).reshape(B, Ho, ph, Wo, pw, cout)
    out = torch.zeros(ph, pw, cout, taps[0], taps[1], Cin, dtype=torch.float64)
    for ty in range(taps[0]):
        for tx in range(taps[1]):
            xs = _gather_tap(xp, PH, PW, ty, tx, pad, stride, Ho, Wo)
            out[:, :, :, ty, tx, :] = torch.einsum("bhwc,bhpwqo->pqoc", xs, g6)
    res = out.reshape(ph * pw * cout, taps[0] * taps[1] * Cin).to(torch.float32 if x.dtype != torch.float64 else torch.float64)
    if gw is not None:
        gw += res
        return gw
    return res


def emu_upfirdn2d(x, k, *, upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0, pady1=0):
    """Documented semantics of tbg_upfirdn2d == upfirdn_2d_ref (upfirdn_2d_v2.py:249-305)."""
    from oracle.stylegan import upfirdn_2d_ref

    return upfirdn_2d_ref(x, k.cpu().numpy(), upx, upy, downx, downy, padx0, padx1, pady0, pady1).contiguous()


def emu_adam_step(p, g, m, v, lr_t, beta1, beta2, eps):
    if torch.is_tensor(lr_t):
        lr_t = float(lr_t)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    p.sub_(lr_t * m / (v.sqrt() + eps))


def emu_ema_step(dst, src, beta):
    dst.copy_(src + (dst - src) * beta)


def emu_lstm_seq_fwd(xp, w_packed):
    """Documented semantics of tbg_lstm_seq_fwd (gate order i,f,g,o; w_packed[d,k,j,g])."""
    D, B, T, H4 = xp.shape
    H = H4 // 4
    w_hh = w_packed.double().permute(0, 1, 3, 2).reshape(D, H, 4 * H)            # [d,k,(g,j)]
    h = torch.zeros(D, B, H, dtype=torch.float64)
    c = torch.zeros(D, B, H, dtype=torch.float64)
    hs, gs, cs = [], [], []
    for t in range(T):
        gates = xp[:, :, t].double() + torch.bmm(h, w_hh)
        i, f, g, o = gates.chunk(4, dim=-1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h)
        cs.append(c)
        gs.append(torch.cat([i, f, g, o], dim=-1))
    return (torch.stack(hs, 2).to(xp.dtype), torch.stack(gs, 2).to(xp.dtype), torch.stack(cs, 2).to(xp.dtype))


def emu_lstm_seq_bwd(g_h, gates, c, wT_packed):
    D, B, T, H = g_h.shape
    w_hh = wT_packed.double().permute(0, 2, 3, 1).reshape(D, H, 4 * H)            # [d,k,(g,j)]
    dh_rec = torch.zeros(D, B, H, dtype=torch.float64)
    dc_next = torch.zeros(D, B, H, dtype=torch.float64)
    out = torch.zeros(D, B, T, 4 * H, dtype=torch.float64)
    for t in range(T - 1, -1, -1):
        i, f, g, o = gates[:, :, t].double().chunk(4, dim=-1)
        cc = c[:, :, t].double()
        cp = c[:, :, t - 1].double() if t > 0 else torch.zeros_like(cc)
        dh = g_h[:, :, t].double() + dh_rec
        tc = torch.tanh(cc)
        do = dh * tc * o * (1 - o)
        dc = dh * o * (1 - tc * tc) + dc_next
        di = dc * g * i * (1 - i)
        df = dc * cp * f * (1 - f)
        dg = dc * i * (1 - g * g)
        dc_next = dc * f
        dgates = torch.cat([di, df, dg, do], dim=-1)
        out[:, :, t] = dgates
        dh_rec = torch.bmm(dgates, w_hh.transpose(1, 2))
    return out.to(g_h.dtype)


def emu_modulate(x, s):
    return (x.double() * s.double().reshape(s.shape[0], *([1] * (x.dim() - 2)), s.shape[1])).to(x.dtype)


def emu_modulate_bwd(gxs, x, s, gs_init=None):
    sb = s.double().reshape(s.shape[0], *([1] * (x.dim() - 2)), s.shape[1])
    gx = (gxs.double() * sb).to(x.dtype)
    gs = (gxs.double() * x.double()).reshape(x.shape[0], -1, x.shape[-1]).sum(1).to(s.dtype)
    if gs_init is not None:
        gs = gs + gs_init
    return gx, gs


def _emu_interp(out_size, in_size, in_max):
    o = torch.arange(out_size, dtype=torch.float64)[None, :]
    src = (o + 0.5) * (in_size.double()[:, None] / out_size) - 0.5
    f0 = torch.floor(src)
    lerp = src - f0
    hi = in_size[:, None] - 1
    i0 = torch.minimum(torch.clamp(f0.long(), min=0), hi)
    i1 = torch.minimum(torch.clamp(torch.ceil(src).long(), min=0), hi)
    m = torch.zeros(in_size.shape[0], out_size, in_max, dtype=torch.float64)
    m.scatter_add_(2, i0[..., None], (1.0 - lerp)[..., None])
    m.scatter_add_(2, i1[..., None], lerp[..., None])
    return m


def _emu_crop_mats(labels, blank, char_width, H, W, out_hw):
    from fractions import Fraction

    cw = Fraction(char_width)
    is_blank = labels == blank
    first = torch.where(is_blank.any(1), is_blank.int().argmax(1), torch.full_like(labels[:, 0], 10 ** 6)).long()
    wc = torch.clamp((first * cw.numerator) // cw.denominator, min=1, max=W)
    my = _emu_interp(out_hw[0], torch.full((1,), H, dtype=torch.long), H)[0]
    mx = _emu_interp(out_hw[1], wc, W)
    return my, mx


def emu_crop_resize_fwd(img, labels, blank, char_width, out_hw):
    """Documented semantics of tbg_crop_resize_fwd (= the reference's crop + tf.image.resize, aster_inferer.py:153-190)."""
    my, mx = _emu_crop_mats(labels, blank, char_width, img.shape[2], img.shape[3], out_hw)
    x = torch.einsum("yh,bchw->bcyw", my, img.double())
    return torch.einsum("bcyw,bxw->byxc", x, mx).to(img.dtype)


def emu_crop_resize_bwd(g, labels, blank, char_width, img_hw):
    my, mx = _emu_crop_mats(labels, blank, char_width, img_hw[0], img_hw[1], (g.shape[1], g.shape[2]))
    t = torch.einsum("byxc,bxw->bcyw", g.double(), mx)
    return torch.einsum("yh,bcyw->bchw", my, t).to(g.dtype)


def emu_fromrgb_fwd(img, w, bias, coef, gain):
    y = torch.einsum("bjhw,jc->bhwc", img.double(), w.double()) * coef + bias.double()
    y = torch.where(y > 0, y, 0.2 * y) * gain
    from textboxgan_b200 import layers as L

    return y.to(L.ACT_DTYPE)


def emu_fromrgb_bwd(img, w, g_out, out, coef, gain, *, want_img=True, want_w=True):
    gp = g_out.double() * gain * torch.where(out.double() > 0, 1.0, 0.2)
    gimg = (torch.einsum("bhwc,jc->bjhw", gp, w.double()) * coef).float() if want_img else None
    gw = (torch.einsum("bjhw,bhwc->jc", img.double(), gp) * coef).float() if want_w else None
    gb = gp.sum(dim=(0, 1, 2)).float() if want_w else None
    return gimg, gw, gb


def emu_fir4(x, out_hw, off, scale, *, d=None, noise=None, noise_strength=None, bias=None, act=0, gain=1.0):
    """Documented semantics of tbg_fir4 (include/tbg.h)."""
    B, IH, IW, C = x.shape
    OH, OW = out_hw
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    # zero-extend so that every index y+m+off lands inside the padded tensor
    lo_y, lo_x = max(0, -off[0]), max(0, -off[1])
    hi_y, hi_x = max(0, OH + 3 + off[0] - IH), max(0, OW + 3 + off[1] - IW)
    xp = torch.nn.functional.pad(x.double(), (0, 0, lo_x, hi_x, lo_y, hi_y))
    acc = torch.zeros(B, OH, OW, C, dtype=torch.float64)
    for m in range(4):
        for n in range(4):
            y0, x0 = m + off[0] + lo_y, n + off[1] + lo_x
            acc += k[m] * k[n] * xp[:, y0:y0 + OH, x0:x0 + OW]
    acc = acc * scale
    if d is not None:
        acc = acc * d.double()[:, None, None, :]
    if noise is not None:
        acc = acc + noise.double()[..., None] * noise_strength.double().reshape(())
    if bias is not None:
        acc = acc + bias.double()
    if act == 1:
        acc = torch.where(acc > 0, acc, 0.2 * acc)
    return (acc * gain).to(x.dtype)


def _fir4_down_f64(x, out_hw, sy, off, scale):
    B, IH, IW, C = x.shape
    OH, OW = out_hw
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    lo_y, lo_x = max(0, -off[0]), max(0, -off[1])
    hi_y, hi_x = max(0, sy * (OH - 1) + 4 + off[0] - IH), max(0, 2 * (OW - 1) + 4 + off[1] - IW)
    xp = torch.nn.functional.pad(x.double(), (0, 0, lo_x, hi_x, lo_y, hi_y))
    acc = torch.zeros(B, OH, OW, C, dtype=torch.float64)
    for m in range(4):
        for n in range(4):
            y0, x0 = m + off[0] + lo_y, n + off[1] + lo_x
            acc = acc + k[m] * k[n] * xp[:, y0: y0 + sy * (OH - 1) + 1: sy, x0: x0 + 2 * (OW - 1) + 1: 2]
    return acc * scale


def emu_fir4_down(x, out_hw, sy, off, scale):
    """Documented semantics of tbg_fir4_down (include/tbg.h)."""
    return _fir4_down_f64(x, out_hw, sy, off, scale).to(x.dtype)


def emu_fir4_down_adjoint(g, in_hw, sy, off, scale, add=None, out=None):
    """Documented semantics of tbg_fir4_down_adjoint: the exact transpose of emu_fir4_down (by autograd) + add."""
    B, OH, OW, C = g.shape
    with torch.enable_grad():
        x = torch.zeros(B, in_hw[0], in_hw[1], C, dtype=torch.float64, requires_grad=True)
        y = _fir4_down_f64(x, (OH, OW), sy, off, scale)
        (gx,) = torch.autograd.grad(y, x, g.double())
    if add is not None:
        gx = gx + add.double()
    if out is not None:
        out.copy_(gx.to(out.dtype))
        return out
    return gx.to(g.dtype)


def emu_wfold_adj(gadj, spec, *, w_raw=None, s=None, t=None, out=None, flip=False):
    taps = spec.KH * spec.KW
    g = gadj.double().reshape(spec.Ipad, taps, spec.Opad)[: spec.I, :, : spec.O]
    if flip:
        g = g.flip(1)
    gw = (g.permute(1, 0, 2) * spec.coef).reshape(spec.KH, spec.KW, spec.I, spec.O)
    if s is not None:
        gq = (s.double() ** 2).t() @ t.double()
        gw = gw + 2.0 * spec.coef * spec.coef * w_raw.double() * gq[None, None]
    gw = gw.float()
    if out is not None:
        out += gw
        return out
    return gw


def emu_style_dense_fwd(style, ws, bs, idxs, coef):
    return [(coef * (style[:, i].double() @ w.double()) + b.double() + 1.0).to(style.dtype)
            for w, b, i in zip(ws, bs, idxs)]


def emu_style_dense_bwd(style, ws, gss, idxs, coef):
    gstyle = torch.zeros_like(style, dtype=torch.float64)
    gws, gbs = [], []
    for w, g, i in zip(ws, gss, idxs):
        gws.append((coef * style[:, i].double().t() @ g.double()).to(style.dtype))
        gbs.append(g.double().sum(0).to(style.dtype))
        gstyle[:, i] += coef * g.double() @ w.double().t()
    return gstyle.to(style.dtype), gws, gbs


def emu_demod_coef(s, q, eps=1e-8):
    return torch.rsqrt((s.double() ** 2) @ q.double() + eps).to(s.dtype)


def emu_demod_bwd(S1, Spre, Snz, d, ns, bias, s, q):
    """Documented semantics of tbg_demod_bwd (include/tbg.h)."""
    nsv = ns.double().reshape(()) if (ns is not None and Snz is not None) else 0.0
    snz = Snz.double() if Snz is not None else torch.zeros_like(S1).double()
    t = -0.5 * (Spre.double() - nsv * snz - bias.double()[None, :] * S1.double()) * d.double() ** 2
    gs = 2.0 * s.double() * (t @ q.double().t())
    return t.to(s.dtype), S1.double().sum(0).to(s.dtype), snz.sum().reshape(1).to(s.dtype), gs.to(s.dtype)


def emu_bias_act_bwd(g_out, out, *, residual=None, noise=None, d=None, act=True, gain=1.0, want_sums=True,
                     bias_grad_only=False):
    B, C = out.shape[0], out.shape[-1]
    o = out.double()
    if residual is not None:
        o = o - residual.double()
    neg = 0.0 if int(act) == 2 else 0.2
    slope = torch.where(o > 0, torch.ones_like(o), torch.full_like(o, neg)) if act else torch.ones_like(o)
    gp = g_out.double() * gain * slope
    pre = torch.where(slope > 0, o / (gain * slope.clamp_min(1e-30)), torch.zeros_like(o))
    dd = d.double().reshape(B, *([1] * (out.dim() - 2)), C) if d is not None else 1.0
    gy0 = (gp * dd).to(out.dtype)
    if bias_grad_only:
        return gy0, gp.reshape(-1, C).sum(0).float(), None, None
    if not want_sums:
        return gy0, None, None, None
    S1 = gp.reshape(B, -1, C).sum(1).float()
    Spre = (gp * pre).reshape(B, -1, C).sum(1).float()
    Snz = (gp * noise.double()[..., None]).reshape(B, -1, C).sum(1).float() if noise is not None \
        else torch.zeros(B, C)
    return gy0, S1, Spre, Snz


def emu_bias_act_rgb_bwd(g_out, out, g_rgb, ws, *, noise, d, act=1, gain=1.0):
    """Documented semantics of tbg_bias_act_rgb_bwd: bias_act_bwd on g_out + g_rgb (x) ws, plus the ToRGB weight gradient."""
    B = out.shape[0]
    gx_rgb = torch.bmm(g_rgb.double().reshape(B, -1, 3), ws.double().transpose(1, 2)).reshape(out.shape)
    g = gx_rgb if g_out is None else g_out.double() + gx_rgb
    gy0, S1, Spre, Snz = emu_bias_act_bwd(g, out, noise=noise, d=d, act=act, gain=gain)
    gws = torch.bmm(out.double().reshape(B, -1, out.shape[-1]).transpose(1, 2), g_rgb.double().reshape(B, -1, 3)).float()
    return gy0, S1, Spre, Snz, gws


def emu_torgb_fwd(x, ws, bias):
    B = x.shape[0]
    y = torch.bmm(x.double().reshape(B, -1, x.shape[-1]), ws.double()).reshape(*x.shape[:-1], 3)
    if bias is not None:
        y = y + bias.double()
    return y.float()


def emu_torgb_bwd(x, ws, gy):
    B = x.shape[0]
    g = gy.double().reshape(B, -1, 3)
    gx = torch.bmm(g, ws.double().transpose(1, 2)).reshape(x.shape).to(x.dtype)
    gws = torch.bmm(x.double().reshape(B, -1, x.shape[-1]).transpose(1, 2), g).float()
    return gx, gws


def emu_wprep(w_raw, spec, *, want_adj=True, want_q=False, act_dtype=None):
    """Documented semantics of tbg_wprep (include/tbg.h) with the spec's per-axis tables."""
    from textboxgan_b200 import layers as L

    fy, fx, ay, ax = [t.double() for t in spec.tables]
    w = w_raw.double() * spec.coef
    wp = torch.zeros(spec.KH, spec.KW, spec.Ipad, spec.Opad, dtype=torch.float64)
    wp[:, :, : spec.I, : spec.O] = w
    fwd = torch.einsum("ptk,qul,klio->pqotui", fy, fx, wp).reshape(spec.fwd_rows, spec.fwd_cols).to(L.ACT_DTYPE)
    adj = torch.einsum("ptk,qul,klio->pqituo", ay, ax, wp).reshape(spec.adj_rows, spec.adj_cols).to(L.ACT_DTYPE) \
        if want_adj else None
    q = (w * w).sum(dim=(0, 1)).float() if want_q else None
    return fwd, adj, q


def emu_wfold(gfwd, spec, *, gq=None, w_raw=None, out=None, s=None, t=None):
    if s is not None:
        gq = (s.double() ** 2).t() @ t.double()
    fy, fx, _, _ = [t.double() for t in spec.tables]
    g6 = gfwd.double().reshape(fy.shape[0], fx.shape[0], spec.Opad, fy.shape[1], fx.shape[1], spec.Ipad)
    gw = torch.einsum("ptk,qul,pqotui->klio", fy, fx, g6)[:, :, : spec.I, : spec.O] * spec.coef
    if gq is not None:
        gw = gw + 2.0 * spec.coef * spec.coef * w_raw.double() * gq.double()[None, None]
    gw = gw.float()
    if out is not None:
        out += gw
        return out
    return gw


def emu_attn_decoder_fwd(mem, keys, w, steps):
    """Documented semantics of tbg_attn_decoder_fwd == oracle/aster.py::attention_decoder."""
    B, T, _ = mem.shape
    f = lambda n: w[n].double()
    m, ky = mem.double(), keys.double()
    h = torch.zeros(B, 256, dtype=torch.float64)
    c = torch.zeros(B, 256, dtype=torch.float64)
    prev = torch.zeros(B, dtype=torch.long)
    sv = dict(a=[], ctx=[], gates=[], c=[], h=[], prev=[])
    logits = []
    for _ in range(steps):
        q = h @ f("wq")
        e = torch.tanh(ky + q[:, None, :]) @ f("v")
        a = torch.softmax(e, dim=1)
        ctx = (a[:, :, None] * m).sum(1)
        gates = torch.cat([f("emb")[prev], ctx, h], 1) @ f("wg") + f("b")
        i, fg, g, o = gates.chunk(4, dim=1)
        i, fg, g, o = torch.sigmoid(i), torch.sigmoid(fg), torch.tanh(g), torch.sigmoid(o)
        c = fg * c + i * g
        h = o * torch.tanh(c)
        lg = torch.cat([h, ctx], 1) @ f("wd") + f("bd")
        sv["prev"].append(prev.clone())
        prev = lg.argmax(dim=1)
        for k, v in (("a", a), ("ctx", ctx), ("gates", torch.cat([i, fg, g, o], 1)), ("c", c), ("h", h)):
            sv[k].append(v)
        logits.append(lg)
    out = {k: torch.stack(v, 1).float() if k != "prev" else torch.stack(v, 1).int() for k, v in sv.items()}
    return torch.stack(logits, 1).float(), out


def emu_attn_decoder_bwd(mem, keys, w, g_logits, sv):
    B, T, _ = mem.shape
    steps = g_logits.shape[1]
    f = lambda n: w[n].double()
    m, ky, gl = mem.double(), keys.double(), g_logits.double()
    g_mem = torch.zeros_like(m)
    g_keys = torch.zeros_like(ky)
    dh = torch.zeros(B, 256, dtype=torch.float64)
    dc = torch.zeros(B, 256, dtype=torch.float64)
    for st in range(steps - 1, -1, -1):
        dcat = gl[:, st] @ f("wd").t()
        dh = dh + dcat[:, :256]
        dctx = dcat[:, 256:]
        i, fg, g, o = sv["gates"][:, st].double().chunk(4, dim=1)
        cc = sv["c"][:, st].double()
        cp = sv["c"][:, st - 1].double() if st > 0 else torch.zeros_like(cc)
        hp = sv["h"][:, st - 1].double() if st > 0 else torch.zeros_like(cc)
        tc = torch.tanh(cc)
        do = dh * tc * o * (1 - o)
        dcc = dh * o * (1 - tc * tc) + dc
        dg = torch.cat([dcc * g * i * (1 - i), dcc * cp * fg * (1 - fg), dcc * i * (1 - g * g), do], 1)
        dc = dcc * fg
        dx = dg @ f("wg").t()
        dctx = dctx + dx[:, 256:768]
        dh_prev = dx[:, 768:]
        a = sv["a"][:, st].double()
        da = (dctx[:, None, :] * m).sum(2)
        g_mem += a[:, :, None] * dctx[:, None, :]
        de = a * (da - (a * da).sum(1, keepdim=True))
        q = hp @ f("wq")
        th = torch.tanh(ky + q[:, None, :])
        dp = de[:, :, None] * f("v")[None, None, :] * (1 - th * th)
        g_keys += dp
        dh = dh_prev + dp.sum(1) @ f("wq").t()
    return g_mem.float(), g_keys.float()


def emu_bias_act_fwd(t, *, noise=None, noise_strength=None, bias=None, act=1, gain=1.0):
    """Documented semantics of tbg_bias_act_fwd (include/tbg.h)."""
    v = t.double()
    if noise is not None:
        v = v + noise.double()[..., None] * noise_strength.double().reshape(())
    if bias is not None:
        v = v + bias.double()
    return (_emu_act(v, act) * gain).to(t.dtype)


def emu_rowdot(a, b):
    return (a.double() * b.double()).reshape(a.shape[0], -1, a.shape[-1]).sum(1).float()


def emu_batch_resize_normalize(packed, offsets, src_h, src_w, dst_w, H, W):
    """Documented semantics of tbg_batch_resize_normalize (float bilinear at pixel centres, rounded to the uint8 grid)."""
    import numpy as np

    B = offsets.shape[0]
    out = np.zeros((B, 3, H, W), dtype=np.float32)
    buf = packed.cpu().numpy()
    for b in range(B):
        sh, sw, dw = int(src_h[b]), int(src_w[b]), int(dst_w[b])
        img = buf[int(offsets[b]): int(offsets[b]) + sh * sw * 3].reshape(sh, sw, 3).astype(np.float32)
        if sw == 2 * dw and sh == 2 * H:
            i = img.astype(np.int64)
            r = ((i[0::2, 0::2] + i[0::2, 1::2] + i[1::2, 0::2] + i[1::2, 1::2] + 2) >> 2).astype(np.float32)
        else:
            def axis(n_dst, n_src):
                f = (np.arange(n_dst, dtype=np.float32) + np.float32(0.5)) * (np.float32(n_src) / np.float32(n_dst)) - np.float32(0.5)
                i0 = np.floor(f).astype(np.int64)
                f = (f - i0).astype(np.float32)
                lo, hi = i0 < 0, i0 >= n_src - 1
                f[lo | hi] = 0
                i0[lo] = 0
                i0[hi] = n_src - 1
                return i0, np.minimum(i0 + 1, n_src - 1), f
            x0, x1, fx = axis(dw, sw)
            y0, y1, fy = axis(H, sh)
            top = img[y0][:, x0] + (img[y0][:, x1] - img[y0][:, x0]) * fx[None, :, None]
            bot = img[y1][:, x0] + (img[y1][:, x1] - img[y1][:, x0]) * fx[None, :, None]
            r = np.rint(top + (bot - top) * fy[:, None, None])
        out[b, :, :, :dw] = (r / np.float32(127.5) - np.float32(1.0)).transpose(2, 0, 1)
    return torch.from_numpy(out)


def _emu_act(pre, act):
    if act == 1:
        return torch.where(pre > 0, pre, 0.2 * pre)
    if act == 2:
        return torch.relu(pre)
    return pre


def emu_dense_fwd(x, w, bias, *, coef=1.0, bias_coef=1.0, act=0, gain=1.0):
    """Documented semantics of tbg_dense_fwd (include/tbg.h)."""
    pre = (x.double() @ w.double()) * coef
    if bias is not None:
        pre = pre + bias.double() * bias_coef
    return (_emu_act(pre, act) * gain).to(x.dtype)


def emu_dense_bwd(x, w, y, gy, *, coef=1.0, bias_coef=1.0, act=0, gain=1.0, want_gx=True, want_gw=True, want_gb=True):
    g = gy.double() * gain
    if act == 1:
        g = g * torch.where(y.double() > 0, 1.0, 0.2)
    elif act == 2:
        g = g * (y.double() > 0)
    gx = (coef * g @ w.double().t()).to(x.dtype) if want_gx else None
    gw = (coef * x.double().t() @ g).to(x.dtype) if want_gw else None
    gb = (bias_coef * g.sum(0)).to(x.dtype) if want_gb else None
    return gx, gw, gb


def emu_pixel_norm_fwd(x):
    xd = x.double()
    return (xd * torch.rsqrt((xd * xd).mean(1, keepdim=True) + 1e-8)).to(x.dtype)


def emu_pixel_norm_bwd(x, gy):
    with torch.enable_grad():                       # called from autograd.Function.backward (grad mode off)
        xd = x.detach().double().requires_grad_(True)
        y = xd * torch.rsqrt((xd * xd).mean(1, keepdim=True) + 1e-8)
        (g,) = torch.autograd.grad(y, xd, gy.double())
    return g.to(x.dtype)


def _emu_word_layout(act, B, mcn, out_hwc):
    oh, ow, oc = out_hwc
    # reference: reshape [B, out_w, out_c, out_h] then transpose (0,2,3,1) -> NCHW; NHWC = permute(0, 3, 1, 2)
    return act.reshape(B, ow, oc, oh).permute(0, 3, 1, 2)


def emu_word_encoder_fwd(words, w0, table, mask, keep, fc_w, fc_b, out_hwc, act_dtype=None):
    """Documented semantics of tbg_word_encoder_fwd (= word_encoder.py:39-63)."""
    from textboxgan_b200 import layers as L

    B, mcn = words.shape
    full = torch.cat([w0.double(), table.double()], 0)
    emb = full[words.long()].reshape(B * mcn, -1)
    if mask is not None:
        emb = emb * mask.double().reshape(B * mcn, -1) / keep
    act = torch.relu(emb @ fc_w.double() + fc_b.double())
    out = _emu_word_layout(act, B, mcn, out_hwc).contiguous().to(L.ACT_DTYPE)
    return out, emb.float(), act.float()


def emu_word_encoder_bwd(words, mask, keep, fc_w, emb, act, g_out, table_rows, out_hwc):
    B, mcn = words.shape
    oh, ow, oc = out_hwc
    g = g_out.double().permute(0, 2, 3, 1).reshape(B * mcn, -1)        # inverse of the forward layout change
    gpre = g * (act.double() > 0)
    g_emb = gpre @ fc_w.double().t()
    if mask is not None:
        g_emb = g_emb * mask.double().reshape(B * mcn, -1) / keep
    g_table = torch.zeros(table_rows + 1, fc_w.shape[0], dtype=torch.float64)
    g_table.index_add_(0, words.long().reshape(-1), g_emb)
    return g_table[1:].float(), (emb.double().t() @ gpre).float(), gpre.sum(0).float()


def _emu_mbstd(xd, n_calls, group_size):
    Bt = xd.shape[0]
    B = Bt // n_calls
    G = min(group_size, B)
    y = xd.reshape(n_calls, G, B // G, -1)
    y = y - y.mean(dim=1, keepdim=True)
    sd = torch.sqrt((y * y).mean(dim=1) + 1e-8).mean(dim=2)               # [n_calls, B/G]
    return sd[:, None, :].expand(n_calls, G, B // G).reshape(Bt)


def emu_minibatch_std_fwd(x, n_calls, cpad, group_size=4):
    """Documented semantics of tbg_minibatch_std_fwd (= mini_batch_std.py:10-35 per call)."""
    Bt, H, W_, C = x.shape
    stat = _emu_mbstd(x.double(), n_calls, group_size)
    xcat = torch.zeros(Bt, H, W_, cpad, dtype=torch.float64)
    xcat[..., :C] = x.double()
    xcat[..., C] = stat[:, None, None]
    return xcat.to(x.dtype), stat.float()


def emu_minibatch_std_bwd(x, gxcat, n_calls, group_size=4):
    C = x.shape[3]
    with torch.enable_grad():
        xd = x.detach().double().requires_grad_(True)
        stat = _emu_mbstd(xd, n_calls, group_size)
        gstat = gxcat.double()[..., C].sum(dim=(1, 2))
        (g,) = torch.autograd.grad(stat, xd, gstat)
    return (g + gxcat.double()[..., :C]).to(x.dtype)


def emu_torgb_skip_fwd(x, ws, bias, y_prev, words, nchw):
    """Documented semantics of tbg_torgb_skip_fwd: ToRGB + upsample_2d(y_prev) (literal upfirdn_2d_ref) + mask + layout."""
    from oracle.stylegan import upfirdn_2d_ref
    import numpy as np

    y = emu_torgb_fwd(x, ws, bias).double()
    B, H, W_, _ = y.shape
    if y_prev is not None:
        t = np.array([1.0, 3.0, 3.0, 1.0])
        k = np.outer(t, t)
        k = k / k.sum() * 4.0
        y = y + upfirdn_2d_ref(y_prev.double(), k, 2, 2, 1, 1, 2, 1, 2, 1)
    if words is not None:
        mcn = words.shape[1]
        idx = (torch.arange(W_) * mcn) // W_
        keep = (words.long()[:, idx] != 0).double()                         # [B, W]
        y = y * keep[:, None, :, None]
    y = y.float()
    return y.permute(0, 3, 1, 2).contiguous() if nchw else y


def emu_image_grad_nhwc(g, words):
    B, _, H, W_ = g.shape
    out = g.permute(0, 2, 3, 1).double()
    if words is not None:
        mcn = words.shape[1]
        idx = (torch.arange(W_) * mcn) // W_
        out = out * (words.long()[:, idx] != 0).double()[:, None, :, None]
    return out.float().contiguous()


class emu_WPrepPlan:
    """K.WPrepPlan on the CPU: the same outputs as one emu_wprep call per entry, in persistent buffers."""

    def __init__(self, entries):
        self.entries = list(entries)
        self.outputs = [tuple(None if t is None else torch.empty_like(t) for t in emu_wprep(w, spec, want_adj=a, want_q=q))
                        for w, spec, a, q in self.entries]

    def run(self):
        for (w, spec, a, q), outs in zip(self.entries, self.outputs):
            for dst, src in zip(outs, emu_wprep(w, spec, want_adj=a, want_q=q)):
                if dst is not None:
                    dst.copy_(src)
        return self.outputs


@contextlib.contextmanager
def emulated_kernels(act_dtype=torch.float32):
    """Route textboxgan_b200.kernels through the CPU emulation (tests only)."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import layers as L

    saved = (K.conv2d_igemm, K.conv2d_wgrad, C._as_bf16, L.ACT_DTYPE, K.upfirdn2d, K.adam_step, K.ema_step,
             K.lstm_seq_fwd, K.lstm_seq_bwd, K.modulate, K.modulate_bwd, K.bias_act_bwd, K.torgb_fwd, K.torgb_bwd)
    K.modulate, K.modulate_bwd, K.bias_act_bwd = emu_modulate, emu_modulate_bwd, emu_bias_act_bwd
    K.torgb_fwd, K.torgb_bwd = emu_torgb_fwd, emu_torgb_bwd
    saved_w = (K.wprep, K.wfold, K.attn_decoder_fwd, K.attn_decoder_bwd)
    saved_c = (K.crop_resize_fwd, K.crop_resize_bwd)
    K.crop_resize_fwd, K.crop_resize_bwd = emu_crop_resize_fwd, emu_crop_resize_bwd
    saved_f = (K.fir4, K.wfold_adj, K.fromrgb_fwd, K.fromrgb_bwd)
    K.fir4, K.wfold_adj = emu_fir4, emu_wfold_adj
    K.fromrgb_fwd, K.fromrgb_bwd = emu_fromrgb_fwd, emu_fromrgb_bwd
    saved_d = (K.demod_coef, K.demod_bwd, K.style_dense_fwd, K.style_dense_bwd)
    K.demod_coef, K.demod_bwd = emu_demod_coef, emu_demod_bwd
    K.style_dense_fwd, K.style_dense_bwd = emu_style_dense_fwd, emu_style_dense_bwd
    K.wprep, K.wfold = emu_wprep, emu_wfold
    K.attn_decoder_fwd, K.attn_decoder_bwd = emu_attn_decoder_fwd, emu_attn_decoder_bwd
    K.conv2d_igemm = emu_conv2d_igemm
    K.conv2d_wgrad = emu_conv2d_wgrad
    K.upfirdn2d = emu_upfirdn2d
    K.adam_step = emu_adam_step
    K.ema_step = emu_ema_step
    K.lstm_seq_fwd = emu_lstm_seq_fwd
    K.lstm_seq_bwd = emu_lstm_seq_bwd
    C._as_bf16 = lambda t: t.contiguous()
    L.ACT_DTYPE = act_dtype
    new_names = ("dense_fwd", "dense_bwd", "pixel_norm_fwd", "pixel_norm_bwd", "word_encoder_fwd", "word_encoder_bwd",
                 "minibatch_std_fwd", "minibatch_std_bwd", "torgb_skip_fwd", "image_grad_nhwc", "bias_act_fwd", "rowdot",
                 "batch_resize_normalize", "fir4_down", "fir4_down_adjoint", "bias_act_rgb_bwd")
    new_names = new_names + ("WPrepPlan",)
    saved_n = {n: getattr(K, n) for n in new_names}
    for n in new_names:
        setattr(K, n, globals()["emu_" + n])
    try:
        yield
    finally:
        for n, f in saved_n.items():
            setattr(K, n, f)
        (K.conv2d_igemm, K.conv2d_wgrad, C._as_bf16, L.ACT_DTYPE, K.upfirdn2d, K.adam_step, K.ema_step,
         K.lstm_seq_fwd, K.lstm_seq_bwd, K.modulate, K.modulate_bwd, K.bias_act_bwd, K.torgb_fwd,
         K.torgb_bwd) = saved
        K.wprep, K.wfold, K.attn_decoder_fwd, K.attn_decoder_bwd = saved_w
        K.demod_coef, K.demod_bwd, K.style_dense_fwd, K.style_dense_bwd = saved_d
        K.fir4, K.wfold_adj, K.fromrgb_fwd, K.fromrgb_bwd = saved_f
        K.crop_resize_fwd, K.crop_resize_bwd = saved_c
