"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the
C ABI, against the CPU oracle on the same seeded inputs.

Tolerances: integer paths bit-exact; fp32-output kernels 2e-5 relative (fp32 accumulation order);
bf16-output kernels 1e-2 relative to the tensor maximum (one bf16 rounding of the output, 2^-8);
whole-model quantities in bf16 activations 5e-2 (losses) — stated next to each assertion.
"""
import copy
import math
import zlib

import pytest
import torch

from common import perturbed_params, rel_err, small_cfg
from emu import emu_conv2d_igemm, emu_conv2d_wgrad
from oracle import aster as OA
from oracle import stylegan as OS
from oracle import train_step as OT

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf16_round(t):
    return t.to(torch.bfloat16).float()


def _rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _geoms():
    from textboxgan_b200 import conv as C

    return {
        "plain3": C.plain_geom(16, 64, 128, 128, 3),
        "plain1": C.plain_geom(8, 32, 64, 64, 1),
        "plain3-lowres": C.plain_geom(4, 16, 512, 256, 3),
        "plain3-oob-batch": C.plain_geom(2, 8, 128, 512, 3),
        "plain3-4x4": C.plain_geom(4, 4, 576, 512, 3),
        "up": C.up_geom(8, 32, 128, 64),
        "down3": C.down_geom(16, 64, 64, 128, 3, True),
        "down1": C.down_geom(16, 64, 64, 128, 1, True),
        "down3-w": C.down_geom(8, 32, 128, 128, 3, False),
        "down1-w": C.down_geom(8, 32, 64, 64, 1, False),
        "aster-s21": C.ConvGeom(8, 32, 128, 128, C.Axis("s2", 1, 0), C.Axis("s1", 1, 0)),
    }


@pytest.mark.parametrize("name", list(_geoms().keys()))
@pytest.mark.parametrize("batch", [3, 8])
def test_conv_forward_adjoint_wgrad_vs_emulated_semantics(name, batch):
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    g = _geoms()[name]
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)    # str hash() is salted per process
    x = _bf16_round(torch.randn(batch, g.H, g.W, g.cin, generator=gen))
    w = _bf16_round(torch.randn(g.n_total, g.k_total, generator=gen) / math.sqrt(g.k_total))
    kw = g.kernel_kwargs()
    # forward, fp32 output: only the accumulation order differs
    y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), out_fp32=True, **kw)
    ref = emu_conv2d_igemm(x, w, out_fp32=True, **kw)
    assert rel_err(y, ref) < 2e-5
    # bf16 output
    y16 = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), **kw)
    assert rel_err(y16.float(), ref) < 1e-2
    # adjoint geometry on re-laid-out weights
    a = g.adjoint()
    wa = _bf16_round(C.relayout_for_adjoint(w, g)).contiguous()
    gy = _bf16_round(torch.randn(ref.shape, generator=gen))
    gx = K.conv2d_igemm(gy.to(DEV).bfloat16(), wa.to(DEV).bfloat16(), out_fp32=True, **a.kernel_kwargs())
    gref = emu_conv2d_igemm(gy, wa, out_fp32=True, **a.kernel_kwargs())
    assert rel_err(gx, gref) < 2e-5
    # <conv(x), gy> == <x, conv^T(gy)>  (size-independent property, exact up to fp32 rounding)
    lhs = (y.double().cpu() * gy.double()).sum()
    rhs = (x.double() * gx.double().cpu()).sum()
    # (both sides are sums with cancellation: the bound is relative to the sum of magnitudes)
    assert abs(lhs - rhs) <= 1e-5 * ((y.double().cpu() * gy.double()).abs().sum() + 1.0)
    # weight gradient
    gw = K.conv2d_wgrad(x.to(DEV).bfloat16(), gy.to(DEV).bfloat16(), **kw)
    gwref = emu_conv2d_wgrad(x, gy, **kw)
    assert rel_err(gw, gwref) < 2e-4


def test_conv_fused_epilogue_matches_reference_order():
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    g = C.up_geom(8, 32, 128, 128)
    gen = torch.Generator().manual_seed(3)
    B = 5
    x = _bf16_round(torch.randn(B, g.H, g.W, g.cin, generator=gen))
    w = _bf16_round(torch.randn(g.n_total, g.k_total, generator=gen) / math.sqrt(g.k_total))
    epi = dict(col_scale=torch.rand(B, g.cout, generator=gen) + 0.5, bias=torch.randn(g.cout, generator=gen) * 0.1,
               noise=torch.randn(B, 16, 64, generator=gen), noise_strength=torch.tensor([0.3]),
               residual=_bf16_round(torch.randn(B, 16, 64, g.cout, generator=gen)), res_scale=1 / math.sqrt(2),
               act=1, act_gain=math.sqrt(2))
    ref = emu_conv2d_igemm(x, w, out_fp32=True, **g.kernel_kwargs(), **epi)
    dev_epi = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in epi.items()}
    dev_epi["residual"] = dev_epi["residual"].bfloat16()
    y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), out_fp32=True, **g.kernel_kwargs(), **dev_epi)
    assert rel_err(y, ref) < 2e-5
    for res_first, act in ((True, 2), (False, 0)):
        e2 = dict(epi, res_first=res_first, act=act, act_gain=1.0, res_scale=1.0)
        d2 = dict(dev_epi, res_first=res_first, act=act, act_gain=1.0, res_scale=1.0)
        ref = emu_conv2d_igemm(x, w, out_fp32=True, **g.kernel_kwargs(), **e2)
        y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), out_fp32=True, **g.kernel_kwargs(), **d2)
        assert rel_err(y, ref) < 2e-5


def test_conv_linearity_at_baseline_size():
    """BASELINE configs[1] top layer (B=32, 32x128, 128->128): conv(a+b) == conv(a)+conv(b) and
    scaling, on fp32 outputs (properties independent of an oracle that could not finish in seconds)."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    g = C.plain_geom(32, 128, 128, 128, 3)
    gen = torch.Generator(device=DEV).manual_seed(0)
    a = torch.randn(32, 32, 128, 128, device=DEV, generator=gen).bfloat16()
    b = torch.randn(32, 32, 128, 128, device=DEV, generator=gen).bfloat16()
    s = (a.float() + b.float()).bfloat16()
    w = (torch.randn(128, 9 * 128, device=DEV, generator=gen) / 34.0).bfloat16()
    ya = K.conv2d_igemm(a, w, out_fp32=True, **g.kernel_kwargs())
    yb = K.conv2d_igemm(b, w, out_fp32=True, **g.kernel_kwargs())
    ys = K.conv2d_igemm(s, w, out_fp32=True, **g.kernel_kwargs())
    # s is a bf16 rounding of a+b: compare against the conv of the rounding error too
    r = (s.float() - a.float() - b.float()).bfloat16()
    yr = K.conv2d_igemm(r, w, out_fp32=True, **g.kernel_kwargs())
    assert rel_err(ys, ya + yb + yr) < 1e-4
    y2 = K.conv2d_igemm((a.float() * 2).bfloat16(), w, out_fp32=True, **g.kernel_kwargs())
    assert rel_err(y2, 2 * ya) < 1e-6


@pytest.mark.parametrize("cfgk", [dict(upx=2, upy=2, padx0=2, padx1=1, pady0=2, pady1=1),
                                  dict(padx0=1, padx1=1, pady0=1, pady1=1),
                                  dict(downx=2, downy=2, padx0=2, padx1=3, pady0=2, pady1=3),
                                  dict(upx=2, upy=1, downx=1, downy=2, padx0=-1, padx1=2, pady0=0, pady1=1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upfirdn2d_matches_reference_op(cfgk, dtype):
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(1)
    x = torch.randn(6, 9, 13, 3, generator=gen)
    if dtype == torch.bfloat16:
        x = _bf16_round(x)
    k = torch.randn(4, 4, generator=gen)
    full = dict(upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0, pady1=0)
    full.update(cfgk)
    ref = OS.upfirdn_2d_ref(x.double(), k.double().numpy(), full["upx"], full["upy"], full["downx"], full["downy"],
                            full["padx0"], full["padx1"], full["pady0"], full["pady1"])
    y = K.upfirdn2d(x.to(DEV).to(dtype), k.to(DEV), **full)
    assert y.shape == ref.shape
    assert rel_err(y.float(), ref) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_upfirdn2d_gradients_any_order():
    from textboxgan_b200 import upfirdn as U

    x = torch.randn(2, 4, 8, 3, device=DEV, requires_grad=True)
    y = U.upsample_2d_nhwc(x)
    assert y.shape == (2, 8, 16, 3)
    r = torch.randn_like(y)
    (gx,) = torch.autograd.grad((y * r).sum(), x, create_graph=True)
    xr = x.detach().cpu().double().requires_grad_(True)
    k, p0, p1 = OS.compute_paddings([1, 3, 3, 1], True, False, is_conv=False)
    yr = OS.upfirdn_2d_ref(xr, k, 2, 2, 1, 1, p0, p1, p0, p1)
    (gr,) = torch.autograd.grad((yr * r.cpu().double()).sum(), xr)
    assert rel_err(y, yr) < 1e-5 and rel_err(gx, gr) < 1e-5
    # second order: gradient of <gx, q> w.r.t. r's coefficient is linear -> just check it runs and is finite
    q = torch.randn_like(gx)
    r2 = r.clone().requires_grad_(True)
    (gx2,) = torch.autograd.grad((U.upsample_2d_nhwc(x) * r2).sum(), x, create_graph=True)
    (gr2,) = torch.autograd.grad((gx2 * q).sum(), r2)
    assert rel_err(gr2, U.upsample_2d_nhwc(q)) < 1e-5


def test_adam_and_ema_kernels():
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(0)
    for n, off in ((1000003, 0), (4099, 1), (7, 3)):
        base = torch.randn(n + 8, generator=gen)
        p = base.to(DEV)[off: off + n]
        g = torch.randn(n + 8, generator=gen).to(DEV)[off: off + n]
        m = torch.randn(n + 8, generator=gen).to(DEV)[off: off + n].abs()
        v = torch.randn(n + 8, generator=gen).to(DEV)[off: off + n].abs()
        pr, gr, mr, vr = (t.detach().cpu().double() for t in (p, g, m, v))
        K.adam_step(p, g, m, v, torch.tensor([0.0017], device=DEV) if off == 1 else 0.0017, 0.0, 0.99, 1e-8)
        mr2 = 0.0 * mr + gr
        vr2 = 0.99 * vr + 0.01 * gr * gr
        ref = pr - 0.0017 * mr2 / (vr2.sqrt() + 1e-8)
        assert rel_err(p, ref) < 1e-6 and rel_err(m, mr2) < 1e-6 and rel_err(v, vr2) < 1e-6
        src = torch.randn(n + 8, generator=gen).to(DEV)[off: off + n]
        before = p.detach().cpu().double()
        K.ema_step(p, src, 0.99)
        assert rel_err(p, src.cpu().double() + (before - src.cpu().double()) * 0.99) < 1e-6


@pytest.mark.parametrize("B,H,W,C", [(4, 4, 16, 512), (3, 32, 128, 128), (5, 2, 8, 64)])
def test_fused_pointwise_kernels_vs_emulated_semantics(B, H, W, C):
    import emu
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(B * 1000 + C)
    x = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    g = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    s = torch.randn(B, C, generator=gen) + 1.0
    d = torch.rand(B, C, generator=gen) + 0.5
    nz = torch.randn(B, H, W, generator=gen)
    res = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    xd, gd, rd = (t.to(DEV).bfloat16() for t in (x, g, res))
    xs = K.modulate(xd, s.to(DEV))
    assert rel_err(xs.float(), emu.emu_modulate(x, s)) < 1e-2
    gx, gs = K.modulate_bwd(gd, xd, s.to(DEV))
    rgx, rgs = emu.emu_modulate_bwd(g, x, s)
    assert rel_err(gx.float(), rgx) < 1e-2 and rel_err(gs, rgs) < 1e-4
    for kw in (dict(noise=nz, d=d, act=True, gain=math.sqrt(2)), dict(residual=res, act=True, gain=1.0),
               dict(act=False, gain=1.0)):
        dkw = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items()}
        if "residual" in dkw:
            dkw["residual"] = dkw["residual"].bfloat16()
        out = K.bias_act_bwd(gd, xd, **dkw)
        ref = emu.emu_bias_act_bwd(g, x, **kw)
        assert rel_err(out[0].float(), ref[0]) < 1e-2
        assert rel_err(out[1], ref[1]) < 1e-4 and rel_err(out[2], ref[2]) < 1e-4
        if "noise" in kw:
            assert rel_err(out[3], ref[3]) < 1e-4
    ws = torch.randn(B, C, 3, generator=gen) / math.sqrt(C)
    bias = torch.randn(3, generator=gen)
    y = K.torgb_fwd(xd, ws.to(DEV), bias.to(DEV))
    assert rel_err(y, emu.emu_torgb_fwd(x, ws, bias)) < 1e-5
    gy = torch.randn(B, H, W, 3, generator=gen)
    tgx, tgws = K.torgb_bwd(xd, ws.to(DEV), gy.to(DEV))
    rtx, rtw = emu.emu_torgb_bwd(x, ws, gy)
    assert rel_err(tgx.float(), rtx) < 1e-2 and rel_err(tgws, rtw) < 1e-4


@pytest.mark.parametrize("kind,k,rh,I,O", [("plain", 3, True, 128, 256), ("plain", 1, True, 64, 128), ("up", 3, True, 128, 64),
                                          ("down", 3, True, 64, 128), ("down", 1, True, 128, 128),
                                          ("down", 3, False, 256, 256), ("plain", 3, True, 513, 512)])
def test_weight_prep_and_fold_vs_emulated_semantics(kind, k, rh, I, O):
    import emu
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import layers as L

    spec = C.weight_spec(kind, 8, 16, I, O, k, rh)
    gen = torch.Generator().manual_seed(I + O + k)
    w = torch.randn(k, k, I, O, generator=gen)
    fwd, adj, q = K.wprep(w.to(DEV), spec, want_adj=True, want_q=True)
    rf, ra, rq = emu.emu_wprep(w, spec, want_adj=True, want_q=True)
    # outputs are bf16 roundings of identical fp32 sums (<= 9 terms): bit-exact up to 1 ulp
    assert rel_err(fwd.float(), rf.float()) < 4e-3 and rel_err(adj.float(), ra.float()) < 4e-3
    assert rel_err(q, rq) < 1e-5
    g = torch.randn(spec.fwd_rows, spec.fwd_cols, generator=gen)
    gq = torch.randn(I, O, generator=gen)
    out = K.wfold(g.to(DEV), spec, gq=gq.to(DEV), w_raw=w.to(DEV))
    ref = emu.emu_wfold(g, spec, gq=gq, w_raw=w)
    assert rel_err(out, ref) < 1e-5
    # fold is the exact transpose of prep: <prep(w), g> == <w, fold(g)>
    lhs = (emu.emu_wprep(w.double(), spec, want_adj=False)[0].double() * g.double()).sum() if L.ACT_DTYPE == torch.float64 else None
    out2 = K.wfold(g.to(DEV), spec)
    rhs = (w.double() * out2.cpu().double()).sum()
    fy, fx, _, _ = [t.double() for t in spec.tables]
    wp = torch.zeros(k, k, spec.Ipad, spec.Opad, dtype=torch.float64)
    wp[:, :, :I, :O] = w.double() * spec.coef
    lhs = (torch.einsum("ptk,qul,klio->pqotui", fy, fx, wp).reshape(g.shape) * g.double()).sum()
    assert abs(lhs - rhs) < 1e-4 * (abs(lhs) + 1)


def test_grouped_weight_preparation_equals_single_launches():
    """tbg_wprep_group (every weight of an iteration in one launch, K.WPrepPlan) writes bit-identical matrices to one
    tbg_wprep call per weight, for every geometry kind, with and without the adjoint / q outputs, and on a re-run after
    the weights changed in place (as the optimiser does)."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    kinds = [("plain", 3, True, 128, 256), ("plain", 1, True, 64, 128), ("up", 3, True, 128, 64), ("down", 3, True, 64, 128),
             ("down", 1, True, 128, 128), ("down", 3, False, 256, 256), ("plain", 3, True, 513, 512), ("upT", 3, True, 128, 128),
             ("downU", 3, True, 64, 128)]
    gen = torch.Generator().manual_seed(5)
    entries = []
    for n, (kind, k, rh, I, O) in enumerate(kinds):
        spec = C.weight_spec(kind, 8, 16, I, O, k, rh, "t%d" % n)
        w = torch.randn(k, k, I, O, generator=gen).to(DEV)
        entries.append((w, spec, n % 3 != 1, n % 2 == 0))
    plan = K.WPrepPlan(entries)
    for rerun in range(2):
        outs = plan.run()
        for (w, spec, want_adj, want_q), got in zip(entries, outs):
            ref = K.wprep(w, spec, want_adj=want_adj, want_q=want_q)
            for a, b in zip(got, ref):
                assert (a is None) == (b is None)
                if a is not None:
                    assert torch.equal(a, b), (spec.geom.tag, rerun)
        for w, _, _, _ in entries:
            w.mul_(1.5).add_(0.01)


@pytest.mark.parametrize("B,I,O", [(4, 512, 512), (32, 128, 64), (3, 192, 512), (5, 64, 3)])
def test_demodulation_kernels_vs_emulated_semantics(B, I, O):
    """tbg_demod_coef / tbg_demod_bwd / tbg_wfold(s, t) / tbg_modulate_bwd(gs_init) / bias-gradient mode of
    tbg_bias_act_bwd against their documented semantics (fp32: 1e-4 relative, summation order only)."""
    import emu
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(B + I + O)
    s = torch.randn(B, I, generator=gen) + 1.0
    q = torch.rand(I, O, generator=gen) / I
    d = K.demod_coef(s.to(DEV), q.to(DEV))
    rd = emu.emu_demod_coef(s, q)
    assert rel_err(d, rd) < 1e-5
    S1, Spre, Snz = (torch.randn(B, O, generator=gen) for _ in range(3))
    ns = torch.randn(1, generator=gen)
    bias = torch.randn(O, generator=gen)
    out = K.demod_bwd(*(t.to(DEV) for t in (S1, Spre, Snz, rd, ns, bias, s, q)))
    ref = emu.emu_demod_bwd(S1, Spre, Snz, rd, ns, bias, s, q)
    for a, b in zip(out, ref):
        assert rel_err(a, b) < 1e-4
    if O % 32 == 0 and I % 32 == 0:
        spec = C.weight_spec("plain", 8, 16, I, O, 3, True)
        w = torch.randn(3, 3, I, O, generator=gen)
        g = torch.randn(spec.fwd_rows, spec.fwd_cols, generator=gen)
        t = ref[0]
        got = K.wfold(g.to(DEV), spec, w_raw=w.to(DEV), s=s.to(DEV), t=t.to(DEV))
        want = emu.emu_wfold(g, spec, w_raw=w, s=s, t=t)
        assert rel_err(got, want) < 1e-4
    if O % 8 == 0:
        x = _bf16_round(torch.randn(B, 4, 8, O, generator=gen))
        gy = _bf16_round(torch.randn(B, 4, 8, O, generator=gen))
        so = torch.randn(B, O, generator=gen)
        init = torch.randn(B, O, generator=gen)
        gx, gs = K.modulate_bwd(gy.to(DEV).bfloat16(), x.to(DEV).bfloat16(), so.to(DEV), gs_init=init.to(DEV).clone())
        rgx, rgs = emu.emu_modulate_bwd(gy, x, so, gs_init=init)
        assert rel_err(gx.float(), rgx) < 1e-2 and rel_err(gs, rgs) < 1e-4
        o2 = K.bias_act_bwd(gy.to(DEV).bfloat16(), x.to(DEV).bfloat16(), act=True, gain=1.0, want_sums=False,
                            bias_grad_only=True)
        r2 = emu.emu_bias_act_bwd(gy, x, act=True, gain=1.0, want_sums=False, bias_grad_only=True)
        assert rel_err(o2[0].float(), r2[0]) < 1e-2 and rel_err(o2[1], r2[1]) < 1e-4 and o2[2] is None


@pytest.mark.parametrize("B,n,S,Is,idxs", [(4, 9, 128, (128, 512, 512, 3), (0, 0, 1, 2)),
                                           (32, 12, 512, (128, 128, 512, 512, 256, 64), (0, 0, 1, 2, 5, 11)),
                                           (37, 3, 96, (40, 200), (2, 0))])
def test_grouped_style_projection_vs_emulated_semantics(B, n, S, Is, idxs):
    """tbg_style_dense_fwd/bwd (all style projections of the synthesis network in one launch) against
    their documented semantics; fp32, 1e-4 relative (summation order only)."""
    import emu
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(B + S)
    style = torch.randn(B, n, S, generator=gen)
    ws = [torch.randn(S, I, generator=gen) for I in Is]
    bs = [torch.randn(I, generator=gen) for I in Is]
    gss = [torch.randn(B, I, generator=gen) for I in Is]
    coef = 1.0 / math.sqrt(S)
    dv = lambda ts: [t.to(DEV) for t in ts]
    out = K.style_dense_fwd(style.to(DEV), dv(ws), dv(bs), idxs, coef)
    ref = emu.emu_style_dense_fwd(style, ws, bs, idxs, coef)
    for a, b in zip(out, ref):
        assert rel_err(a, b) < 1e-4
    gstyle, gws, gbs = K.style_dense_bwd(style.to(DEV), dv(ws), dv(gss), idxs, coef)
    rstyle, rws, rbs = emu.emu_style_dense_bwd(style, ws, gss, idxs, coef)
    assert rel_err(gstyle, rstyle) < 1e-4
    for a, b in zip(gws + gbs, rws + rbs):
        assert rel_err(a, b) < 1e-4


@pytest.mark.parametrize("B,h,w,I,O", [(3, 4, 16, 128, 64), (32, 2, 8, 128, 512), (5, 16, 64, 64, 128), (2, 8, 32, 256, 256)])
def test_unfolded_upsample_conv_kernels_vs_emulated_semantics(B, h, w, I, O):
    """upsample_conv_2d as transposed conv (tap masks, (h+1) x (w+1) non-power-of-two tile grid) + tbg_fir4:
    every kernel against its documented semantics, and the whole layer (forward and all gradients)
    against the FIR-folded formulation of the same layer."""
    import emu
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import fused as F
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(B * 7 + h + I)
    spec = C.weight_spec("upT", h, w, I, O, 3, True, "modconv")
    x = _bf16_round(torch.randn(B, h, w, I, generator=gen))
    wr = torch.randn(3, 3, I, O, generator=gen)
    fwd, adj, q = K.wprep(wr.to(DEV), spec, want_adj=True, want_q=True)
    rf, ra, rq = emu.emu_wprep(wr, spec, want_adj=True, want_q=True)
    assert rel_err(fwd.float(), rf.float()) < 4e-3 and rel_err(adj.float(), ra.float()) < 4e-3
    # transposed conv with masked taps, fp32 out: accumulation order only
    T = K.conv2d_igemm(x.to(DEV).bfloat16(), fwd, out_fp32=True, **spec.fwd_kwargs)
    rT = emu.emu_conv2d_igemm(x, fwd.float().cpu(), out_fp32=True, **spec.fwd_kwargs)
    assert T.shape == (B, 2 * h + 2, 2 * w + 2, O) and rel_err(T, rT) < 2e-5
    # FIR pass (+ epilogue) and its adjoint
    Tb = _bf16_round(rT)
    d = torch.rand(B, O, generator=gen) + 0.5
    nz = torch.randn(B, 2 * h, 2 * w, generator=gen)
    ns = torch.tensor([0.3])
    bias = torch.randn(O, generator=gen)
    kw = dict(d=d, noise=nz, noise_strength=ns, bias=bias, act=1, gain=math.sqrt(2))
    o = K.fir4(Tb.to(DEV).bfloat16(), spec.out_hw, (-1, -1), 1 / 16, **{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items()})
    ro = emu.emu_fir4(Tb, spec.out_hw, (-1, -1), 1 / 16, **kw)
    assert rel_err(o.float(), ro) < 1e-2
    gy = _bf16_round(torch.randn(B, 2 * h, 2 * w, O, generator=gen))
    gT = K.fir4(gy.to(DEV).bfloat16(), spec.t_hw, (-2, -2), 1 / 16)
    rgT = emu.emu_fir4(gy, spec.t_hw, (-2, -2), 1 / 16)
    assert rel_err(gT.float(), rgT) < 1e-2
    # adjoint identity <fir(T), gy> == <T, fir_adj(gy)> on the emulated semantics (fp64)
    lhs = (emu.emu_fir4(Tb.double(), spec.out_hw, (-1, -1), 1 / 16).double() * gy.double()).sum()
    rhs = (Tb.double() * emu.emu_fir4(gy.double(), spec.t_hw, (-2, -2), 1 / 16).double()).sum()
    assert abs(lhs - rhs) < 1e-6 * (abs(lhs) + 1)
    # adjoint-layout fold
    g = torch.randn(spec.adj_rows, spec.adj_cols, generator=gen)
    s = torch.randn(B, I, generator=gen)
    t = torch.randn(B, O, generator=gen)
    got = K.wfold_adj(g.to(DEV), spec, w_raw=wr.to(DEV), s=s.to(DEV), t=t.to(DEV), flip=True)
    want = emu.emu_wfold_adj(g, spec, w_raw=wr, s=s, t=t, flip=True)
    assert rel_err(got, want) < 1e-4
    # whole layer: unfolded vs folded formulation on the GPU (bf16 activations: 2e-2 of the tensor maximum)
    sc = torch.randn(B, I, generator=gen) * 0.2 + 1.0
    outs = []
    for kind, Fn in (("up", F.ModConvAct), ("upT", F.ModUpConvAct)):
        F.clear_step_cache()
        xa = x.to(DEV).bfloat16().requires_grad_(True)
        sa = sc.to(DEV).requires_grad_(True)
        wa = wr.to(DEV).requires_grad_(True)
        y = Fn.apply(xa, sa, wa, nz.to(DEV), ns.to(DEV).reshape(()), bias.to(DEV), C.weight_spec(kind, h, w, I, O, 3, True, "modconv"),
                     math.sqrt(2))
        y.backward(gy.to(DEV).bfloat16())
        outs.append((y.float(), xa.grad.float(), sa.grad, wa.grad))
    F.clear_step_cache()
    # outputs: one bf16 rounding (2e-2 of the maximum).  Gradients: the leaky-ReLU slope is recovered from
    # the bf16 output, so the two formulations pick different slopes where |pre-activation| is below their
    # rounding difference; compare in relative L2 norm (5e-2).
    assert rel_err(outs[0][0], outs[1][0]) < 2e-2
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert _rel_l2(a, b) < 5e-2


@pytest.mark.parametrize("B,H,W,I,O,k,rh", [(4, 16, 64, 128, 128, 3, True), (6, 8, 32, 256, 256, 3, False),
                                             (4, 16, 64, 128, 256, 1, True), (3, 8, 16, 256, 512, 1, False),
                                             (64, 4, 8, 512, 512, 3, False)])
def test_unfolded_downsample_conv_layer_matches_folded(B, H, W, I, O, k, rh):
    """conv_downsample_2d as FIR pre-pass + strided k x k conv (forward, weight gradient) against the
    FIR-folded single convolution, whole layer incl. bias/lrelu/residual (bf16 activations: 2e-2)."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import fused as F

    gen = torch.Generator().manual_seed(B + H + I + k)
    x = _bf16_round(torch.randn(B, H, W, I, generator=gen))
    wr = torch.randn(k, k, I, O, generator=gen)
    bias = torch.randn(O, generator=gen) * 0.1 if k == 3 else None
    oh, ow = (H // 2 if rh else H), W // 2
    res = _bf16_round(torch.randn(B, oh, ow, O, generator=gen)) if k == 3 else None
    gy = _bf16_round(torch.randn(B, oh, ow, O, generator=gen))
    outs = []
    for kind in ("down", "downU"):
        F.clear_step_cache()
        xa = x.to(DEV).bfloat16().requires_grad_(True)
        wa = wr.to(DEV).requires_grad_(True)
        ba = bias.to(DEV).requires_grad_(True) if bias is not None else None
        ra = res.to(DEV).bfloat16().requires_grad_(True) if res is not None else None
        y = F.ConvAct.apply(xa, wa, ba, ra, C.weight_spec(kind, H, W, I, O, k, rh, "dconv"), 1.0)
        assert y.shape == (B, oh, ow, O)
        y.backward(gy.to(DEV).bfloat16())
        outs.append([y.float(), xa.grad.float(), wa.grad] + ([ba.grad, ra.grad.float()] if k == 3 else []))
    F.clear_step_cache()
    assert rel_err(outs[0][0], outs[1][0]) < 2e-2
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert _rel_l2(a, b) < 5e-2          # see test_unfolded_upsample_conv_kernels_vs_emulated_semantics


def test_halo_kernel_store_paths_and_pipeline_depths_are_bit_identical():
    """The halo kernel's tuning switches (staged / direct epilogue stores, number of halo and weight boxes in flight) change
    the schedule, never the arithmetic: every configuration produces bit-identical outputs."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib

    def run():
        outs = []
        for (B, H, W, I, O) in [(3, 16, 64, 64, 64), (2, 32, 128, 128, 128), (5, 16, 32, 256, 96)]:
            g = C.plain_geom(H, W, I, O, 3)
            gen = torch.Generator().manual_seed(B + H)
            x = torch.randn(B, H, W, I, generator=gen).to(DEV).bfloat16()
            w = (torch.randn(g.n_total, g.k_total, generator=gen) / g.k_total ** 0.5).to(DEV).bfloat16()
            bias = torch.randn(O, generator=gen).to(DEV)
            outs.append(K.conv2d_igemm(x, w, bias=bias, act=1, act_gain=1.4, **g.kernel_kwargs()).float().cpu())
        return outs

    keys = ("conv_halo", "halo_staged", "halo_a_stages", "halo_b_stages")
    saved = {k: lib.get_tuning(k) for k in keys}
    try:
        lib.set_tuning("conv_halo", 1)
        res = {}
        for staged, a_st, b_st in ((1, 2, 4), (0, 2, 4), (0, 3, 3), (1, 2, 2), (0, 2, 8)):
            lib.set_tuning("halo_staged", staged)
            lib.set_tuning("halo_a_stages", a_st)
            lib.set_tuning("halo_b_stages", b_st)
            res[(staged, a_st, b_st)] = run()
    finally:
        for k, v in saved.items():
            lib.set_tuning(k, v)
    ref = res[(1, 2, 4)]
    for key, outs in res.items():
        for a, b in zip(ref, outs):
            assert torch.equal(a, b), key


@pytest.mark.parametrize("B,H,W,I,O", [(2, 16, 16, 64, 32), (3, 32, 64, 128, 128), (2, 16, 64, 256, 256),
                                       (1, 64, 256, 64, 64), (5, 16, 32, 64, 96)])
def test_conv3x3_halo_kernel_vs_emulated_semantics_and_igemm(B, H, W, I, O):
    """The halo-reuse kernel behind tbg_conv2d_igemm (tuning conv_halo = 1: one activation box per 64-channel block
    shared by the nine taps) against the documented semantics of the entry point and against conv_igemm_kernel on the
    same arguments: full modulated-conv epilogue (demodulation scale, noise, bias, leaky-ReLU) and the bare product."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib

    g = C.plain_geom(H, W, I, O, 3)
    gen = torch.Generator().manual_seed(B * 1000 + H + O)
    x = _bf16_round(torch.randn(B, H, W, I, generator=gen))
    w = _bf16_round(torch.randn(g.n_total, g.k_total, generator=gen) / math.sqrt(g.k_total))
    d = torch.rand(B, O, generator=gen) + 0.5
    nz = torch.randn(B, H, W, generator=gen)
    ns = torch.tensor([0.3])
    bias = torch.randn(O, generator=gen) * 0.2
    epi = dict(col_scale=d, noise=nz, noise_strength=ns, bias=bias, act=1, act_gain=math.sqrt(2.0))
    dev = lambda t: t.to(DEV)
    saved = lib.get_tuning("conv_halo")
    try:
        got = {}
        for halo in (1, 0):
            lib.set_tuning("conv_halo", halo)
            full = K.conv2d_igemm(dev(x).bfloat16(), dev(w).bfloat16(), **g.kernel_kwargs(),
                                  **{k: (dev(v) if torch.is_tensor(v) else v) for k, v in epi.items()})
            bare = K.conv2d_igemm(dev(x).bfloat16(), dev(w).bfloat16(), **g.kernel_kwargs())
            got[halo] = (full.float().cpu(), bare.float().cpu())
    finally:
        lib.set_tuning("conv_halo", saved)
    want_full = emu_conv2d_igemm(x, w, **g.kernel_kwargs(), **epi).float()
    want_bare = emu_conv2d_igemm(x, w, **g.kernel_kwargs()).float()
    for halo in (1, 0):
        assert rel_err(got[halo][0], want_full) < 1e-2, halo          # bf16 output
        assert rel_err(got[halo][1], want_bare) < 1e-2, halo
    # same bf16 inputs, fp32 accumulation in a different order: the two kernels agree to output rounding
    assert rel_err(got[1][0], got[0][0]) < 1e-2 and rel_err(got[1][1], got[0][1]) < 1e-2


def test_conv3x3_halo_kernel_phase_scatter_and_tap_masks():
    """Up-sampling geometry on the halo kernel: columns = (phase_y, phase_x, cout), phase (py,px) of row (b,i,j) goes
    to pixel (2i+py, 2j+px); per-phase tap masks skip weight blocks (treated as zero) — against the emulated semantics."""
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib

    B, H, W, I, O = 2, 16, 32, 64, 64
    gen = torch.Generator().manual_seed(7)
    x = _bf16_round(torch.randn(B, H, W, I, generator=gen))
    w = _bf16_round(torch.randn(4 * O, 9 * I, generator=gen) / math.sqrt(9 * I))
    bias = torch.randn(O, generator=gen) * 0.1
    kw = dict(Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), stride=(1, 1), up=(1, 1))
    masks = (0b111111111, 0b000111111, 0b110110110, 0b000010000)
    saved = lib.get_tuning("conv_halo")
    try:
        lib.set_tuning("conv_halo", 1)
        for tm in (None, masks):
            y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), bias=bias.to(DEV), act=1, act_gain=1.25,
                               tap_mask=tm, **kw)
            want = emu_conv2d_igemm(x, w, bias=bias, act=1, act_gain=1.25, tap_mask=tm, **kw)
            assert y.shape == (B, 2 * H, 2 * W, O)
            assert rel_err(y.float().cpu(), want.float()) < 1e-2, tm
        # 64 channels per phase: one 128-column tile holds two phases when their tap masks agree (pairs (0,1), (2,3)),
        # with per-sample column scales and per-output-pixel noise; also the width-only up-sampling geometry
        for up in ((1, 1), (0, 1)):
            nph = (1 + up[0]) * (1 + up[1])
            kw2 = dict(kw, up=up)
            w2 = _bf16_round(torch.randn(nph * O, 9 * I, generator=gen) / math.sqrt(9 * I))
            scale = torch.rand(B, O, generator=gen) + 0.5
            noise = torch.randn(B, H * (1 + up[0]), W * (1 + up[1]), generator=gen)
            ns = torch.tensor([0.3])
            for tm in (None, (0b101111101, 0b101111101, 0b000111111, 0b000111111)[:nph] + (0,) * (4 - nph)):
                y = K.conv2d_igemm(x.to(DEV).bfloat16(), w2.to(DEV).bfloat16(), bias=bias.to(DEV), col_scale=scale.to(DEV),
                                   noise=noise.to(DEV), noise_strength=ns.to(DEV), act=1, act_gain=1.25, tap_mask=tm, **kw2)
                want = emu_conv2d_igemm(x, w2, bias=bias, col_scale=scale, noise=noise, noise_strength=ns, act=1,
                                        act_gain=1.25, tap_mask=tm, **kw2)
                assert rel_err(y.float().cpu(), want.float()) < 1e-2, (up, tm)
    finally:
        lib.set_tuning("conv_halo", saved)


@pytest.mark.parametrize("Ho,Wo,B", [(5, 17, 3), (17, 65, 2), (3, 9, 32), (33, 129, 1), (7, 5, 9)])
def test_conv_forward_non_power_of_two_grids(Ho, Wo, B):
    """Tile boxes are chosen per grid (any bw x bh x bn <= 128): plain 3x3 SAME conv on odd-sized grids."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K

    g = C.plain_geom(Ho, Wo, 64, 96, 3)
    gen = torch.Generator().manual_seed(Ho * 100 + Wo)
    x = _bf16_round(torch.randn(B, Ho, Wo, 64, generator=gen))
    w = _bf16_round(torch.randn(g.n_total, g.k_total, generator=gen) / math.sqrt(g.k_total))
    y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), out_fp32=True, **g.kernel_kwargs())
    assert rel_err(y, emu_conv2d_igemm(x, w, out_fp32=True, **g.kernel_kwargs())) < 2e-5


@pytest.mark.parametrize("B,H,W,C", [(4, 16, 64, 128), (3, 5, 7, 64), (64, 32, 128, 128), (2, 64, 256, 64)])
def test_fromrgb_kernels_vs_emulated_semantics(B, H, W, C):
    """tbg_fromrgb_fwd / bwd against their documented semantics (bf16 output 1e-2; fp32 gradients 1e-4)."""
    import emu
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(B + H + C)
    img = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
    w = torch.randn(3, C, generator=gen)
    bias = torch.randn(C, generator=gen) * 0.1
    coef, gain = 1 / math.sqrt(3.0), math.sqrt(2.0)
    out = K.fromrgb_fwd(img.to(DEV), w.to(DEV), bias.to(DEV), coef, gain)
    with emu.emulated_kernels(act_dtype=torch.float32):
        ref = emu.emu_fromrgb_fwd(img, w, bias, coef, gain)
    assert rel_err(out.float(), ref) < 1e-2
    g = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    ob = out.float().cpu()
    gi, gw, gb = K.fromrgb_bwd(img.to(DEV), w.to(DEV), g.to(DEV).bfloat16(), out, coef, gain)
    ri, rw, rb = emu.emu_fromrgb_bwd(img, w, g, ob, coef, gain)
    assert rel_err(gi, ri) < 1e-4 and rel_err(gw, rw) < 1e-4 and rel_err(gb, rb) < 1e-4
    gi2, gw2, gb2 = K.fromrgb_bwd(img.to(DEV), w.to(DEV), g.to(DEV).bfloat16(), out, coef, gain, want_w=False)
    assert gw2 is None and gb2 is None and rel_err(gi2, ri) < 1e-4


def test_fromrgb_fwd_channel_counts_outside_the_ladders():
    """C / 8 that does not divide the block (24, 40 channels) takes the generic kernel: same semantics."""
    import emu
    from textboxgan_b200 import kernels as K

    gen = torch.Generator().manual_seed(11)
    for C in (24, 40, 8):
        img = torch.rand(3, 3, 9, 13, generator=gen) * 2 - 1
        w = torch.randn(3, C, generator=gen)
        bias = torch.randn(C, generator=gen) * 0.1
        out = K.fromrgb_fwd(img.to(DEV), w.to(DEV), bias.to(DEV), 0.5, 1.4)
        with emu.emulated_kernels(act_dtype=torch.float32):
            ref = emu.emu_fromrgb_fwd(img, w, bias, 0.5, 1.4)
        assert rel_err(out.float(), ref) < 1e-2


@pytest.mark.parametrize("B,H,W,mcn,cw", [(5, 16, 64, 8, 8), (4, 32, 128, 8, 16), (3, 64, 256, 12, "64/3")])
def test_crop_resize_kernels_vs_emulated_semantics(B, H, W, mcn, cw):
    """tbg_crop_resize_fwd / bwd (convert_inputs) against the crop + tf.image.resize semantics (fp32, 1e-5), incl.
    words without a blank label, a blank in the first position (clamped to one column) and a fractional char_width."""
    from fractions import Fraction

    import emu
    from textboxgan_b200 import kernels as K

    cw = Fraction(cw)
    gen = torch.Generator().manual_seed(B + H)
    img = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
    labels = torch.randint(2, 96, (B, mcn), generator=gen, dtype=torch.int32)
    labels[0, 3:] = 1
    labels[1, :] = 1
    if B > 3:
        labels[3, mcn - 1:] = 1
    out = K.crop_resize_fwd(img.to(DEV), labels.to(DEV), 1, cw, (64, 256))
    ref = emu.emu_crop_resize_fwd(img, labels, 1, cw, (64, 256))
    assert rel_err(out, ref) < 1e-5
    g = torch.randn(B, 64, 256, 3, generator=gen)
    gi = K.crop_resize_bwd(g.to(DEV), labels.to(DEV), 1, cw, (H, W))
    ri = emu.emu_crop_resize_bwd(g, labels, 1, cw, (H, W))
    assert rel_err(gi, ri) < 1e-4
    # adjoint identity on the device results
    lhs, rhs = (out.cpu().double() * g.double()).sum(), (img.double() * gi.cpu().double()).sum()
    assert abs(lhs - rhs) < 1e-4 * (abs(lhs) + 1)


def test_relu_mask_epilogue_and_fused_encoder_backward():
    """relu_mask epilogue of tbg_conv2d_igemm (fused ReLU backward) against its documented semantics, and the
    one-node ResNet encoder (masks and residual sums folded into the input-gradient convs) against the
    layer-by-layer autograd formulation (bf16 activations, identical kernels: 2e-2)."""
    from textboxgan_b200 import aster_inferer as AI
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200.config import baseline_config

    g = C.ConvGeom(8, 32, 128, 64, C.Axis("s2", 1, 0), C.Axis("s1", 1, 0))
    gen = torch.Generator().manual_seed(11)
    x = _bf16_round(torch.randn(3, 8, 32, 128, generator=gen))
    w = _bf16_round(torch.randn(g.n_total, g.k_total, generator=gen) / math.sqrt(g.k_total))
    res = _bf16_round(torch.randn(3, 4, 32, 64, generator=gen))
    m = _bf16_round(torch.randn(3, 4, 32, 64, generator=gen))
    kw = dict(g.kernel_kwargs(), res_scale=1.0, res_first=True, out_fp32=True)
    y = K.conv2d_igemm(x.to(DEV).bfloat16(), w.to(DEV).bfloat16(), residual=res.to(DEV).bfloat16(),
                       relu_mask=m.to(DEV).bfloat16(), **kw)
    ref = emu_conv2d_igemm(x, w, residual=res, relu_mask=m, **kw)
    assert rel_err(y, ref) < 2e-5 and float((y.cpu()[m <= 0]).abs().max()) == 0.0

    cfg = baseline_config(0)
    aster = AI.AsterInferer(cfg, device=DEV, synthetic_weights=True)
    img = torch.randn(4, 64, 256, 3, generator=gen).to(DEV)
    gm = torch.randn(4, 64, 512, generator=gen).to(DEV)       # T = 256 / 4 feature columns
    outs = []
    for fused in (False, True):
        AI.FUSED_ENCODER = fused
        xi = img.clone().requires_grad_(True)
        mem = aster._encoder(xi)
        (gi,) = torch.autograd.grad((mem * gm[:, : mem.shape[1]]).sum(), xi)
        outs.append((mem, gi))
    AI.FUSED_ENCODER = True
    assert rel_err(outs[0][0], outs[1][0]) < 1e-6          # same forward kernels
    assert _rel_l2(outs[1][1], outs[0][1]) < 2e-2


@pytest.mark.parametrize("B,T,steps", [(3, 32, 8), (5, 17, 4)])
def test_attention_decoder_and_lstm_kernels_vs_emulated_semantics(B, T, steps):
    import emu
    from textboxgan_b200 import kernels as K
    from textboxgan_b200.aster_inferer import _LstmLayer, _pack_decoder, init_aster_params

    P = init_aster_params()
    w_dev = _pack_decoder(P, DEV)
    w_cpu = {k: v.float().cpu() for k, v in w_dev.items()}          # the same bf16-rounded weights
    gen = torch.Generator().manual_seed(B * 100 + T)
    mem = torch.randn(B, T, 512, generator=gen)
    keys = mem @ P["dec/memory_layer/w"]
    logits, sv = K.attn_decoder_fwd(mem.to(DEV), keys.to(DEV), w_dev, steps)
    rl, rsv = emu.emu_attn_decoder_fwd(mem, keys, w_cpu, steps)
    assert torch.equal(sv["prev"].cpu(), rsv["prev"])                # greedy symbols: integer path, bit-exact
    assert rel_err(logits, rl) < 1e-4
    for k in ("a", "ctx", "gates", "c", "h"):
        assert rel_err(sv[k], rsv[k]) < 1e-4, k
    gl = torch.randn(B, steps, 96, generator=gen)
    gm, gk = K.attn_decoder_bwd(mem.to(DEV), keys.to(DEV), w_dev, gl.to(DEV), sv)
    rgm, rgk = emu.emu_attn_decoder_bwd(mem, keys, w_cpu, gl, rsv)
    assert rel_err(gm, rgm) < 1e-4 and rel_err(gk, rgk) < 1e-4
    # LSTM sequence kernels
    lay = _LstmLayer(P, "rnn/l0", DEV)
    xp = torch.randn(2, B, T, 1024, generator=gen)
    h, gates, c = K.lstm_seq_fwd(xp.to(DEV), lay.w_packed)
    rh, rg, rc = emu.emu_lstm_seq_fwd(xp, lay.w_packed.float().cpu())
    assert rel_err(h, rh) < 1e-4 and rel_err(gates, rg) < 1e-4 and rel_err(c, rc) < 1e-4
    gh = torch.randn(2, B, T, 256, generator=gen)
    gxp = K.lstm_seq_bwd(gh.to(DEV), gates, c, lay.wT_packed)
    rgxp = emu.emu_lstm_seq_bwd(gh, rg, rc, lay.wT_packed.float().cpu())
    assert rel_err(gxp, rgxp) < 1e-4


def _product(cfg, GP, DP, with_ocr=True):
    from textboxgan_b200.aster_inferer import AsterInferer
    from textboxgan_b200.discriminator import Discriminator
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.optimizers import Adam, update_optimizer_params
    from textboxgan_b200.training_step import TrainingStep

    G = Generator(cfg, device=DEV, seed=0)
    G.load_state_dict(GP)
    D = Discriminator(cfg, device=DEV, seed=0)
    D.load_state_dict(DP)
    aster = AsterInferer(cfg, device=DEV, synthetic_weights=True) if with_ocr else None
    go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=DEV), cfg)
    return G, D, aster, ts


def _to_dev(draws):
    return {k: (v.to(DEV) if torch.is_tensor(v) else ([t.to(DEV) for t in v] if isinstance(v, list) else v))
            for k, v in draws.items()}


def test_generator_and_discriminator_forward_vs_oracle():
    cfg = small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    real, words, labels = OT.synthetic_batch(cfg, 4, g)
    draws = OT.make_draws(cfg, 4, g)
    G, D, _, _ = _product(cfg, GP, DP, with_ocr=False)
    ref = OS.generator(words, draws["z"], GP, cfg, training=True, draws=draws)
    for grad_mode in (True, False):           # differentiable path and fused-epilogue (no_grad) path
        G.load_state_dict(GP)
        with torch.set_grad_enabled(grad_mode):
            img = G((words.to(DEV), draws["z"].to(DEV)), training=True, draws=_to_dev(draws))
        # bf16 activations through 7 modulated convs: 3e-2 of the image range
        assert rel_err(img, ref) < 3e-2, (grad_mode, rel_err(img, ref))
        sc = D(real.to(DEV)) if grad_mode else None
        with torch.set_grad_enabled(grad_mode):
            sc = D(real.to(DEV))
        rsc = OS.discriminator(real, DP, cfg)
        assert (sc.cpu() - rsc).abs().max() < 3e-2 * max(1.0, float(rsc.abs().max()))


@pytest.mark.parametrize("do_r1,do_pl,with_ocr", [(False, False, True), (True, True, False)])
def test_train_step_vs_oracle(do_r1, do_pl, with_ocr):
    """Whole _train_step on the GPU vs the oracle at BASELINE configs[0] (batch 4, 64x16): the seven
    losses within 5e-2 relative (bf16 activations), updated weights within the Adam step size."""
    cfg = small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    real, words, labels = OT.synthetic_batch(cfg, 4, g)
    draws = OT.make_draws(cfg, 4, g, with_pl=do_pl)
    st = OT.StepState(copy.deepcopy(GP), copy.deepcopy(DP), OA.init_aster_params(), OT.make_adam(cfg.g_opt),
                      OT.make_adam(cfg.g_opt), OT.make_adam(cfg.d_opt), torch.zeros(()))
    ref_out, ref_grads, _ = OT.train_step(st, cfg, real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws,
                                          fused=False, with_ocr=with_ocr, ret_grads=True)
    G, D, aster, ts = _product(cfg, GP, DP, with_ocr)
    d2 = _to_dev(draws)
    d2["keep_grads"] = True
    out = ts.dist_train_step(real.to(DEV), torch.zeros((), device=DEV), words.to(DEV), labels.to(DEV), do_r1, do_pl,
                             1e-4, draws=d2)
    flat = lambda o: [float(v) for v in (*o[0], *o[1], o[2])]
    got, want = flat(out), flat(ref_out)
    for a, b in zip(got, want):
        assert abs(a - b) <= 5e-2 * max(1.0, abs(b)), (got, want)
    # gradient direction of the big groups agrees (cosine similarity; bf16 + lrelu sign flips)
    for names, grads, ref in ((ts._g_names, ts.last_grads[0], ref_grads[0]), (ts._d_names, ts.last_grads[2], ref_grads[2])):
        a = torch.cat([gr.reshape(-1).float().cpu() for n, gr in zip(names, grads) if gr is not None and n in ref])
        b = torch.cat([ref[n].reshape(-1) for n, gr in zip(names, grads) if gr is not None and n in ref])
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.98, cos
    # every weight moved by at most ~lr (Adam with beta1 = 0) and in the oracle's direction on average
    lr = 0.002
    moved = same = 0
    for n, p in G.params.items():
        if n in OS.NON_TRAINABLE:
            continue
        d_prod = p.detach().cpu() - GP[n]
        d_ref = st.G[n] - GP[n]
        assert float(d_prod.abs().max()) <= 2.5 * lr * 2 + 1e-6, n
        moved += d_ref.numel()
        same += int(((d_prod * d_ref) > 0).sum()) + int(((d_prod == 0) & (d_ref == 0)).sum())
    assert same / moved > 0.9


def test_launches_are_counted_and_library_is_loaded():
    from textboxgan_b200 import lib

    h = lib.load()
    h.tbg_reset_launch_count()
    cfg = small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    G, D, _, _ = _product(cfg, GP, DP, with_ocr=False)
    real, words, labels = OT.synthetic_batch(cfg, 4, g)
    with torch.no_grad():
        G((words.to(DEV), torch.randn(4, cfg.z_dim, device=DEV)))
    torch.cuda.synchronize()
    assert h.tbg_launch_count() >= 6
    maps = open("/proc/self/maps").read()
    assert "libtbg.so" in maps


@pytest.mark.parametrize("B,H,W,I,O", [(8, 32, 64, 128, 128), (4, 16, 64, 256, 256), (16, 32, 32, 64, 64),
                                       (6, 16, 128, 192, 96)])
def test_wgrad_halo_kernel_vs_emulated_semantics_and_wgrad(B, H, W, I, O):
    """The halo-reuse weight-gradient kernel behind tbg_conv2d_wgrad (tuning wgrad_halo = 1: one x halo box per 64-channel
    block shared by the taps of a work item, MN-major shifted windows) against the documented semantics of the entry
    point and against conv_wgrad_kernel on the same arguments; accumulation into a non-zero gw."""
    from textboxgan_b200 import conv as C
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib

    g = C.plain_geom(H, W, I, O, 3)
    gen = torch.Generator().manual_seed(B * 100 + H + I)
    x = _bf16_round(torch.randn(B, H, W, I, generator=gen))
    gy = _bf16_round(torch.randn(B, H, W, O, generator=gen))
    init = torch.randn(O, 9 * I, generator=gen)
    want = emu_conv2d_wgrad(x, gy, **g.kernel_kwargs()).double() + init.double()
    saved = lib.get_tuning("wgrad_halo")
    try:
        got = {}
        for halo in (1, 0):
            lib.set_tuning("wgrad_halo", halo)
            gw = init.clone().to(DEV)
            K.conv2d_wgrad(x.to(DEV).bfloat16(), gy.to(DEV).bfloat16(), gw=gw, **g.kernel_kwargs())
            got[halo] = gw.cpu()
    finally:
        lib.set_tuning("wgrad_halo", saved)
    for halo in (1, 0):
        assert rel_err(got[halo], want) < 2e-5, halo            # fp32 accumulation of exact bf16 products
    assert rel_err(got[1], got[0]) < 2e-5


@pytest.mark.parametrize("B,H,W,I,O,up", [(2, 16, 32, 64, 128, (0, 0)), (3, 32, 64, 128, 128, (0, 0)), (4, 16, 16, 128, 256, (0, 0)),
                                          (2, 16, 32, 128, 128, (1, 1)), (2, 16, 32, 64, 64, (1, 1)), (64, 16, 64, 128, 128, (0, 0))])
def test_halo_kernel_cta_pairs_bit_identical(B, H, W, I, O, up):
    """conv3x3_halo_kernel<CTA2>: two CTAs of a cluster issue one tcgen05.mma.cta_group::2 over M = 256 rows, each staging
    half of every weight box.  Same summation order per output element as the single-CTA kernel: bit-identical results,
    with the full epilogue, on plain and 4-phase (up) geometries, few and many pair items."""
    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib

    gen = torch.Generator().manual_seed(B * 7 + H + O)
    nph = (1 + up[0]) * (1 + up[1])
    x = _bf16_round(torch.randn(B, H, W, I, generator=gen)).to(DEV).bfloat16()
    w = _bf16_round(torch.randn(nph * O, 9 * I, generator=gen) / math.sqrt(9 * I)).to(DEV).bfloat16()
    epi = dict(col_scale=(torch.rand(B, O, generator=gen) + 0.5).to(DEV),
               noise=torch.randn(B, H * (1 + up[0]), W * (1 + up[1]), generator=gen).to(DEV),
               noise_strength=torch.tensor([0.3], device=DEV), bias=(torch.randn(O, generator=gen) * 0.1).to(DEV),
               act=1, act_gain=1.4)
    kw = dict(Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), stride=(1, 1), up=up)
    saved = lib.get_tuning("halo_cta2")
    got = {}
    try:
        for v in (0, 1):
            lib.set_tuning("halo_cta2", v)
            got[v] = (K.conv2d_igemm(x, w, **kw).float().cpu(), K.conv2d_igemm(x, w, **kw, **epi).float().cpu())
    finally:
        lib.set_tuning("halo_cta2", saved)
    torch.cuda.synchronize()
    assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1])
    want = emu_conv2d_igemm(x.float().cpu(), w.float().cpu(), **kw)
    assert rel_err(got[1][0], want.float()) < 1e-2
