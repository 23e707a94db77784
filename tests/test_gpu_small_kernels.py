"""GPU parity of the small kernels behind the word encoder, the mapping network, the discriminator head, the RGB branch and
the generic resampling op — every call goes through the C ABI (ctypes) and is compared with the documented semantics in
tests/emu.py (fp64) or with the oracle's literal restatement of the reference."""
import math

import pytest
import torch

import emu
from common import rel_err
from oracle import stylegan as OS

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,K,N,act,bias", [(128, 512, 512, 1, True), (64, 8192, 512, 1, True), (5, 37, 3, 0, False),
                                            (768, 32, 256, 2, True), (128, 512, 1, 0, True), (33, 70, 65, 1, True),
                                            (128, 4096, 64, 1, True), (200, 300, 130, 2, True)])
def test_dense_kernels_vs_emulated_semantics(M, K, N, act, bias):
    """tbg_dense_fwd / tbg_dense_bwd (Dense.call dense.py:23-29 + bias + activation), exact fp32: 1e-5."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=gen)
    w = torch.randn(K, N, generator=gen)
    b = torch.randn(N, generator=gen) if bias else None
    gy = torch.randn(M, N, generator=gen)
    kw = dict(coef=1.0 / math.sqrt(K) * 0.7, bias_coef=0.3, act=act, gain=math.sqrt(2.0) if act else 1.0)
    y = Kn.dense_fwd(x.to(DEV), w.to(DEV), b.to(DEV) if bias else None, **kw)
    want = emu.emu_dense_fwd(x, w, b, **kw)
    assert rel_err(y, want) < 1e-5
    gx, gw, gb = Kn.dense_bwd(x.to(DEV), w.to(DEV), y, gy.to(DEV), **kw, want_gb=bias)
    ex, ew, eb = emu.emu_dense_bwd(x, w, want, gy, **kw, want_gb=bias)
    assert rel_err(gx, ex) < 1e-5 and rel_err(gw, ew) < 1e-5
    if bias:
        assert rel_err(gb, eb) < 1e-5
    # pieces can be skipped
    gx2, gw2, gb2 = Kn.dense_bwd(x.to(DEV), w.to(DEV), y, gy.to(DEV), **kw, want_gx=False, want_gw=False, want_gb=False)
    assert gx2 is None and gw2 is None and gb2 is None


def test_pixel_norm_kernels():
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(3)
    for M, K in ((128, 512), (7, 128), (1, 33)):
        x = torch.randn(M, K, generator=gen)
        gy = torch.randn(M, K, generator=gen)
        assert rel_err(Kn.pixel_norm_fwd(x.to(DEV)), emu.emu_pixel_norm_fwd(x)) < 1e-6
        assert rel_err(Kn.pixel_norm_bwd(x.to(DEV), gy.to(DEV)), emu.emu_pixel_norm_bwd(x, gy)) < 1e-5


@pytest.mark.parametrize("B,mcn,fm0,with_mask", [(4, 8, 128, True), (3, 12, 192, True), (2, 16, 256, False)])
def test_word_encoder_kernels_vs_emulated_semantics(B, mcn, fm0, with_mask):
    """tbg_word_encoder_fwd / bwd (word_encoder.py:39-63): gather is exact (index arithmetic), dense fp32 1e-5, bf16 map."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(B + mcn)
    E, D = 32, 256
    words = torch.randint(0, 70, (B, mcn), generator=gen).to(torch.int32)
    words[0, -1] = 0
    w0 = torch.zeros(1, E)
    table = torch.randn(69, E, generator=gen)
    mask = (torch.rand(B, mcn, E, generator=gen) < 0.7).float() if with_mask else None
    fc_w = torch.randn(E, D, generator=gen) * 0.2
    fc_b = torch.randn(D, generator=gen) * 0.1
    hwc = (2, 8, fm0)
    d = lambda t: t.to(DEV) if t is not None else None
    out, emb, act = Kn.word_encoder_fwd(d(words), d(w0), d(table), d(mask), 0.7, d(fc_w), d(fc_b), hwc)
    eo, ee, ea = emu.emu_word_encoder_fwd(words, w0, table, mask, 0.7, fc_w, fc_b, hwc)
    assert out.shape == (B, 2, 8, fm0) and out.dtype == torch.bfloat16
    assert rel_err(emb, ee) < 1e-6 and rel_err(act, ea) < 1e-5
    assert rel_err(out.float(), eo.float()) < 1e-2
    g_out = _bf16_round(torch.randn(B, 2, 8, fm0, generator=gen))
    gt, gw, gb = Kn.word_encoder_bwd(d(words), d(mask), 0.7, d(fc_w), emb, act, d(g_out).bfloat16(), 69, hwc)
    et, ew, eb = emu.emu_word_encoder_bwd(words, mask, 0.7, fc_w, ee, ea, g_out, 69, hwc)
    assert rel_err(gt, et) < 1e-5 and rel_err(gw, ew) < 1e-5 and rel_err(gb, eb) < 1e-5


@pytest.mark.parametrize("B,n_calls,C,H,W", [(8, 1, 512, 4, 4), (4, 2, 512, 4, 4), (2, 1, 64, 2, 3), (64, 2, 512, 4, 4)])
def test_minibatch_std_kernels_vs_emulated_semantics(B, n_calls, C, H, W):
    """tbg_minibatch_std_fwd / bwd (mini_batch_std.py:10-35, one statistic per group of min(4, B) samples, per call)."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(B * 10 + n_calls)
    x = _bf16_round(torch.randn(B * n_calls, H, W, C, generator=gen))
    cpad = (C + 1 + 63) // 64 * 64
    xcat, stat = Kn.minibatch_std_fwd(x.to(DEV).bfloat16(), n_calls, cpad)
    ecat, estat = emu.emu_minibatch_std_fwd(x, n_calls, cpad)
    assert rel_err(stat, estat) < 1e-5
    assert torch.equal(xcat[..., :C].float().cpu(), x)                       # the copy is exact
    assert rel_err(xcat[..., C].float(), ecat[..., C]) < 1e-2 and float(xcat[..., C + 1:].abs().max()) == 0.0
    gxcat = _bf16_round(torch.randn(B * n_calls, H, W, cpad, generator=gen))
    gx = Kn.minibatch_std_bwd(x.to(DEV).bfloat16(), gxcat.to(DEV).bfloat16(), n_calls)
    assert rel_err(gx.float(), emu.emu_minibatch_std_bwd(x, gxcat, n_calls)) < 1e-2


def test_r1_sqnorm_kernels():
    from textboxgan_b200 import kernels as Kn
    from textboxgan_b200.fused import R1SqNorm

    gen = torch.Generator().manual_seed(5)
    g = torch.randn(6, 3, 16, 64, generator=gen)
    out = Kn.r1_sqnorm(g.to(DEV))
    assert rel_err(out, (g.double() ** 2).sum(dim=(1, 2, 3))) < 1e-6
    go = torch.randn(6, generator=gen)
    assert rel_err(Kn.r1_sqnorm_bwd(g.to(DEV), go.to(DEV)), 2 * g.double() * go.double()[:, None, None, None]) < 1e-6
    ga = g.to(DEV).requires_grad_(True)
    (R1SqNorm.apply(ga) * go.to(DEV)).sum().backward()
    assert rel_err(ga.grad, 2 * g.double() * go.double()[:, None, None, None]) < 1e-6


@pytest.mark.parametrize("B,H,W,C,mcn,prev,nchw,masked", [
    (2, 2, 8, 192, 12, False, False, False), (3, 4, 16, 512, 8, True, False, False), (2, 16, 64, 256, 8, True, False, False),
    (2, 32, 128, 128, 8, True, True, True), (2, 64, 256, 128, 12, True, True, True), (1, 16, 64, 64, 16, True, True, False)])
def test_torgb_skip_kernel_vs_emulated_semantics(B, H, W, C, mcn, prev, nchw, masked):
    """tbg_torgb_skip_fwd (ToRGB to_rgb.py:28-33 + upsample_2d skip synthesis_block.py:152 + mask_text_box + NCHW) against
    the literal upfirdn_2d_ref / mask restatement, incl. the fractional char_width of BASELINE configs[2] (256 / 12); the
    autograd Function's gradients against autograd through the emulation."""
    from textboxgan_b200 import kernels as Kn
    from textboxgan_b200.fused import ToRGBSkip

    gen = torch.Generator().manual_seed(H + C)
    x = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    ws = torch.randn(B, C, 3, generator=gen) / math.sqrt(C)
    bias = torch.randn(3, generator=gen) * 0.1
    y_prev = torch.randn(B, H // 2, W // 2, 3, generator=gen) if prev else None
    words = None
    if masked:
        lens = torch.randint(1, mcn + 1, (B,), generator=gen)
        words = torch.where(torch.arange(mcn)[None] < lens[:, None], torch.randint(1, 70, (B, mcn), generator=gen),
                            torch.zeros(B, mcn, dtype=torch.long)).to(torch.int32)
    d = lambda t: t.to(DEV) if t is not None else None
    y = Kn.torgb_skip_fwd(d(x).bfloat16(), d(ws), d(bias), d(y_prev), d(words), nchw)
    want = emu.emu_torgb_skip_fwd(x, ws, bias, y_prev, words, nchw)
    assert y.shape == want.shape and rel_err(y, want) < 1e-5
    if masked:
        from oracle import train_step as OT
        from fractions import Fraction

        cw = Fraction(W, mcn)
        assert torch.equal(y.cpu(), OT.mask_text_box(y.cpu(), words, int(cw) if cw.denominator == 1 else cw))
    # gradients through the Function
    xa, wa, ba = d(x).bfloat16().requires_grad_(True), d(ws).requires_grad_(True), d(bias).requires_grad_(True)
    pa = d(y_prev).requires_grad_(True) if prev else None
    r = torch.randn(want.shape, generator=gen)
    (ToRGBSkip.apply(xa, wa, ba, pa, d(words), nchw) * d(r)).sum().backward()
    xe, we, be = x.double().requires_grad_(True), ws.double().requires_grad_(True), bias.double().requires_grad_(True)
    pe = y_prev.double().requires_grad_(True) if prev else None
    ye = torch.einsum("bhwc,bcj->bhwj", xe, we) + be
    if prev:
        import numpy as np
        t = np.array([1.0, 3.0, 3.0, 1.0])
        k = np.outer(t, t)
        ye = ye + OS.upfirdn_2d_ref(pe, k / k.sum() * 4.0, 2, 2, 1, 1, 2, 1, 2, 1)
    if masked:
        keep = (words.long()[:, (torch.arange(W) * mcn) // W] != 0).double()
        ye = ye * keep[:, None, :, None]
    if nchw:
        ye = ye.permute(0, 3, 1, 2)
    (ye * r.double()).sum().backward()
    assert rel_err(xa.grad.float(), xe.grad) < 1e-2 and rel_err(wa.grad, we.grad) < 1e-4 and rel_err(ba.grad, be.grad) < 1e-4
    if prev:
        assert rel_err(pa.grad, pe.grad) < 1e-5


@pytest.mark.parametrize("cfgk", [dict(upx=2, upy=2, padx0=2, padx1=1, pady0=2, pady1=1),
                                  dict(downx=2, downy=2, padx0=1, padx1=1, pady0=1, pady1=1),
                                  dict(upx=3, upy=2, downx=2, downy=3, padx0=4, padx1=-1, pady0=-2, pady1=5),
                                  dict(padx0=3, padx1=3, pady0=3, pady1=3)])
@pytest.mark.parametrize("shape,ksz", [((2, 33, 70, 3), (4, 4)), ((5, 8, 9, 1), (3, 5)), ((1, 16, 16, 19), (6, 6))])
def test_upfirdn2d_tiled_kernel_matches_reference_semantics(cfgk, shape, ksz):
    """tbg_upfirdn2d (the reference's UpFirDn2D op contract) against the literal upfirdn_2d_ref (upfirdn_2d_v2.py:249-305):
    tiles that overhang the output, minor sizes that are not multiples of the staged chunk, non-symmetric filters,
    negative pads, up and down at once."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen)
    k = torch.randn(*ksz, generator=gen)
    full = dict(upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0, pady1=0)
    full.update(cfgk)
    ref = OS.upfirdn_2d_ref(x.double(), k.double().numpy(), full["upx"], full["upy"], full["downx"], full["downy"],
                            full["padx0"], full["padx1"], full["pady0"], full["pady1"])
    if ref.shape[1] < 1 or ref.shape[2] < 1:
        pytest.skip("empty output for this combination")
    y = Kn.upfirdn2d(x.to(DEV), k.to(DEV), **full)
    assert y.shape == ref.shape and rel_err(y, ref) < 1e-5


def test_bias_act_fwd_and_rowdot_kernels():
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(11)
    for B, H, W, C in ((3, 8, 16, 64), (2, 16, 64, 256), (5, 3, 7, 128)):
        t = _bf16_round(torch.randn(B, H, W, C, generator=gen))
        u = _bf16_round(torch.randn(B, H, W, C, generator=gen))
        nz = torch.randn(B, H, W, generator=gen)
        ns = torch.tensor([0.4])
        bias = torch.randn(C, generator=gen) * 0.3
        for act, gain in ((1, math.sqrt(2.0)), (0, 1.0), (2, 1.0)):
            got = Kn.bias_act_fwd(t.to(DEV).bfloat16(), noise=nz.to(DEV), noise_strength=ns.to(DEV), bias=bias.to(DEV), act=act, gain=gain)
            assert rel_err(got.float(), emu.emu_bias_act_fwd(t, noise=nz, noise_strength=ns, bias=bias, act=act, gain=gain).float()) < 1e-2
        got = Kn.bias_act_fwd(t.to(DEV).bfloat16(), bias=bias.to(DEV), act=1, gain=1.0)
        assert rel_err(got.float(), emu.emu_bias_act_fwd(t, bias=bias, act=1, gain=1.0).float()) < 1e-2
        assert rel_err(Kn.rowdot(t.to(DEV).bfloat16(), u.to(DEV).bfloat16()), emu.emu_rowdot(t, u)) < 1e-5


def test_second_order_primitives_match_autograd_of_plain_torch():
    """The closed primitive set of second_order.py (modulate, rowdot, bias_act, mask_mul, to_rgb and its adjoints): first
    and SECOND derivatives of a path-length-like scalar against fp64 autograd through the same composite written with
    plain torch ops: J = d(sum(img * n)) / ds with create_graph, then d(sum(J^2)) / d(x, s, ws, d)."""
    from textboxgan_b200 import second_order as SO

    gen = torch.Generator().manual_seed(5)
    B, H, W, C = 3, 8, 16, 64
    x0 = _bf16_round(torch.randn(B, H, W, C, generator=gen))
    s0 = torch.rand(B, C, generator=gen) + 0.5
    d0 = torch.rand(B, C, generator=gen) + 0.5
    nz = torch.randn(B, H, W, generator=gen)
    ns0 = torch.tensor(0.3)
    b0 = torch.randn(C, generator=gen) * 0.2
    ws0 = torch.randn(B, C, 3, generator=gen) / math.sqrt(C)
    n_img = torch.randn(B, H, W, 3, generator=gen)
    g2 = math.sqrt(2.0)

    def run(device, dtype, prim):
        x = x0.to(device, dtype).requires_grad_(True)
        s, d, ws = (t.to(device).double().requires_grad_(True) if not prim else t.to(device).requires_grad_(True)
                    for t in (s0, d0, ws0))
        nzd, nsd, bd, nd = nz.to(device), ns0.to(device), b0.to(device), n_img.to(device)
        if prim:
            t = SO.modulate(SO.modulate(x, s), d)
            y = SO.bias_act(t, nzd, nsd, bd, 1, g2)
            img = SO.to_rgb(y, ws)
        else:
            # same bf16 rounding points as the kernels (values rounded, gradients passed straight through), so that the
            # leaky-ReLU masks of near-zero pre-activations agree
            r = lambda v: v + (v.to(torch.bfloat16).double() - v).detach()
            t = r(r(x.double() * s[:, None, None, :]) * d[:, None, None, :])
            pre = t + nzd.double()[..., None] * nsd.double() + bd.double()
            y = r(torch.nn.functional.leaky_relu(pre, 0.2) * g2)
            img = torch.einsum("bhwc,bcj->bhwj", y, ws)
        (J,) = torch.autograd.grad((img * nd.to(img.dtype)).sum(), s, create_graph=True)
        pen = (J ** 2).sum()
        grads = torch.autograd.grad(pen, [x, s, ws, d])
        return J.detach().double().cpu(), [g.detach().double().cpu() for g in grads]

    J_ref, g_ref = run("cpu", torch.float64, False)
    J_got, g_got = run(DEV, torch.bfloat16, True)
    rl2 = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    print("second-order primitives: rel-L2 of J", rl2(J_got, J_ref), "of d(sum J^2)/d(x, s, ws, d)",
          [rl2(a, b) for a, b in zip(g_got, g_ref)])
    assert rl2(J_got, J_ref) < 2e-2                    # gradients travel as bf16 tensors between the kernels
    for name, a, b in zip(("x", "s", "ws", "d"), g_got, g_ref):
        assert rl2(a, b) < 5e-2, name


def test_batch_resize_normalize_kernel_vs_emulation_and_cv2():
    """tbg_batch_resize_normalize (the loader transform of training_data_loader.py:63-82 for a whole batch): against the
    numpy emulation (identical up to fused-multiply-add rounding at exact .5 ties: <= 1 LSB on < 0.1 % of pixels) and
    against cv2 itself (fixed-point bilinear: <= 1 LSB)."""
    import numpy as np
    from textboxgan_b200 import kernels as Kn

    cv2 = pytest.importorskip("cv2")
    H, W = 64, 256
    rng = np.random.RandomState(11)
    dst_w = [256, 128, 16, 96, 80, 256]
    shapes = [(37, 301), (128, 256), (9, 5), (64, 96), (200, 31), (64, 256)]
    imgs = [rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8) for h, w in shapes]
    sizes = [im.size for im in imgs]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)[:-1]]), dtype=torch.int64)
    packed = torch.from_numpy(np.concatenate([im.reshape(-1) for im in imgs]))
    sh = torch.tensor([s[0] for s in shapes], dtype=torch.int32)
    sw = torch.tensor([s[1] for s in shapes], dtype=torch.int32)
    dw = torch.tensor(dst_w, dtype=torch.int32)
    got = Kn.batch_resize_normalize(packed.to(DEV), offsets.to(DEV), sh.to(DEV), sw.to(DEV), dw.to(DEV), H, W).cpu()
    want = emu.emu_batch_resize_normalize(packed, offsets, sh, sw, dw, H, W)
    lsb = 1.0 / 127.5
    d = (got - want).abs()
    assert float(d.max()) <= lsb + 1e-6 and float((d > 1e-6).float().mean()) < 1e-3
    for b, (im, w_) in enumerate(zip(imgs, dst_w)):
        ref = cv2.resize(im, (w_, H)).astype(np.float32) / 127.5 - 1.0
        ref = torch.from_numpy(ref.transpose(2, 0, 1))
        assert float((got[b, :, :, :w_] - ref).abs().max()) <= lsb + 1e-6
        if w_ < W:
            assert float(got[b, :, :, w_:].abs().max()) == 0.0
    assert torch.equal(got[1], want[1]) and torch.equal(got[5], want[5])      # 2:1 area mean and identity are exact


@pytest.mark.parametrize("sy", [1, 2])
@pytest.mark.parametrize("B,H,W,C", [(3, 16, 64, 64), (2, 8, 34, 128), (1, 66, 70, 8), (2, 4, 4, 256)])
def test_fir4_down_and_adjoint_kernels(sy, B, H, W, C):
    """tbg_fir4_down / tbg_fir4_down_adjoint (skip branch of a residual block: upfirdn_2d_v2.py:106-113 with a 1x1 kernel):
    the decimated filter equals the full-resolution tbg_fir4 semantics sub-sampled, the adjoint is the exact transpose
    (+ add); one bf16 rounding of the result."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(H * 7 + W + sy)
    x = torch.randn(B, H, W, C, generator=gen).bfloat16()
    OH, OW = H // sy, W // 2
    got = Kn.fir4_down(x.to(DEV), (OH, OW), sy, (-1, -1), 1.0 / 64.0).cpu()
    want = emu.emu_fir4_down(x, (OH, OW), sy, (-1, -1), 1.0 / 64.0)
    full = emu.emu_fir4(x, (H, W), (-1, -1), 1.0 / 64.0)[:, ::sy, ::2][:, :OH, :OW]
    assert torch.equal(want, full)
    assert rel_err(got, want) < 4e-3 and float((got.float() - want.float()).abs().max()) <= 2.0 ** -7 * float(want.abs().max())
    g = torch.randn(B, OH, OW, C, generator=gen).bfloat16()
    add = torch.randn(B, H, W, C, generator=gen).bfloat16()
    for a in (None, add):
        gx = Kn.fir4_down_adjoint(g.to(DEV), (H, W), sy, (-1, -1), 1.0 / 64.0, add=a.to(DEV) if a is not None else None).cpu()
        ex = emu.emu_fir4_down_adjoint(g, (H, W), sy, (-1, -1), 1.0 / 64.0, add=a)
        assert rel_err(gx, ex) < 4e-3, (sy, a is not None, rel_err(gx, ex))
    # <fir_down(x), g> == <x, adjoint(g)> within bf16 rounding of the two results
    lhs = float((got.double() * g.double()).sum())
    rhs = float((x.double() * Kn.fir4_down_adjoint(g.to(DEV), (H, W), sy, (-1, -1), 1.0 / 64.0).cpu().double()).sum())
    assert abs(lhs - rhs) <= 2e-2 * (got.double() * g.double()).abs().sum() ** 0.5 + 1e-3 * abs(lhs)


@pytest.mark.parametrize("B,H,W,C,has_g,has_noise", [(4, 16, 64, 128, True, True), (3, 8, 32, 256, False, True),
                                                     (2, 64, 256, 128, True, True), (5, 4, 16, 512, True, False),
                                                     (2, 5, 7, 64, False, True)])
def test_bias_act_rgb_bwd_kernel_vs_emulated_semantics(B, H, W, C, has_g, has_noise):
    """tbg_bias_act_rgb_bwd: the activation backward of a modulated layer with the ToRGB input gradient formed in the
    kernel (g_out + g_rgb (x) ws), plus the ToRGB weight gradient — against the fp64 emulation (bf16 output, fp32 sums)."""
    from textboxgan_b200 import kernels as Kn

    gen = torch.Generator().manual_seed(B * 100 + C)
    out = torch.randn(B, H, W, C, generator=gen).bfloat16()
    g_out = torch.randn(B, H, W, C, generator=gen).bfloat16() if has_g else None
    g_rgb = torch.randn(B, H, W, 3, generator=gen)
    ws = torch.randn(B, C, 3, generator=gen) * 0.3
    noise = torch.randn(B, H, W, generator=gen) if has_noise else None
    d = torch.rand(B, C, generator=gen) + 0.5
    got = Kn.bias_act_rgb_bwd(g_out.to(DEV) if has_g else None, out.to(DEV), g_rgb.to(DEV), ws.to(DEV),
                              noise=noise.to(DEV) if has_noise else None, d=d.to(DEV), act=1, gain=math.sqrt(2.0))
    want = emu.emu_bias_act_rgb_bwd(g_out, out, g_rgb, ws, noise=noise, d=d, act=1, gain=math.sqrt(2.0))
    names = ("gy0", "S1", "Spre", "Snz", "gws")
    for n, a, b in zip(names, got, want):
        if n == "Snz" and not has_noise:
            continue
        tol = 4e-3 if n == "gy0" else 2e-4
        assert rel_err(a.float().cpu(), b.float()) < tol, (n, rel_err(a.float().cpu(), b.float()))


@pytest.mark.parametrize("kind,k,rh,H,W,I,O", [("plain", 3, True, 16, 32, 64, 128), ("plain", 1, True, 8, 16, 128, 64),
                                               ("upT", 3, True, 8, 16, 128, 64), ("upT", 3, True, 16, 16, 64, 64),
                                               ("downU", 3, True, 16, 32, 64, 128), ("downU", 3, False, 8, 32, 128, 128)])
def test_lin_conv_triple_first_and_second_derivatives_vs_reference_convolution(kind, k, rh, H, W, I, O):
    """second_order.lin_conv (F / A / G on the master weight, the regulariser passes' convolution): value, the
    path-length-like first derivative J = d<y, n>/dx taken with create_graph, and d(sum J^2 + <y, m>)/d(x, w), against fp64
    torch convolutions that restate the reference ops (plain SAME conv; upsample_conv_2d = transposed stride-2 conv + FIR,
    upfirdn_2d_v2.py:65-103; conv_downsample_2d = FIR + strided conv, :106-113)."""
    import torch.nn.functional as F_
    from textboxgan_b200 import second_order as SO

    gen = torch.Generator().manual_seed(H * 3 + W + I)
    B = 3
    x0 = _bf16_round(torch.randn(B, H, W, I, generator=gen))
    w0 = torch.randn(k, k, I, O, generator=gen)
    coef = 1.0 / math.sqrt(k * k * I)
    kf = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    k2 = torch.outer(kf, kf) / 64.0

    def ref_conv(x, w):                                   # x [B,H,W,I] fp64, w [k,k,I,O] fp64 -> NHWC
        xn = x.permute(0, 3, 1, 2)
        wk = (w * coef).permute(3, 2, 0, 1)               # OIHW
        if kind == "plain":
            y = F_.conv2d(xn, wk, padding=k // 2)
        elif kind == "upT":
            t = F_.conv_transpose2d(xn, (w * coef).flip(0, 1).permute(2, 3, 0, 1), stride=2)  # flipped kernel (:80)
            kk = (k2 * 4.0)[None, None].expand(O, 1, 4, 4)
            y = F_.conv2d(F_.pad(t, (1, 1, 1, 1)), kk, groups=O)                             # pad (1,1): 2H x 2W
        else:
            kk = k2[None, None].expand(I, 1, 4, 4)
            xb = F_.conv2d(F_.pad(xn, (2, 2, 2, 2 if rh else 3)), kk, groups=I)
            y = F_.conv2d(xb, wk, stride=(2 if rh else 1, 2))
        return y.permute(0, 2, 3, 1)

    def run(device, prim):
        x = x0.to(device, torch.bfloat16 if prim else torch.float64).requires_grad_(True)
        w = w0.to(device, torch.float32 if prim else torch.float64).requires_grad_(True)
        y = SO.lin_conv(x, w, kind, k, rh, "t") if prim else ref_conv(x, w)
        gn = torch.Generator().manual_seed(99)
        n = torch.randn(y.shape, generator=gn).to(device, y.dtype)
        m = torch.randn(y.shape, generator=gn).to(device, y.dtype)
        (J,) = torch.autograd.grad((y * n).sum(), x, create_graph=True)
        pen = (J.double() ** 2).sum() + (y.double() * m.double()).sum()
        gx, gw = torch.autograd.grad(pen, [x, w])
        return [t.detach().double().cpu() for t in (y, J, gx, gw)]

    ref = run("cpu", False)
    got = run(DEV, True)
    rl2 = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    errs = [rl2(a, b) for a, b in zip(got, ref)]
    print(kind, "rel-L2 of (y, J, d/dx, d/dw):", errs)
    assert ref[0].shape == got[0].shape
    assert errs[0] < 1e-2 and errs[1] < 1e-2 and errs[2] < 3e-2 and errs[3] < 3e-2, errs


def test_training_data_loader_device_transform_on_gpu(tmp_path):
    """Row f2 on the device: TrainingDataLoader(device_transform=True) — host PNG decode, one packed host-to-device copy, one
    tbg_batch_resize_normalize launch per batch — against the host cv2 transform of the same stream (same seed, same
    order): labels identical, images within one uint8 level (cv2's fixed-point bilinear), zero pad exact."""
    import os

    import numpy as np
    from common import small_cfg
    from textboxgan_b200.data_loader import TrainingDataLoader

    cv2 = pytest.importorskip("cv2")
    cfg = small_cfg(2)
    boxes = tmp_path / "text_boxes"
    boxes.mkdir()
    rng = np.random.RandomState(3)
    words = ["Hi", "a,b", "World!", "x", "Text", "Boxes"]
    lines = []
    for i, w in enumerate(words):
        cv2.imwrite(str(boxes / f"{i}.png"), rng.randint(0, 256, size=(17 + 3 * i, 25 + 9 * len(w), 3), dtype=np.uint8))
        lines.append(f"{i}.png,{w}\n")
    (boxes / "annotations_filtered.txt").write_text("".join(lines))
    host = TrainingDataLoader(cfg, str(boxes), None, device=DEV, seed=5)
    devl = TrainingDataLoader(cfg, str(boxes), None, device=DEV, seed=5, device_transform=True)
    a = list(host.load_dataset(batch_size=3, repeat=False))
    b = list(devl.load_dataset(batch_size=3, repeat=False))
    assert len(a) == len(b) == 2
    for (ra, _, ia, la), (rb, _, ib, lb) in zip(a, b):
        assert torch.equal(ia, ib) and torch.equal(la, lb)
        assert rb.is_cuda and rb.shape == ra.shape and rb.dtype == torch.float32
        d = (ra - rb).abs()
        assert float(d.max()) <= 1.0 / 127.5 + 1e-6
        lens = (ia > 0).sum(1)
        for i in range(ra.shape[0]):
            wpx = int(cfg.char_width * int(lens[i]))
            assert float(rb[i, :, :, wpx:].abs().max()) == 0.0
