"""Scratch: isolated per-layer kernel timings (back-to-back launches over rotating, > L2 buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C

def bench(fn, n_rot, iters=30):
    for i in range(3): fn(i % n_rot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def run(name, B, g):
    dev = "cuda"
    oh, ow = g.out_hw
    xbytes = B * g.H * g.W * g.cin * 2
    n_rot = max(2, int(300e6 // max(xbytes, 1)) + 1); n_rot = min(n_rot, 64)
    xs = [torch.randn(B, g.H, g.W, g.cin, device=dev).bfloat16() for _ in range(n_rot)]
    gys = [torch.randn(B, oh, ow, g.cout, device=dev).bfloat16() for _ in range(n_rot)]
    w = (torch.randn(g.n_total, g.k_total, device=dev) / g.k_total ** 0.5).bfloat16()
    out = torch.empty(B, oh, ow, g.cout, device=dev, dtype=torch.bfloat16)
    gw = torch.zeros(g.n_total, g.k_total, device=dev)
    kw = g.kernel_kwargs()
    t_f = bench(lambda i: K.conv2d_igemm(xs[i], w, out=out, **kw), n_rot)
    t_w = bench(lambda i: K.conv2d_wgrad(xs[i], gys[i], gw=gw, **kw), n_rot)
    fl = g.flops(B)
    print(f"{name:28s} B={B} fwd {t_f:7.1f} us {fl/t_f/1e6:7.1f} TF/s (algo {fl*g.algo_frac/t_f/1e6:6.1f}) | wgrad {t_w:7.1f} us {fl/t_w/1e6:7.1f} TF/s", flush=True)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
print("TBG_WGRAD_ITEMS_PER_SM", os.environ.get("TBG_WGRAD_ITEMS_PER_SM"))
run("mod 4x16 512->512", B, C.plain_geom(4, 16, 512, 512, 3))
run("mod up 2x8 128->512", B, C.up_geom(2, 8, 128, 512))
run("mod 8x32 256->256", B, C.plain_geom(8, 32, 256, 256, 3))
run("mod up 4x16 512->256", B, C.up_geom(4, 16, 512, 256))
run("mod 16x64 256->256", B, C.plain_geom(16, 64, 256, 256, 3))
run("mod up 8x32 256->256", B, C.up_geom(8, 32, 256, 256))
run("mod 32x128 128->128", B, C.plain_geom(32, 128, 128, 128, 3))
run("mod up 16x64 256->128", B, C.up_geom(16, 64, 256, 128))
run("d 32x128 128->128", B, C.plain_geom(32, 128, 128, 128, 3))
run("d down 32x128 128->128", B, C.down_geom(32, 128, 128, 128, 3, True))
run("d skip 32x128 128->128", B, C.down_geom(32, 128, 128, 128, 1, True))
run("d 16x64 128->128", B, C.plain_geom(16, 64, 128, 128, 3))
run("d down 16x64 128->256", B, C.down_geom(16, 64, 128, 256, 3, True))
run("d 8x32 256->256", B, C.plain_geom(8, 32, 256, 256, 3))
run("d 4x8 512->512", B, C.plain_geom(4, 8, 512, 512, 3))
run("big 64x256 128->128", B, C.plain_geom(64, 256, 128, 128, 3))


def run_upT(name, B, h, w, I, O):
    """Unfolded upsample_conv_2d: transposed conv (tap masks) + FIR pass; dgrad = stride-2 conv; wgrad role-swapped."""
    dev = "cuda"
    spec = C.weight_spec("upT", h, w, I, O, 3, True, "modconv")
    n_rot = min(64, max(2, int(300e6 // (B * h * w * I * 2)) + 1))
    xs = [torch.randn(B, h, w, I, device=dev).bfloat16() for _ in range(n_rot)]
    th, tw = spec.t_hw
    gTs = [torch.randn(B, th, tw, O, device=dev).bfloat16() for _ in range(min(n_rot, 8))]
    wf = (torch.randn(spec.fwd_rows, spec.fwd_cols, device=dev) / (9 * I) ** 0.5).bfloat16()
    wa = (torch.randn(spec.adj_rows, spec.adj_cols, device=dev) / (9 * I) ** 0.5).bfloat16()
    T = torch.empty(B, th, tw, O, device=dev, dtype=torch.bfloat16)
    d = torch.rand(B, O, device=dev) + 0.5
    nz = torch.randn(B, 2 * h, 2 * w, device=dev)
    ns = torch.ones(1, device=dev)
    bias = torch.randn(O, device=dev)
    gx = torch.empty(B, h, w, I, device=dev, dtype=torch.bfloat16)
    gw = torch.zeros(spec.adj_rows, spec.adj_cols, device=dev)
    t_c = bench(lambda i: K.conv2d_igemm(xs[i], wf, out=T, **spec.fwd_kwargs), n_rot)
    t_f = bench(lambda i: K.fir4(gTs[i % len(gTs)], spec.out_hw, (-1, -1), 1 / 16, d=d, noise=nz, noise_strength=ns, bias=bias, act=1, gain=1.4), n_rot)
    t_d = bench(lambda i: K.conv2d_igemm(gTs[i % len(gTs)], wa, out=gx, **spec.s2_kwargs), n_rot)
    t_w = bench(lambda i: K.conv2d_wgrad(gTs[i % len(gTs)], xs[i], gw=gw, **spec.s2_kwargs), n_rot)
    algo = 2.0 * B * h * w * 9 * I * O
    print(f"{name:28s} B={B} convT {t_c:6.1f} us (algo {algo/t_c/1e6:6.1f} TF/s) fir {t_f:6.1f} us | dgrad {t_d:6.1f} us "
          f"(algo {algo/t_d/1e6:6.1f}) | wgrad {t_w:6.1f} us (algo {algo/t_w/1e6:6.1f})", flush=True)


run_upT("modT up 2x8 128->512", B, 2, 8, 128, 512)
run_upT("modT up 4x16 512->256", B, 4, 16, 512, 256)
run_upT("modT up 8x32 256->256", B, 8, 32, 256, 256)
run_upT("modT up 16x64 256->128", B, 16, 64, 256, 128)
run_upT("modT up 32x128 128->128", B, 32, 128, 128, 128)
