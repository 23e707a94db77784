"""Isolated timings of the bandwidth-bound kernels at BASELINE configs[2] shapes: achieved GB/s (algorithmic bytes = one
read of each input + one write of each output) against the measured HBM copy bandwidth.
   python scripts/perf_pointwise.py [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from textboxgan_b200 import kernels as K

peak = 6552.3
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def bench(fn, n_rot, iters=20):
    for i in range(3): fn(i % n_rot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:58s} {us:8.1f} us  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of measured HBM copy", flush=True)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
for (H, W, C) in [(64, 256, 128), (32, 128, 128), (16, 64, 256), (64, 256, 64)]:
    n = B * H * W * C
    n_rot = max(2, min(8, int(600e6 // (n * 2)) + 1))
    g = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(n_rot)]
    o = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(n_rot)]
    nz = torch.randn(B, H, W, device=dev)
    d = torch.rand(B, C, device=dev) + 0.5
    s = torch.rand(B, C, device=dev) + 0.5
    t = bench(lambda i: K.bias_act_bwd(g[i], o[i], noise=nz, d=d, act=True, gain=1.4), n_rot)
    report(f"bias_act_bwd {B}x{H}x{W}x{C} (modconv: sums)", t, 3 * n * 2)
    t = bench(lambda i: K.bias_act_bwd(g[i], o[i], act=True, gain=1.4, want_sums=False, bias_grad_only=True), n_rot)
    report(f"bias_act_bwd {B}x{H}x{W}x{C} (dconv: bias grad)", t, 3 * n * 2)
    grgb = torch.randn(B, H, W, 3, device=dev)
    wsr = torch.randn(B, C, 3, device=dev)
    t = bench(lambda i: K.bias_act_rgb_bwd(g[i], o[i], grgb, wsr, noise=nz, d=d, act=1, gain=1.4), n_rot)
    report(f"bias_act_rgb_bwd {B}x{H}x{W}x{C} (+ ToRGB gradient)", t, 3 * n * 2)
    t = bench(lambda i: K.modulate_bwd(g[i], o[i], s), n_rot)
    report(f"modulate_bwd {B}x{H}x{W}x{C}", t, 3 * n * 2)
    t = bench(lambda i: K.modulate(g[i], s), n_rot)
    report(f"modulate {B}x{H}x{W}x{C}", t, 2 * n * 2)
    ws = torch.randn(B, C, 3, device=dev)
    gy = torch.randn(B, H, W, 3, device=dev)
    t = bench(lambda i: K.torgb_bwd(g[i], ws, gy), n_rot)
    report(f"torgb_bwd {B}x{H}x{W}x{C}", t, 2 * n * 2)
    t = bench(lambda i: K.torgb_skip_fwd(g[i], ws, None, None, None, False), n_rot)
    report(f"torgb_skip_fwd {B}x{H}x{W}x{C}", t, n * 2)
    yprev = torch.randn(B, H // 2, W // 2, 3, device=dev)
    wordsi = torch.randint(0, 2, (B, 12), device=dev, dtype=torch.int32)
    t = bench(lambda i: K.torgb_skip_fwd(g[i], ws, None, yprev, wordsi, True), n_rot)
    report(f"torgb_skip_fwd {B}x{H}x{W}x{C} + skip + mask, NCHW", t, n * 2)
    # FIR adjoint of an up layer's backward: [B,H,W,C] -> [B,H+2,W+2,C]
    t = bench(lambda i: K.fir4(g[i], (H + 2, W + 2), (-2, -2), 1.0 / 16.0), n_rot)
    report(f"fir4 adjoint {B}x{H}x{W}x{C} -> +2", t, 2 * n * 2)
    t = bench(lambda i: K.fir4(g[i], (H + 2, W + 2), (-2, -2), 1.0 / 64.0), n_rot)
    report(f"fir4 down pre-pass {B}x{H}x{W}x{C}", t, 2 * n * 2)

# FromRGB backward of the discriminator (3 -> 64 channels at the full resolution; fake + real = 2 x batch)
for Bf in (2 * B, B):
    H, W, C = 64, 256, 64
    n = Bf * H * W * C
    n_rot = max(2, min(8, int(600e6 // (n * 2)) + 1))
    g = [torch.randn(Bf, H, W, C, device=dev).bfloat16() for _ in range(n_rot)]
    o = [torch.randn(Bf, H, W, C, device=dev).bfloat16() for _ in range(n_rot)]
    img = torch.randn(Bf, 3, H, W, device=dev)
    wf = torch.randn(3, C, device=dev)
    t = bench(lambda i: K.fromrgb_bwd(img, wf, g[i], o[i], 0.5, 1.4), n_rot)
    report(f"fromrgb_bwd {Bf}x{H}x{W}x{C} (image + weight gradients)", t, 2 * n * 2 + 2 * Bf * 3 * H * W * 4)
    bf = torch.randn(C, device=dev)
    t = bench(lambda i: K.fromrgb_fwd(img, wf, bf, 0.5, 1.4), n_rot)
    report(f"fromrgb_fwd {Bf}x{H}x{W}x{C}", t, n * 2 + Bf * 3 * H * W * 4)
