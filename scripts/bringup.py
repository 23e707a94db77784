"""GPU bring-up of the tcgen05 conv kernels against torch fp32 convs on bf16-rounded inputs.

Scratch tooling (not a test): each group runs in its own subprocess under a timeout so that a
trapped kernel cannot take the other groups down.  Usage on the GPU box:
    python scripts/bringup.py            # all groups, log to gpurun_out/bringup.log
    python scripts/bringup.py --group fwd_basic
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["fwd_basic", "fwd_epi", "fwd_up", "fwd_stride", "wgrad_basic", "wgrad_up", "wgrad_stride", "perf"]


def rel_err(a, b):
    import torch

    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def ref_conv(x_nhwc, w_nk, taps, pad, stride, Ho, Wo):
    """fp32 reference: x [B,H,W,C] bf16, w [N, taps*C] bf16 -> [B,Ho,Wo,N] fp32."""
    import torch
    import torch.nn.functional as F

    B, H, W, C = x_nhwc.shape
    N = w_nk.shape[0]
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_nk.float().view(N, taps[0], taps[1], C).permute(0, 3, 1, 2)
    # explicit (possibly asymmetric) padding so that out = Ho x Wo
    need_h = (Ho - 1) * stride[0] + taps[0]
    need_w = (Wo - 1) * stride[1] + taps[1]
    pad_b = max(need_h - H - pad[0], 0)
    pad_r = max(need_w - W - pad[1], 0)
    xp = F.pad(x, (pad[1], pad_r, pad[0], pad_b))
    y = F.conv2d(xp, w, stride=stride)
    y = y[:, :, :Ho, :Wo]
    return y.permute(0, 2, 3, 1).contiguous()


def run_fwd_case(name, B, H, W, Cin, Cout, taps=(3, 3), pad=(1, 1), stride=(1, 1), Ho=None, Wo=None, epi=False,
                 up=False, out_fp32=False, seed=0):
    import torch

    from textboxgan_b200 import kernels as K

    torch.manual_seed(seed)
    dev = "cuda"
    Ho = Ho or H
    Wo = Wo or W
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    n_total = 4 * Cout if up else Cout
    w = (torch.randn(n_total, taps[0] * taps[1] * Cin, device=dev) / (taps[0] * taps[1] * Cin) ** 0.5).to(torch.bfloat16)
    kw = {}
    oH, oW = (2 * Ho, 2 * Wo) if up else (Ho, Wo)
    if epi:
        kw["col_scale"] = torch.rand(B, Cout, device=dev) + 0.5
        kw["bias"] = torch.randn(Cout, device=dev) * 0.1
        kw["noise"] = torch.randn(B, oH, oW, device=dev)
        kw["noise_strength"] = torch.tensor(0.3, device=dev)
        kw["residual"] = torch.randn(B, oH, oW, Cout, device=dev).to(torch.bfloat16)
        kw["res_scale"] = 0.70710678
        kw["act"] = 1
        kw["act_gain"] = 1.41421356
    y = K.conv2d_igemm(x, w, Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=up, out_fp32=out_fp32, **kw)
    torch.cuda.synchronize()
    ref = ref_conv(x, w, taps, pad, stride, Ho, Wo)  # [B,Ho,Wo,n_total]
    if up:
        ref = ref.view(B, Ho, Wo, 2, 2, Cout).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * Ho, 2 * Wo, Cout)
    if epi:
        ref = ref * kw["col_scale"][:, None, None, :]
        ref = ref + kw["noise"][..., None] * kw["noise_strength"]
        ref = ref + kw["bias"]
        ref = torch.where(ref > 0, ref, 0.2 * ref) * kw["act_gain"]
        ref = (ref + kw["residual"].float()) * kw["res_scale"]
    err = rel_err(y, ref)
    tol = 2e-5 if out_fp32 else 1e-2
    status = "OK " if err < tol else "BAD"
    print(f"[{status}] fwd {name}: B={B} H={H} W={W} Cin={Cin} Cout={Cout} taps={taps} pad={pad} stride={stride} "
          f"Ho={Ho} Wo={Wo} up={up} epi={epi} fp32={out_fp32} rel_err={err:.3e}", flush=True)
    if err >= tol:
        d = (y.float() - ref).abs()
        idx = d.flatten().argmax().item()
        print("     worst idx", idx, "got", y.flatten()[idx].item(), "ref", ref.flatten()[idx].item(),
              " nonfinite:", (~torch.isfinite(y.float())).sum().item(), flush=True)
        # per-row error pattern hint
        bad = (d > tol * ref.abs().max()).float()
        print("     bad frac", bad.mean().item(), "by batch", bad.mean(dim=(1, 2, 3)).tolist()[:8], flush=True)
        print("     bad by channel(first 16)", bad.mean(dim=(0, 1, 2))[:16].tolist(), flush=True)
        print("     bad by row h", bad.mean(dim=(0, 2, 3))[:16].tolist(), flush=True)
        print("     bad by col w(first 16)", bad.mean(dim=(0, 1, 3))[:16].tolist(), flush=True)
    return err < tol


def run_wgrad_case(name, B, H, W, Cin, Cout, taps=(3, 3), pad=(1, 1), stride=(1, 1), Ho=None, Wo=None, up=False, seed=0):
    import torch

    from textboxgan_b200 import kernels as K

    torch.manual_seed(seed)
    dev = "cuda"
    Ho = Ho or H
    Wo = Wo or W
    oH, oW = (2 * Ho, 2 * Wo) if up else (Ho, Wo)
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    gy = torch.randn(B, oH, oW, Cout, device=dev).to(torch.bfloat16)
    gw = K.conv2d_wgrad(x, gy, Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=up)
    torch.cuda.synchronize()
    # reference through autograd of the fp32 conv
    n_total = 4 * Cout if up else Cout
    w = torch.zeros(n_total, taps[0] * taps[1] * Cin, device=dev, requires_grad=True)
    ref = ref_conv(x, w, taps, pad, stride, Ho, Wo)
    if up:
        ref = ref.view(B, Ho, Wo, 2, 2, Cout).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * Ho, 2 * Wo, Cout)
    (gref,) = torch.autograd.grad(ref, w, gy.float())
    err = rel_err(gw, gref)
    tol = 2e-3
    status = "OK " if err < tol else "BAD"
    print(f"[{status}] wgrad {name}: B={B} H={H} W={W} Cin={Cin} Cout={Cout} taps={taps} pad={pad} stride={stride} "
          f"Ho={Ho} Wo={Wo} up={up} rel_err={err:.3e}", flush=True)
    if err >= tol:
        d = (gw - gref).abs()
        bad = (d > tol * gref.abs().max()).float().view(n_total, taps[0] * taps[1], Cin)
        print("     bad frac", bad.mean().item(), " nonfinite:", (~torch.isfinite(gw)).sum().item(), flush=True)
        print("     bad by tap", bad.mean(dim=(0, 2)).tolist(), flush=True)
        print("     bad by n (first 16)", bad.mean(dim=(1, 2))[:16].tolist(), flush=True)
        print("     bad by c (first 16)", bad.mean(dim=(0, 1))[:16].tolist(), flush=True)
        print("     ratio sample", (gw.flatten()[:8] / gref.flatten()[:8]).tolist(), flush=True)
    return err < tol


def perf_case(name, B, H, W, Cin, Cout, up=False, iters=20):
    import torch

    from textboxgan_b200 import kernels as K

    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    n_total = 4 * Cout if up else Cout
    w = (torch.randn(n_total, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    oH, oW = (2 * H, 2 * W) if up else (H, W)
    out = torch.empty(B, oH, oW, Cout, device=dev, dtype=torch.bfloat16)
    gy = torch.randn(B, oH, oW, Cout, device=dev).to(torch.bfloat16)
    gw = torch.zeros(n_total, 9 * Cin, device=dev)
    for fn_name in ("fwd", "wgrad"):
        def call():
            if fn_name == "fwd":
                K.conv2d_igemm(x, w, Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), up=up, out=out)
            else:
                K.conv2d_wgrad(x, gy, Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), up=up, gw=gw)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * B * H * W * n_total * 9 * Cin
        print(f"[perf] {fn_name} {name}: B={B} H={H} W={W} Cin={Cin} Cout={Cout} up={up}: {ms * 1e3:.1f} us, "
              f"{flops / ms / 1e9:.1f} TFLOP/s", flush=True)


def run_group(group: str) -> int:
    import torch

    assert torch.cuda.is_available()
    print(f"== group {group} on {torch.cuda.get_device_name(0)}", flush=True)
    ok = True
    if group == "fwd_basic":
        ok &= run_fwd_case("1x1-min", 1, 2, 64, 64, 32, taps=(1, 1), pad=(0, 0), out_fp32=True)
        ok &= run_fwd_case("1x1-K128", 2, 8, 32, 128, 64, taps=(1, 1), pad=(0, 0), out_fp32=True)
        ok &= run_fwd_case("3x3-a", 4, 16, 64, 64, 64, out_fp32=True)
        ok &= run_fwd_case("3x3-b", 2, 32, 128, 128, 128)
        ok &= run_fwd_case("3x3-lowres", 4, 4, 16, 512, 512)
        ok &= run_fwd_case("3x3-oob-batch", 4, 2, 8, 128, 512)
        ok &= run_fwd_case("3x3-wide", 2, 8, 256, 64, 128)
        ok &= run_fwd_case("3x3-N256", 3, 8, 32, 256, 256)
        ok &= run_fwd_case("3x3-4x4", 8, 4, 4, 512, 512)
        ok &= run_fwd_case("3x3-many-tiles", 16, 64, 256, 128, 128)
    elif group == "fwd_epi":
        ok &= run_fwd_case("epi-fp32", 4, 16, 64, 64, 64, epi=True, out_fp32=True)
        ok &= run_fwd_case("epi-bf16", 4, 16, 64, 128, 128, epi=True)
        ok &= run_fwd_case("epi-lowres", 6, 4, 16, 256, 512, epi=True)
    elif group == "fwd_up":
        ok &= run_fwd_case("up-a", 2, 8, 32, 64, 64, up=True, out_fp32=True)
        ok &= run_fwd_case("up-b", 4, 4, 16, 512, 256, up=True)
        ok &= run_fwd_case("up-epi", 4, 16, 64, 128, 128, up=True, epi=True)
        ok &= run_fwd_case("up-lowres", 4, 2, 8, 128, 512, up=True, epi=True)
    elif group == "fwd_stride":
        ok &= run_fwd_case("s2-valid", 2, 34, 130, 64, 128, pad=(0, 0), stride=(2, 2), Ho=16, Wo=64, out_fp32=True)
        ok &= run_fwd_case("s12-valid", 2, 10, 34, 128, 256, pad=(0, 0), stride=(1, 2), Ho=8, Wo=16)
        ok &= run_fwd_case("s2-1x1", 2, 32, 128, 64, 128, taps=(1, 1), pad=(0, 0), stride=(2, 2), Ho=16, Wo=64)
        ok &= run_fwd_case("s2-6x6", 2, 16, 64, 64, 64, taps=(6, 6), pad=(2, 2), stride=(2, 2), Ho=8, Wo=32)
    elif group == "wgrad_basic":
        ok &= run_wgrad_case("1x1", 2, 8, 32, 64, 128, taps=(1, 1), pad=(0, 0))
        ok &= run_wgrad_case("3x3-a", 4, 16, 64, 64, 64)
        ok &= run_wgrad_case("3x3-b", 2, 32, 128, 128, 128)
        ok &= run_wgrad_case("3x3-256", 3, 8, 32, 256, 256)
        ok &= run_wgrad_case("3x3-512", 4, 4, 16, 512, 512)
        ok &= run_wgrad_case("3x3-oob", 4, 2, 8, 128, 512)
        ok &= run_wgrad_case("3x3-4x4", 8, 4, 4, 512, 512)
    elif group == "wgrad_up":
        ok &= run_wgrad_case("up-a", 2, 8, 32, 64, 64, up=True)
        ok &= run_wgrad_case("up-b", 4, 4, 16, 512, 256, up=True)
    elif group == "wgrad_stride":
        ok &= run_wgrad_case("s2", 2, 34, 130, 64, 128, pad=(0, 0), stride=(2, 2), Ho=16, Wo=64)
        ok &= run_wgrad_case("s12", 2, 10, 34, 128, 256, pad=(0, 0), stride=(1, 2), Ho=8, Wo=16)
    elif group == "perf":
        perf_case("top-conv1", 32, 64, 256, 128, 128)
        perf_case("mid", 32, 16, 64, 256, 256)
        perf_case("low", 32, 4, 16, 512, 512)
        perf_case("up-top", 32, 32, 128, 128, 128, up=True)
    print(f"== group {group} {'PASS' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--groups", default=",".join(GROUPS))
    ap.add_argument("--timeout", type=int, default=180)
    args = ap.parse_args()
    if args.group:
        sys.exit(run_group(args.group))
    os.makedirs("gpurun_out", exist_ok=True)
    log = open("gpurun_out/bringup.log", "a")
    rc_all = 0
    for g in args.groups.split(","):
        t0 = time.time()
        try:
            proc = subprocess.run([sys.executable, __file__, "--group", g], capture_output=True, text=True,
                                  timeout=args.timeout)
            out = proc.stdout + proc.stderr[-3000:]
            rc = proc.returncode
        except subprocess.TimeoutExpired as ex:
            out = (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
            out += f"\n== group {g} TIMEOUT after {args.timeout}s\n"
            rc = 124
        msg = f"{out}\n-- group {g} rc={rc} ({time.time() - t0:.1f}s)\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()
        rc_all |= rc
    sys.exit(1 if rc_all else 0)


if __name__ == "__main__":
    main()
