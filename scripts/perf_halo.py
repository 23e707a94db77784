"""Round-2: correctness + speed of the experimental halo-reuse 3x3 convolution (csrc/conv_halo.cu) against
tbg_conv2d_igemm on the same tensors.  Usage (GPU box): python scripts/perf_halo.py [batch]"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from textboxgan_b200 import kernels as K, conv as CV, lib

h = lib.load()
fn = h.tbg_conv3x3_halo
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_int, C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"

def halo(x, w, out, use_bo, **epi):
    Bq, H, W, Cin = x.shape
    st = fn(x.data_ptr(), w.data_ptr(), out.data_ptr(), Bq, H, W, Cin, w.shape[0],
            epi["col_scale"].data_ptr() if epi.get("col_scale") is not None else None,
            epi["bias"].data_ptr() if epi.get("bias") is not None else None,
            epi["noise"].data_ptr() if epi.get("noise") is not None else None,
            epi["noise_strength"].data_ptr() if epi.get("noise") is not None else None,
            epi.get("act", 0), epi.get("act_gain", 1.0), use_bo, torch.cuda.current_stream().cuda_stream)
    lib.check(st, "tbg_conv3x3_halo")
    return out

def timeit(f, iters=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for (H, W, Cin, Cout) in [(16, 64, 64, 64), (32, 128, 128, 128), (16, 64, 256, 128), (64, 256, 128, 128)]:
    g = CV.plain_geom(H, W, Cin, Cout, 3)
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
    epi = dict(col_scale=torch.rand(B, Cout, device=dev) + 0.5, bias=torch.randn(Cout, device=dev),
               noise=torch.randn(B, H, W, device=dev), noise_strength=torch.ones(1, device=dev), act=1, act_gain=1.41421356)
    ref = K.conv2d_igemm(x, w, **g.kernel_kwargs(), **epi)
    fl = g.flops(B)
    t_ref = timeit(lambda: K.conv2d_igemm(x, w, out=ref, **g.kernel_kwargs(), **epi))
    line = f"{H}x{W} {Cin}->{Cout} B={B}: igemm {t_ref:7.1f} us {fl / t_ref / 1e6:7.1f} TF/s"
    for bo in (0, 1):
        out = torch.zeros_like(ref)
        try:
            halo(x, w, out, bo, **epi)
            torch.cuda.synchronize()
            err = ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
            t = timeit(lambda: halo(x, w, out, bo, **epi)) if err < 2e-2 else float("nan")
            line += f" | halo(base_offset={bo}) rel_err {err:.2e} {t:7.1f} us {fl / t / 1e6 if t == t else 0:7.1f} TF/s"
        except Exception as ex:
            line += f" | halo(base_offset={bo}) FAILED: {str(ex)[:80]}"
    print(line, flush=True)
