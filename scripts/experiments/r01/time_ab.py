import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
import torch
from textboxgan_b200 import lib, conv as Cv

old = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libr01.so"))
old.tbg_conv2d_igemm_r01.restype = C.c_int
old.tbg_conv2d_igemm_r01.argtypes = [C.POINTER(lib.ConvArgs), C.c_void_p]
new = lib.load()
lib.set_tuning("conv_halo", 0)
epi4 = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libepi4.so"))
epi4.tbg_conv2d_igemm.restype = C.c_int
epi4.tbg_conv2d_igemm.argtypes = [C.POINTER(lib.ConvArgs), C.c_void_p]
epi4.tbg_set_tuning.argtypes = [C.c_char_p, C.c_int]
epi4.tbg_set_tuning(b"conv_halo", 0)


def bench(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def run(name, B, g):
    kw = g.kernel_kwargs()
    oh, ow = g.out_hw
    n_rot = min(32, max(2, int(300e6 // (B * g.H * g.W * g.cin * 2)) + 1))
    xs = [torch.randn(B, g.H, g.W, g.cin, device="cuda").bfloat16() for _ in range(n_rot)]
    w = (torch.randn(g.n_total, g.k_total, device="cuda") / g.k_total ** 0.5).bfloat16()
    out = torch.empty(B, oh, ow, g.cout, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty_like(out)
    st = torch.cuda.current_stream().cuda_stream
    i = [0]

    def args(o):
        x = xs[i[0] % n_rot]; i[0] += 1
        return lib.ConvArgs(x=x.data_ptr(), w=w.data_ptr(), out=o.data_ptr(), B=B, H=g.H, W=g.W, Cin=g.cin, Ho=kw["Ho"], Wo=kw["Wo"],
                            n_total=g.n_total, cout=g.cout, taps_h=kw["taps"][0], taps_w=kw["taps"][1], pad_h=kw["pad"][0],
                            pad_w=kw["pad"][1], stride_h=kw["stride"][0], stride_w=kw["stride"][1], up_h=kw["up"][0],
                            up_w=kw["up"][1], act_gain=1.0, res_scale=1.0)

    def f_old():
        a = args(out); assert old.tbg_conv2d_igemm_r01(C.byref(a), st) == 0
    def f_new():
        a = args(out2); assert new.tbg_conv2d_igemm(C.byref(a), st) == 0
    t_old = bench(f_old)
    res = [f"r01 {t_old:6.1f} us"]
    for staged in (1, 0):
        lib.set_tuning("igemm_staged", staged)
        res.append(f"now(staged={staged}) {bench(f_new):6.1f} us")
    def f_epi4():
        a = args(out2); assert epi4.tbg_conv2d_igemm(C.byref(a), st) == 0
    for staged in (1, 0):
        epi4.tbg_set_tuning(b"igemm_staged", staged)
        res.append(f"4-warp(staged={staged}) {bench(f_epi4):6.1f} us")
    i[0] = 0; f_old(); i[0] = 0; f_new(); torch.cuda.synchronize()
    res.append("bit-identical" if torch.equal(out, out2) else f"max diff {(out.float() - out2.float()).abs().max():.3g}")
    print(f"{name:26s} B={B} | " + " | ".join(res), flush=True)


B = 32
run("mod 4x16 512->512", B, Cv.plain_geom(4, 16, 512, 512, 3))
run("mod 8x32 256->256", B, Cv.plain_geom(8, 32, 256, 256, 3))
run("mod 32x128 128->128", B, Cv.plain_geom(32, 128, 128, 128, 3))
run("up(folded) 4x16 512->256", B, Cv.up_geom(4, 16, 512, 256))
run("d down 32x128 128->128", B, Cv.down_geom(32, 128, 128, 128, 3, True))
run("d skip 32x128 128->128", B, Cv.down_geom(32, 128, 128, 128, 1, True))
run("d 4x8 512->512", B, Cv.plain_geom(4, 8, 512, 512, 3))
