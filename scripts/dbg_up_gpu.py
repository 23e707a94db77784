import sys, os, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu
from textboxgan_b200 import conv as C, fused as F, kernels as K
DEV = "cuda"
def rnd(t): return t.to(torch.bfloat16).float()
def rep(name, a, b):
    a, b = a.float().cpu(), b.float().cpu()
    dlt = (a - b).abs(); i = dlt.argmax().item(); idx = []
    for s_ in reversed(a.shape): idx.append(i % s_); i //= s_
    print(f"  {name}: max|d| {dlt.max():.4f} at {list(reversed(idx))} (a {a.flatten()[dlt.argmax()]:.4f} b {b.flatten()[dlt.argmax()]:.4f}), max|a| {a.abs().max():.3f}, mean|d| {dlt.mean():.5f}")
def run(Fn, spec, dev, x, sc, wr, nz, ns, bias, gy):
    F.clear_step_cache()
    cv = (lambda t: t.to(dev).bfloat16()) if dev == DEV else (lambda t: t.clone())
    xa = cv(x).requires_grad_(True); sa = sc.to(dev).clone().requires_grad_(True); wa = wr.to(dev).clone().requires_grad_(True)
    y = Fn.apply(xa, sa, wa, nz.to(dev), ns.to(dev), bias.to(dev), spec, math.sqrt(2))
    y.backward(cv(gy))
    F.clear_step_cache()
    return y.detach(), xa.grad, sa.grad, wa.grad
for (B, h, w, I, O) in [(3, 4, 16, 128, 64), (2, 8, 32, 256, 256)]:
    gen = torch.Generator().manual_seed(B * 7 + h + I)
    spec = C.weight_spec("upT", h, w, I, O, 3, True, "modconv")
    fold = C.weight_spec("up", h, w, I, O, 3, True, "modconv")
    x = rnd(torch.randn(B, h, w, I, generator=gen)); wr = torch.randn(3, 3, I, O, generator=gen)
    sc = torch.randn(B, I, generator=gen) * 0.2 + 1.0
    nz = torch.randn(B, 2 * h, 2 * w, generator=gen); ns = torch.tensor(0.3); bias = torch.randn(O, generator=gen)
    gy = rnd(torch.randn(B, 2 * h, 2 * w, O, generator=gen))
    ga = run(F.ModConvAct, fold, DEV, x, sc, wr, nz, ns, bias, gy)
    gb = run(F.ModUpConvAct, spec, DEV, x, sc, wr, nz, ns, bias, gy)
    with emu.emulated_kernels():
        ea = run(F.ModConvAct, fold, "cpu", x, sc, wr, nz, ns, bias, gy)
        eb = run(F.ModUpConvAct, spec, "cpu", x, sc, wr, nz, ns, bias, gy)
    print((B, h, w, I, O))
    for i, n in enumerate(("y", "gx", "gs", "gw")):
        rep(n + " gpu folded   vs emu", ga[i], ea[i]); rep(n + " gpu unfolded vs emu", gb[i], eb[i])
