import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C
def bench(fn, iters=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for kind, k, I, O in (("plain",3,128,128),("plain",3,512,512),("up",3,512,256),("up",3,128,512),("down",3,512,512),("down",1,512,512),("down",3,128,128),("plain",3,513,512)):
    spec = C.weight_spec(kind, 8, 16, I, O, k, True)
    w = torch.randn(k,k,I,O, device="cuda")
    g = torch.randn(spec.fwd_rows, spec.fwd_cols, device="cuda")
    gq = torch.randn(I, O, device="cuda")
    out = torch.zeros(k,k,I,O, device="cuda")
    t1 = bench(lambda: K.wprep(w, spec, want_adj=True, want_q=True))
    t2 = bench(lambda: K.wprep(w, spec, want_adj=False, want_q=False))
    t3 = bench(lambda: K.wfold(g, spec, gq=gq, w_raw=w, out=out))
    mb = (spec.fwd_rows*spec.fwd_cols*2*2 + w.numel()*4)/1e6
    print(f"{kind} k={k} {I}->{O}: wprep(full) {t1:.1f} us, wprep(fwd only) {t2:.1f} us, wfold {t3:.1f} us; bytes {mb:.1f} MB")
