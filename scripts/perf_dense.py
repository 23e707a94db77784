"""In-graph time of the mapping network's dense chain (8 x [M=128, 512 -> 512] forward, then the 8 backward steps): the
chain is captured in a CUDA graph and replayed (what the training step does).
   python scripts/perf_dense.py [M]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda"
ws = [torch.randn(512, 512, device=dev) for _ in range(8)]
bs = [torch.randn(512, device=dev) for _ in range(8)]
x0 = torch.randn(M, 512, device=dev)
g0 = torch.randn(M, 512, device=dev)
kw = dict(coef=0.01 / math.sqrt(512), bias_coef=0.01, act=1, gain=math.sqrt(2.0))


def chain():
    xs = [x0]
    for w, b in zip(ws, bs):
        xs.append(K.dense_fwd(xs[-1], w, b, **kw))
    return xs


def chain_bwd(xs):
    g = g0
    outs = []
    for i in range(7, -1, -1):
        g, gw, gb = K.dense_bwd(xs[i], ws[i], xs[i + 1], g, **kw, want_gx=True)
        outs.append((gw, gb))
    return g, outs


def timed(fn, label):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): r = fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        r = fn()
    for _ in range(5): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:60s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us per replay", flush=True)
    return r


xs = timed(chain, f"M={M}: 8-layer forward chain")
timed(lambda: chain_bwd(xs), f"M={M}: 8-layer backward chain (gpre, gx, gw, gb)")
# the discriminator head: 8192 -> 512 at M = 128
xd = torch.randn(128, 8192, device=dev); wd = torch.randn(8192, 512, device=dev)
timed(lambda: K.dense_fwd(xd, wd, None, coef=1.0, act=0), "M=128: 8192 -> 512")
