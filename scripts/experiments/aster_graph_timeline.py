"""In-graph kernel list of the frozen OCR branch alone (crop-resize -> ResNet encoder -> BiLSTMs -> attention
decoder -> loss -> gradient w.r.t. the images), captured in ONE CUDA graph on one stream and replayed under
torch.profiler: every launch in issue order with its duration.  Shows what the branch costs when nothing overlaps it.
Usage: python scripts/experiments/aster_graph_timeline.py [cfg] [batch]"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.config import baseline_config
from textboxgan_b200.losses import softmax_cross_entropy_loss

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = baseline_config(idx)
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch_size_per_gpu
dev = "cuda:0"
aster = AsterInferer(cfg, device=dev, synthetic_weights=True)
g = torch.Generator().manual_seed(1)
H, W = cfg.char_height, cfg.image_width
img = torch.randn(B, 3, H, W, generator=g).to(dev).requires_grad_(True)
labels = torch.randint(1, 30, (B, cfg.max_char_number), generator=g).to(dev)


def branch():
    x = AsterInferer.convert_inputs(img, labels, 1, cfg)
    logits = aster(x)
    loss = softmax_cross_entropy_loss(logits, labels, B)
    (gi,) = torch.autograd.grad(loss, img)
    return gi


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): branch()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=s):
    out = branch()
for _ in range(3): graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): graph.replay()
e1.record(); torch.cuda.synchronize()
print(f"OCR branch alone, config {idx}, batch {B}: {e0.elapsed_time(e1) / 10:.3f} ms per replay (CUDA events)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    graph.replay()
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
t0 = evs[0][0]
prev_end = t0
tot = 0.0
for s_, e_, n in evs:
    n = re.sub(r"^void ", "", n).replace("at::native::", "")[:90]
    print(f"{(s_ - t0):9.1f} us  +gap {(s_ - prev_end):6.1f}  dur {(e_ - s_):7.1f} us  {n}")
    prev_end = e_
    tot += e_ - s_
print(f"{len(evs)} kernels, sum of durations {tot / 1e3:.3f} ms, span {(prev_end - t0) / 1e3:.3f} ms")
