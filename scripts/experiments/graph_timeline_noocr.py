"""Kernel timeline of CUDA-graph replays of the training step (torch.profiler, CUDA activity only):
per-kernel-name busy time inside the graph, idle gaps, concurrency.
Usage: python scripts/graph_timeline.py [cfg] [steps] [plain|pl|r1pl] [per-GPU batch]"""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.config import baseline_config
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep
from oracle import train_step as OT

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = baseline_config(idx)
if len(sys.argv) > 4:
    cfg.batch_size_per_gpu = cfg.batch_size = int(sys.argv[4])
dev = "cuda:0"
G = Generator(cfg, device=dev, seed=0); D = Discriminator(cfg, device=dev, seed=1)
aster = AsterInferer(cfg, device=dev, synthetic_weights=True) if os.environ.get("NO_OCR") is None else None
go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=dev), cfg)
ts.use_cuda_graph = True
g = torch.Generator().manual_seed(4444)
real, words, labels = OT.synthetic_batch(cfg, cfg.batch_size_per_gpu, g)
real, words, labels = real.to(dev), words.to(dev), labels.to(dev)
zero = torch.zeros((), device=dev)
variant = sys.argv[3] if len(sys.argv) > 3 else "plain"
do_r1, do_pl = {"plain": (False, False), "pl": (False, True), "r1pl": (True, True)}[variant]
step = lambda: ts.dist_train_step(real, zero, words, labels, do_r1, do_pl, 1e-4)
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(nsteps): step()
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)
busy = 0.0; cur_s, cur_e = evs[0][0], evs[0][1]
for s, e, _ in evs[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in evs:
    n = re.sub(r"^void ", "", n)
    n = n.replace("at::native::", "").replace("(anonymous namespace)::", "")[:100]
    agg[n][0] += 1; agg[n][1] += e - s
tot = sum(v[1] for v in agg.values())
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/graph_timeline.txt", "w") as f:
    f.write(f"config {idx} [{variant}]: {nsteps} graph replays, span {(t1 - t0) / nsteps / 1e3:.3f} ms/step, GPU busy (union) {busy / nsteps / 1e3:.3f} ms/step, "
            f"sum of kernel durations {tot / nsteps / 1e3:.3f} ms/step, {len(evs) / nsteps:.0f} kernels/step\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        f.write(f"{v[1] / tot * 100:6.2f}% {v[0] / nsteps:7.1f}/step {v[1] / v[0]:8.1f}us avg {v[1] / nsteps / 1e3:7.3f} ms/step  {k}\n")
print(open("gpurun_out/graph_timeline.txt").read())
