"""Attribute the GPU time of one eager training step to the Python call sites that launch the kernels
(torch.profiler with stacks).  Usage (GPU box): python scripts/profile_sites.py [config_index] [steps]
Writes gpurun_out/profile_sites.txt."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.config import baseline_config
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep
from oracle import train_step as OT

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = baseline_config(idx)
dev = "cuda:0"
G = Generator(cfg, device=dev, seed=0); D = Discriminator(cfg, device=dev, seed=1)
aster = AsterInferer(cfg, device=dev, synthetic_weights=True)
go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=dev), cfg)
g = torch.Generator().manual_seed(4444)
real, words, labels = OT.synthetic_batch(cfg, cfg.batch_size_per_gpu, g)
real, words, labels = real.to(dev), words.to(dev), labels.to(dev)
zero = torch.zeros((), device=dev)
step = lambda: ts.dist_train_step(real, zero, words, labels, False, False, 1e-4)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    for _ in range(nsteps): step()
    torch.cuda.synchronize()
sites = collections.defaultdict(lambda: [0.0, 0, collections.Counter()])
for e in prof.events():
    t = getattr(e, "self_device_time_total", 0.0)
    if not t:
        continue
    site = None
    for fr in (e.stack or []):
        if "textboxgan_b200/" in fr:
            site = fr.split("textboxgan_b200/")[-1]
            break
    if site is None:
        site = "<no python frame> " + e.name
    s = sites[site]
    s[0] += t; s[1] += 1; s[2][e.name] += 1
tot = sum(v[0] for v in sites.values())
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/profile_sites.txt", "w") as f:
    f.write(f"config {idx}: {tot / nsteps / 1e3:.3f} ms GPU self time per eager step\n")
    for k, v in sorted(sites.items(), key=lambda kv: -kv[1][0])[:140]:
        ops = ", ".join(f"{n}x{c // nsteps if c >= nsteps else c}" for n, c in v[2].most_common(3))
        f.write(f"{v[0] / tot * 100:6.2f}% {v[0] / nsteps:9.1f}us {v[1] // nsteps:5d}  {k[:90]:90s} {ops[:100]}\n")
print(open("gpurun_out/profile_sites.txt").read()[:6000])
