"""Every tensor-core launch of one eager plain training iteration, grouped by (kernel, tag, shape): count, total time
(CUDA events around each launch), executed TF/s.  Tells which conv shapes the conv_igemm / conv_wgrad time goes to.
Usage: python scripts/conv_census.py [cfg] [per-GPU batch]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.config import baseline_config
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep
from oracle import train_step as OT

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = baseline_config(idx)
if len(sys.argv) > 2:
    cfg.batch_size_per_gpu = cfg.batch_size = int(sys.argv[2])
dev = "cuda:0"
G = Generator(cfg, device=dev, seed=0); D = Discriminator(cfg, device=dev, seed=1)
aster = AsterInferer(cfg, device=dev, synthetic_weights=True)
go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=dev), cfg)
ts.use_cuda_graph = False
g = torch.Generator().manual_seed(4444)
real, words, labels = OT.synthetic_batch(cfg, cfg.batch_size_per_gpu, g)
real, words, labels = real.to(dev), words.to(dev), labels.to(dev)
zero = torch.zeros((), device=dev)
step = lambda: ts.dist_train_step(real, zero, words, labels, False, False, 1e-4)
for _ in range(3): step()
torch.cuda.synchronize()
K.PROFILE = []
step()
torch.cuda.synchronize()
recs, K.PROFILE = K.PROFILE, None
agg = collections.OrderedDict()
for r in recs:
    name, tag, flops, e0, e1 = r[:5]
    info = r[5] if len(r) > 5 else {}
    key = (name, str(tag[0] if isinstance(tag, tuple) else tag),
           str(info.get("x_shape", "")), str(info.get("w_shape", "")), str(info.get("stride", "")), str(info.get("up", "")))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3; a[2] += flops
tot = sum(a[1] for a in agg.values())
by_tag = collections.defaultdict(float)
lines = [f"config {idx} batch {cfg.batch_size_per_gpu}: {len(recs)} recorded launches, {tot / 1e3:.3f} ms (event-timed, eager)"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    by_tag[(k[0], k[1])] += a[1]
    lines.append(f"{a[1]:8.1f} us {a[1] / tot * 100:5.1f}%  x{a[0]:<3d} {a[1] / a[0]:7.1f} us/launch {a[2] / a[1] / 1e6:7.1f} TF/s  {k[0]:11s} {k[1]:10s} x{k[2]} w{k[3]} s{k[4]} up{k[5]}")
lines.append("-- by (kernel, tag)")
for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1]):
    lines.append(f"{v:8.1f} us {v / tot * 100:5.1f}%  {k}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/conv_census.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
