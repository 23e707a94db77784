"""Scratch: time the training step on the GPU and print a torch.profiler kernel summary."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import lib
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.config import baseline_config
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep
from oracle import train_step as OT

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
with_ocr = (sys.argv[2] != "noocr") if len(sys.argv) > 2 else True
use_graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
do_prof = not (len(sys.argv) > 4 and sys.argv[4] == "noprof")
cfg = baseline_config(idx)
dev = "cuda:0"
G = Generator(cfg, device=dev, seed=0); D = Discriminator(cfg, device=dev, seed=1)
aster = AsterInferer(cfg, device=dev, synthetic_weights=True) if with_ocr else None
go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=dev), cfg)
ts.use_cuda_graph = use_graph
g = torch.Generator().manual_seed(4444)
real, words, labels = OT.synthetic_batch(cfg, cfg.batch_size_per_gpu, g)
real, words, labels = real.to(dev), words.to(dev), labels.to(dev)
zero = torch.zeros((), device=dev)
def step(r1=False, pl=False):
    return ts.dist_train_step(real, zero, words, labels, r1, pl, 1e-4)
for _ in range(3): step()
torch.cuda.synchronize()
lib.load().tbg_reset_launch_count()
t0 = time.time(); n = 10
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): out = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"config {idx} ocr={with_ocr} graph={use_graph}: plain step {ms:.2f} ms (wall {(time.time()-t0)/n*1e3:.2f} ms) -> {cfg.batch_size_per_gpu/ms*1e3:.1f} img/s; tbg launches/step {lib.load().tbg_launch_count()/n}")
print("losses", [float(v) for v in (*out[0], *out[1], out[2])])
for name, kw in (("pl", dict(pl=True)), ("r1+pl", dict(r1=True, pl=True))):
    step(**kw); torch.cuda.synchronize()
    e0.record(); step(**kw); e1.record(); torch.cuda.synchronize()
    print(f"  {name} step {e0.elapsed_time(e1):.2f} ms")
if not do_prof: sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/profile_table.txt", "w").write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=80, max_name_column_width=90))
# attribute GPU time to the python call sites issuing the kernels (by module of the stack top)
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70))
