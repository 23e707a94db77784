"""CPU census of the aten ops a plain training step dispatches outside the C-ABI kernels, by call
site (TorchDispatchMode over the emulated-kernel host path of tests/emu.py).  Each non-view op is
one (or more) library kernel launch on the GPU.  Usage: python scripts/op_census.py"""
import collections, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
from common import perturbed_params, small_cfg
from emu import emulated_kernels
from oracle import train_step as OT
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep

VIEW = {"view", "reshape", "_unsafe_view", "expand", "permute", "transpose", "t", "slice", "select", "unsqueeze",
        "squeeze", "as_strided", "detach", "alias", "unbind", "split", "_reshape_alias", "unfold", "narrow",
        "empty", "empty_like", "empty_strided", "new_empty", "size", "stride", "is_same_size", "sym_size",
        "lift_fresh", "_local_scalar_dense", "new_empty_strided", "split_with_sizes", "chunk", "diagonal"}

class Census(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.sites = collections.defaultdict(collections.Counter)
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.overloadpacket.__name__
        if name not in VIEW:
            site, in_emu = None, False
            for fr in reversed(traceback.extract_stack(limit=40)):
                fn = fr.filename
                if fn.endswith("tests/emu.py"):
                    in_emu = True; break
                if "textboxgan_b200/" in fn:
                    site = f"{fn.split('textboxgan_b200/')[-1]}:{fr.lineno} {fr.name}"; break
            if not in_emu:
                self.sites[site or "<other>"][name] += 1
        return func(*args, **(kwargs or {}))

cfg = small_cfg(4)
GP, DP, g = perturbed_params(cfg)
real, words, labels = OT.synthetic_batch(cfg, 4, g)
with emulated_kernels():
    G = Generator(cfg, device="cpu", seed=0); G.load_state_dict(GP)
    D = Discriminator(cfg, device="cpu", seed=0); D.load_state_dict(DP)
    aster = AsterInferer(cfg, device="cpu", synthetic_weights=True)
    go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros(()), cfg)
    ts.dist_train_step(real, torch.zeros(()), words, labels, False, False, 1e-4)
    c = Census()
    with c:
        ts.dist_train_step(real, torch.zeros(()), words, labels, False, False, 1e-4)
tot = sum(sum(v.values()) for v in c.sites.values())
print(f"{tot} non-view aten ops per plain step (tiny ladder: {len(cfg.generator_resolutions) - 1} synthesis blocks)")
for site, ops in sorted(c.sites.items(), key=lambda kv: -sum(kv[1].values())):
    print(f"{sum(ops.values()):5d}  {site:60s} " + ", ".join(f"{k}x{v}" for k, v in ops.most_common(8)))
