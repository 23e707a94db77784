"""World-size-2 data parallelism on CPU (gloo): two ranks holding half of the global batch each
reach the same updated weights and losses as one process holding the whole batch — the SUM (not
mean) gradient semantics of the reference (gan_losses.py:10,16; training_step.py:233-235)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, ret, shard=None, late_comm=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from common import perturbed_params, small_cfg
    from emu import emulated_kernels
    from oracle import train_step as OT
    from textboxgan_b200.discriminator import Discriminator
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.optimizers import Adam, update_optimizer_params
    from textboxgan_b200.strategy import Strategy
    from textboxgan_b200.training_step import TrainingStep

    torch.set_num_threads(2)
    gb = 4
    cfg = small_cfg(gb // world)
    strategy = Strategy(backend="gloo") if world > 1 else None
    if strategy is not None:
        cfg.attach_strategy(strategy)
    cfg.batch_size = gb
    GP, DP, g = perturbed_params(cfg)
    real, words, labels = OT.synthetic_batch(cfg, gb, g)
    draws = OT.make_draws(cfg, gb, g)
    per = gb // 2
    cfg.batch_size_per_gpu = per
    idx = rank if shard is None else shard
    sl = slice(idx * per, (idx + 1) * per)
    d = {k: (v[sl] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == gb else v) for k, v in draws.items()}
    d["noises"] = [n[sl] for n in draws["noises"]]
    with emulated_kernels():
        G = Generator(cfg, device="cpu", seed=0)
        G.load_state_dict(GP)
        D = Discriminator(cfg, device="cpu", seed=0)
        D.load_state_dict(DP)
        go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
        mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
        ts = TrainingStep(G, D, None, mk(go), mk(go), mk(do), 8, 16, torch.zeros(()), cfg)
        ts.overlap_comm = not late_comm
        out = ts.dist_train_step(real[sl], torch.zeros(()), words[sl], labels[sl], False, False, 1e-4, draws=d)
    if rank == 0:
        ret["losses"] = [float(v) for v in (*out[0], *out[1], out[2])]
        ret["G"] = G.flat.detach().clone()
        ret["D"] = D.flat.detach().clone()
        # gradient buffers the optimisers consumed (after the cross-replica SUM when world > 1)
        ret["gG"] = list(ts.g_optimizer._slots.values())[0][0].detach().clone()
        ret["gD"] = list(ts.d_optimizer._slots.values())[0][0].detach().clone()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port() -> int:
    """A port nobody listens on right now (a fixed port can collide with a socket of an earlier run in TIME_WAIT)."""
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def test_two_ranks_sum_gradients_of_their_shards():
    """The gradient buffer each optimiser consumes on 2 ranks == the SUM of the gradients two
    independent single-rank runs compute on the two shards (losses carry 1/global_batch), and the
    reported losses are the SUM of the per-shard losses (training_step.py:106-134)."""
    mgr = mp.Manager()
    r2, s0, s1 = mgr.dict(), mgr.dict(), mgr.dict()
    mp.spawn(_run, args=(2, _free_port(), r2), nprocs=2, join=True)
    mp.spawn(_run, args=(1, _free_port(), s0, 0), nprocs=1, join=True)
    mp.spawn(_run, args=(1, _free_port(), s1, 1), nprocs=1, join=True)
    for key in ("gG", "gD"):
        want = s0[key] + s1[key]
        err = ((r2[key] - want).abs().max() / (want.abs().max() + 1e-30)).item()
        assert err < 1e-5, (key, err)
    for a, b, c in zip(r2["losses"], s0["losses"], s1["losses"]):
        assert abs(a - (b + c)) < 1e-5 * max(1.0, abs(a))
    # identical replicas stay identical: rank 0's weights after the update are what a single process
    # applying the summed gradient would hold (checked through the Adam step bound)
    assert (r2["G"] - s0["G"]).abs().max().item() < 5e-3


def test_all_reduces_after_the_last_backward_pass_give_the_same_update():
    """TrainingStep.overlap_comm = False (the three cross-replica sums issued after the last backward pass instead of
    after each group's own pass) changes only the issue order: same summed gradients, same weights."""
    mgr = mp.Manager()
    a, b = mgr.dict(), mgr.dict()
    mp.spawn(_run, args=(2, _free_port(), a), nprocs=2, join=True)
    mp.spawn(_run, args=(2, _free_port(), b, None, True), nprocs=2, join=True)
    for key in ("gG", "gD", "G", "D"):
        assert torch.equal(a[key], b[key]), key
    assert a["losses"] == b["losses"]


def _init_run(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from common import small_cfg
    from emu import emulated_kernels
    from textboxgan_b200.model_loader import ModelLoader
    from textboxgan_b200.strategy import Strategy

    torch.set_num_threads(2)
    torch.manual_seed(1000 + rank)             # every rank has its own RNG stream, like independent processes do
    cfg = small_cfg(2)
    cfg.attach_strategy(Strategy(backend="gloo"))
    with emulated_kernels():
        D, G, g_clone = ModelLoader(cfg, device="cpu").initiate_models()
    ret[rank] = {"G": G.flat.detach().clone(), "D": D.flat.detach().clone(), "C": g_clone.flat.detach().clone(),
                 "w_avg": G.params["latent_encoder/w_avg"].detach().clone(),
                 "next_rand": float(torch.rand(()))}
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_start_from_identical_weights():
    """ModelLoader.initiate_models on 2 ranks with different RNG states: rank 0's initialisation is broadcast, so the
    generator, its clone and the discriminator are bit-identical across replicas (MirroredStrategy semantics); the
    user's global RNG stream is consumed, not reseeded."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_init_run, args=(2, _free_port(), ret), nprocs=2, join=True)
    for k in ("G", "D", "C", "w_avg"):
        assert torch.equal(ret[0][k], ret[1][k]), k
    assert torch.equal(ret[0]["G"], ret[0]["C"])
    assert float(ret[0]["G"].abs().sum()) > 0
    assert ret[0]["next_rand"] != ret[1]["next_rand"]          # per-rank streams stay distinct (manual_seed respected)
