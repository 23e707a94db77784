#!/usr/bin/env python
"""bench.py — images/sec of the TextBoxGAN training step (G + D + OCR loss) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C]

Workload: N = 1 defaults to BASELINE.json ``configs[2]`` (256x64, batch 64, max_char_number 12, z 512, style mixing
on, R1 penalty), N > 1 to ``configs[3]`` (the same ladder data-parallel, 256 / 8 = 32 images per GPU, weak scaling).
The K timed steps follow the reference's lazy-regularisation schedule (train.py:182-183: iteration i runs the
path-length regulariser when (i+1) % 8 == 0 and the R1 penalty when (i+1) % 16 == 0), so ``value`` is the
throughput of real training iterations; ``plain_step`` is the rate of the non-regularised iteration alone and
``mix16`` the rate over whole 16-step cycles.

N > 1 is launched by the driver under ``torch.distributed.run`` (one rank per GPU, NCCL); the
step is pure data parallel on the batch axis (weak scaling: the per-GPU batch is fixed).

One JSON line on rank 0 with the contract keys plus
  roofline      modulated-conv2d tensor-core roofline of ``conv_igemm_kernel``: every modconv launch configuration of
                one iteration re-timed in isolation with CUDA events (algorithmic FLOPs per SURVEY.md §8d; peak =
                MEASURED_PEAKS.json bf16_tflops, the burst figure), ``in_step`` = the same launches timed inside the
                iteration against bf16_tflops_sustained, ``traffic`` = ncu DRAM bytes of the dominant launch,
  cpu_baseline  the oracle's restatement of the reference's ``cpu_only`` path timed on this box's
                host cores on a bounded sample (rank 0, N = 1 only),
  e2e           the same metric through the public API with pinned-host inputs copied in and
                the seven loss scalars read back every step,
  gpu_launches  launches of this repo's kernels inside the timed region,
  plain_step    rate of the non-regularised iteration alone,
  mix16         rate over whole 16-iteration schedule cycles (14 plain, one path-length, one path-length + R1).
``--impl reference`` times the oracle port of the reference's CPU path (TensorFlow 2.8 is not
installable offline — see DESIGN.md) with all host threads on a bounded sample of the same workload and schedule,
and prints the steps / warm-up / batch that actually ran.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_step_images_per_sec"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clocks and clock-event (throttle) reasons through NVML during the timed region."""

    def __init__(self, gpu_index: int, period_s: float = 0.05):
        self.gpu = gpu_index
        self.period = period_s
        self.samples = []
        self.reasons = set()
        self.smax = None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def _run(self):
        nv = self.nv
        bits = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as ex:  # pragma: no cover
                self.err = repr(ex)
                break
            time.sleep(self.period)

    def stop(self) -> dict:
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
               "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


def _cpu_oracle_rate(cfg_index: int, sample_batch: int, steps: int, threads: int):
    """Images/sec of the oracle's restatement of the reference ``cpu_only`` training step
    (non-fused modulated conv, upfirdn_2d_ref, three TF-style Adam updates) on the host cores."""
    import copy

    import torch

    from oracle import aster as OA
    from oracle import stylegan as OS
    from oracle import train_step as OT
    from textboxgan_b200.config import baseline_config

    torch.set_num_threads(threads)
    cfg = baseline_config(cfg_index)
    cfg.batch_size_per_gpu = sample_batch
    cfg.batch_size = sample_batch
    g = torch.Generator().manual_seed(4444)
    GP = OS.init_generator_params(cfg, g)
    DP = OS.init_discriminator_params(cfg, g)
    st = OT.StepState(GP, DP, OA.init_aster_params(), OT.make_adam(cfg.g_opt), OT.make_adam(cfg.g_opt),
                      OT.make_adam(cfg.d_opt), torch.zeros(()))
    real, words, labels = OT.synthetic_batch(cfg, sample_batch, g)
    times = []
    for i in range(steps + 1):
        draws = OT.make_draws(cfg, sample_batch, g)
        t0 = time.perf_counter()
        OT.train_step(st, cfg, real, torch.zeros(()), words, labels, False, False, 1e-4, draws, fused=False)
        times.append(time.perf_counter() - t0)
    timed = times[1:] if steps >= 1 else times
    per_step = sum(timed) / len(timed)
    return sample_batch / per_step, per_step


def synthetic_inputs(cfg, batch: int, seed: int):
    """Seeded synthetic word batch with the loader's tensor contract (dataset_utils/training_data_loader.py:56-97,
    SURVEY.md §8d): words of length U{1..mcn} over the 69 main characters (pad 0), their ASTER labels (pad 1), uniform
    "real" images zeroed right of the word.  Product code only (the oracle is not involved in the measured arm)."""
    import torch

    from textboxgan_b200.char_tokens import main_to_aster_ids
    from textboxgan_b200.utils import mask_text_box

    g = torch.Generator().manual_seed(seed)
    mcn = cfg.max_char_number
    lens = torch.randint(1, mcn + 1, (batch,), generator=g)
    chars = torch.randint(1, 70, (batch, mcn), generator=g)
    words = torch.where(torch.arange(mcn)[None, :] < lens[:, None], chars, torch.zeros_like(chars)).to(torch.int32)
    labels = torch.from_numpy(main_to_aster_ids(words.numpy())).to(torch.int32)
    real = torch.rand(batch, 3, cfg.char_height, cfg.image_width, generator=g) * 2 - 1
    real = mask_text_box(real, words, cfg.char_width)
    return real.contiguous(), words, labels


def _schedule(i: int, cfg):
    """train.py:182-183 — (do_r1_reg, do_pl_reg) of training iteration ``i`` (0-based)."""
    return (i + 1) % cfg.d_opt["reg_interval"] == 0, (i + 1) % cfg.g_opt["reg_interval"] == 0


def _bench_config(args, n_gpus: int):
    """BASELINE.json configs[C] with the per-GPU batch of the run: configs[3] is quoted on 8 GPUs (256 / 8 = 32 images
    per GPU) and keeps that per-GPU batch at every N (weak scaling); configs[4] likewise 512 / 8 = 64."""
    from textboxgan_b200.config import baseline_config

    c = args.config if args.config is not None else (2 if n_gpus == 1 else 3)
    cfg = baseline_config(c, n_gpus=8 if c in (3, 4) else 1)
    return c, cfg


def run_reference(args) -> None:
    """The reference's CPU path (oracle port: TensorFlow 2.8 is not installable offline) on the host
    cores.  One "step" = one training iteration of the same workload shape and schedule on a bounded sample batch;
    the line states the steps, warm-up and batch that actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import aster as OA
    from oracle import stylegan as OS
    from oracle import train_step as OT

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    c, cfg = _bench_config(args, max(1, args.gpus))
    workload = _config_dict(c, cfg, max(1, args.gpus))
    sample_batch = 4
    cfg.batch_size_per_gpu = sample_batch
    cfg.batch_size = sample_batch
    g = torch.Generator().manual_seed(4444)
    st = OT.StepState(OS.init_generator_params(cfg, g), OS.init_discriminator_params(cfg, g), OA.init_aster_params(),
                      OT.make_adam(cfg.g_opt), OT.make_adam(cfg.g_opt), OT.make_adam(cfg.d_opt), torch.zeros(()))
    real, words, labels = OT.synthetic_batch(cfg, sample_batch, g)

    def step(i):
        do_r1, do_pl = _schedule(i, cfg)
        draws = OT.make_draws(cfg, sample_batch, g, with_pl=do_pl)
        t0 = time.perf_counter()
        OT.train_step(st, cfg, real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws, fused=False)
        return time.perf_counter() - t0

    # bounded: at most 16 timed iterations (one full schedule cycle) and ~150 s of CPU work
    budget_s = 150.0
    warm = max(1, min(args.warmup, 2))
    t_start = time.perf_counter()
    for _ in range(warm):
        step(0)                                   # plain iterations
    k_max = max(1, min(args.steps, 16))
    times = []
    for i in range(k_max):
        times.append(step(i))
        if time.perf_counter() - t_start > budget_s:
            break
    k = len(times)
    per_step = sum(times) / k
    value = sample_batch / per_step
    n_pl = sum(1 for i in range(k) if _schedule(i, cfg)[1])
    n_r1 = sum(1 for i in range(k) if _schedule(i, cfg)[0])
    sample = (f"{k} training iterations (schedule of train.py:182-183: {k - n_pl} plain, {n_pl - n_r1} path-length, "
              f"{n_r1} path-length+R1) after {warm} warm-up, at batch {sample_batch} of the configs[{c}] shape "
              f"({cfg.image_width}x{cfg.char_height}, mcn {cfg.max_char_number}, z {cfg.z_dim}), fp32, {threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": warm, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload, sample_batch=sample_batch, requested_steps=args.steps, requested_warmup=args.warmup),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference cpu_only path (TensorFlow 2.8 not installable offline); steps / warmup / "
                "sample_batch are what actually ran (bounded sample of the workload in config)",
    }
    print(json.dumps(line), flush=True)


def _config_dict(c: int, cfg, n_gpus: int):
    return {
        "workload": (f"BASELINE.json configs[{c}]: TextBoxGAN training iteration G+D+OCR loss (style mixing on, lazy "
                     f"regularisation schedule of train.py:182-183: path length every 8th, R1 every 16th iteration), "
                     f"per-GPU batch {cfg.batch_size_per_gpu}, max_char_number={cfg.max_char_number}, z_dim={cfg.z_dim}, "
                     f"{cfg.image_width}x{cfg.char_height}"),
        "baseline_config_index": c,
        "global_batch": cfg.batch_size_per_gpu * n_gpus,
        "per_gpu_batch": cfg.batch_size_per_gpu,
        "image": f"{cfg.image_width}x{cfg.char_height}",
        "parallelism": f"dp{n_gpus}",
        "l2": "per-step working set (parameters + Adam state + activations, > 1 GB) exceeds the 126 MB L2; no explicit flush",
    }


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib
    from textboxgan_b200.aster_inferer import AsterInferer
    from textboxgan_b200.discriminator import Discriminator
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.optimizers import Adam, update_optimizer_params
    from textboxgan_b200.strategy import Strategy
    from textboxgan_b200.training_step import TrainingStep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    lib.load()
    strategy = Strategy()
    rank, world = strategy.rank, strategy.world_size
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    c, cfg = _bench_config(args, world)
    cfg.attach_strategy(strategy)
    B = cfg.batch_size_per_gpu

    G = Generator(cfg, device=dev, seed=1)
    D = Discriminator(cfg, device=dev, seed=2)
    g_clone = Generator(cfg, device=dev, seed=1)
    aster = AsterInferer(cfg, device=dev, synthetic_weights=True)
    go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), cfg.g_opt["reg_interval"], cfg.d_opt["reg_interval"],
                      torch.zeros((), device=dev), cfg)
    ts.use_cuda_graph = not args.eager      # whole iteration replayed from a CUDA graph (one per schedule variant)
    if getattr(args, "late_comm", False):
        ts.overlap_comm = False

    real_h, words_h, labels_h = synthetic_inputs(cfg, B, 4444 + rank)          # reference shuffle_seed, config.py:114
    real_h, words_h, labels_h = real_h.pin_memory(), words_h.pin_memory(), labels_h.pin_memory()
    real, words, labels = real_h.to(dev), words_h.to(dev), labels_h.to(dev)
    zero = torch.zeros((), device=dev)
    loss_host = torch.empty(7, dtype=torch.float32).pin_memory()

    def step_resident(i, plain=False):
        do_r1, do_pl = (False, False) if plain else _schedule(i, cfg)
        out = ts.dist_train_step(real, zero, words, labels, do_r1, do_pl, cfg.ocr_loss_weight)
        g_clone.set_as_moving_average_of(G)      # train.py:208 — part of every iteration
        return out

    # end-to-end input path = the Trainer's (train.py): pinned host batches through the package's DevicePrefetcher — the
    # copy of batch i+1 runs on its copy stream while iteration i computes; every iteration still copies its own inputs
    from textboxgan_b200.prefetch import DevicePrefetcher

    def host_batches():
        while True:
            yield real_h, words_h, labels_h

    feed = DevicePrefetcher(host_batches(), dev)

    def step_e2e(i, plain=False):
        do_r1, do_pl = (False, False) if plain else _schedule(i, cfg)
        r, w, l = next(feed)
        out = ts.dist_train_step(r, zero, w, l, do_r1, do_pl, cfg.ocr_loss_weight)
        g_clone.set_as_moving_average_of(G)
        packed = torch.stack([*out[0], *out[1], out[2]]).float()
        loss_host.copy_(packed)                   # device -> host read of the step's losses (synchronising)
        return loss_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, **kw):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i, **kw)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        # the step time that counts is the slower of device time and host wall time per step
        return max(ms, wall * 1e3) / steps, ms / steps, wall * 1e3 / steps

    # every schedule variant: first use runs eagerly, the second captures its graph, then replays
    plain_only = bool(args.plain_only)
    variants = [(False, False)] if plain_only else [(False, False), (False, True), (True, True)]
    for do_r1, do_pl in variants:
        for _ in range(3):
            ts.dist_train_step(real, zero, words, labels, do_r1, do_pl, cfg.ocr_loss_weight)
            g_clone.set_as_moving_average_of(G)
    W = max(args.warmup, 3)
    for i in range(W):
        step_resident(i, plain=plain_only)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.load().tbg_reset_launch_count()
    ms_eff, ms_step, wall_step = timed(step_resident, args.steps, plain=plain_only)
    # host-issued launches (EMA, eager mode) + launches replayed from the captured graphs
    launches = int(lib.load().tbg_launch_count())
    for i in range(args.steps):
        do_r1, do_pl = (False, False) if plain_only else _schedule(i, cfg)
        launches += ts.graph_launches(do_r1, do_pl)
    clocks = sampler.stop() if rank == 0 else None

    step_e2e(0, plain=True)
    ms_e2e, _, _ = timed(step_e2e, args.steps, plain=plain_only)
    h2d = real_h.numel() * 4 + words_h.numel() * 4 + labels_h.numel() * 4
    d2h = 7 * 4

    # secondary rates: the plain (non-regularised) iteration alone, and whole 16-step schedule cycles
    n_plain = 32
    ms_plain, _, _ = timed(step_resident, n_plain, plain=True)
    mix16 = None
    if not plain_only:
        cycles = 2
        ms_mix, _, _ = timed(step_resident, 16 * cycles)
        mix16 = {"value": B * world / (ms_mix * 1e-3), "unit": UNIT, "ms_per_16_steps": ms_mix * 16, "cycles": cycles,
                 "schedule": "14 plain + 1 path-length + 1 path-length+R1 iteration (train.py:182-183), CUDA-graph replay"}

    # ---- roofline pass: CUDA events around every tensor-core launch of two instrumented plain steps ----
    roof = None
    if args.no_roofline:
        if rank == 0:
            gb = B * world
            print(json.dumps({"metric": METRIC, "value": gb / (ms_eff * 1e-3), "unit": UNIT, "n_gpus": world,
                              "steps": args.steps, "ms_per_step": ms_eff, "per_gpu_batch": B,
                              "plain_step": {"value": gb / (ms_plain * 1e-3), "ms_per_step": ms_plain},
                              "config": _config_dict(c, cfg, world)["workload"]}), flush=True)
        if world > 1:
            ts._graphs.clear()
            _teardown(dist)
        return
    # every rank runs the two instrumented eager steps (they contain collectives); rank 0 records
    K.PROFILE = [] if rank == 0 else None
    ts.use_cuda_graph = False                     # events around every launch need eager launches
    for i in range(2):
        step_resident(i, plain=True)
    torch.cuda.synchronize()
    recs = K.PROFILE
    K.PROFILE = None
    if rank == 0:
        roof = _roofline(recs, K, dev)

    if world > 1:
        # clean NCCL teardown: all ranks idle, then destroy (no collective may still be in flight on any stream)
        barrier()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, per = _cpu_oracle_rate(c, 4, 1, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"1 timed plain training iteration (after 1 warm-up) at batch 4 of the configs[{c}] shape, "
                             f"oracle restatement of the reference cpu_only path, fp32, {threads} host threads"}
        # single-GPU rate at the per-GPU batch of the multi-GPU workload (configs[3]: 32 images per GPU), so that weak
        # scaling can be read against the same per-GPU work; measured by a child process after this one is done
        base = None
        if world == 1 and c == 2 and not args.no_scaling_base:
            base = _scaling_base_subprocess()
        gb = B * world
        n_pl = sum(1 for i in range(args.steps) if not plain_only and _schedule(i, cfg)[1])
        n_r1 = sum(1 for i in range(args.steps) if not plain_only and _schedule(i, cfg)[0])
        line = {
            "metric": METRIC, "value": gb / (ms_eff * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_eff, "device_ms_per_step": ms_step,
            "host_wall_ms_per_step": wall_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (tensor-core operands; fp32 accumulate, parameters, optimiser state; small dense layers tf32)",
            "data": "synthetic", "config": _config_dict(c, cfg, world),
            "steps_by_kind": {"plain": args.steps - n_pl, "path_length": n_pl - n_r1, "path_length_r1": n_r1},
            "clocks": clocks,
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "plain_step": {"value": gb / (ms_plain * 1e-3), "unit": UNIT, "ms_per_step": ms_plain, "steps": n_plain},
            "mix16": mix16,
            "weak_scaling_base": base,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graphs hold NCCL kernel nodes and communicator references: release them before the process group
        ts._graphs.clear()
        del ts
        _teardown(dist)


def _scaling_base_subprocess(timeout_s: int = 240) -> dict:
    """`bench.py --config 3 --gpus 1` (the multi-GPU workload's per-GPU batch on one GPU) in a child process."""
    import subprocess

    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", "3", "--steps", "32", "--warmup", "3",
                            "--no-roofline", "--no-cpu-baseline"], capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
        d = json.loads(lines[-1])
        return {"value": d["value"], "unit": UNIT, "ms_per_step": d["ms_per_step"], "per_gpu_batch": d["per_gpu_batch"],
                "plain_step": d["plain_step"], "workload": d["config"],
                "note": "one GPU at the per-GPU batch of BASELINE configs[3] (256 / 8): the base of the weak-scaling runs"}
    except Exception as ex:   # timeout, JSON error, ...
        return {"error": repr(ex)[:300]}


def _teardown(dist) -> None:
    """NCCL collectives were captured inside CUDA graphs; in round 1 ``destroy_process_group()`` hung at exit with
    those graphs alive.  Drop every captured graph first (they hold NCCL kernel nodes and communicator
    references), synchronise, then destroy; a watchdog exits the process if NCCL still does not come back."""
    import gc
    import threading

    import torch

    def _bail():
        os._exit(0)

    t = threading.Timer(20.0, _bail)
    t.daemon = True
    t.start()
    gc.collect()
    torch.cuda.synchronize()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


def _roofline(recs, K, dev):
    """Tensor-core roofline of ``conv_igemm_kernel`` on the modulated-conv launches of one plain iteration."""
    import torch

    agg = {}
    recipes = {}
    for rec in recs:
        name, tag, flops, e0, e1 = rec[:5]
        t, frac = tag if isinstance(tag, tuple) else (str(tag), 1.0)
        a = agg.setdefault((name, t), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1) * 1e-3
        a[2] += flops * frac
        a[3] += flops
        if name == "conv_igemm" and t == "modconv" and len(rec) > 5:
            key = repr(sorted(rec[5].items()))
            r = recipes.setdefault(key, [rec[5], 0, flops, frac])
            r[1] += 1
    # Isolated kernel timing of every modulated-conv launch configuration of the step: the same
    # launch re-issued back-to-back over rotating inputs larger than L2, CUDA events on the
    # launching stream (in-step per-launch events also include host launch gaps in eager mode).
    iso_secs = iso_algo = iso_exec = 0.0
    iso_launches = 0
    per_shape = []
    for recipe, count, flops, frac in recipes.values():
        xs_bytes = 2
        for d_ in recipe["x_shape"]:
            xs_bytes *= d_
        n_rot = max(2, min(48, int(300e6 // xs_bytes) + 1))
        xs = [torch.randn(recipe["x_shape"], device=dev).to(torch.bfloat16) for _ in range(n_rot)]
        wt = (torch.randn(recipe["w_shape"], device=dev) / recipe["w_shape"][1] ** 0.5).to(torch.bfloat16)
        Bq = recipe["x_shape"][0]
        up_ = recipe["up"]
        cout = recipe["w_shape"][0] // ((1 + up_[0]) * (1 + up_[1]))
        oh, ow = recipe["Ho"] * (1 + up_[0]), recipe["Wo"] * (1 + up_[1])
        kw = dict(Ho=recipe["Ho"], Wo=recipe["Wo"], taps=recipe["taps"], pad=recipe["pad"], stride=recipe["stride"],
                  up=up_, act=recipe["act"], act_gain=recipe["act_gain"], res_scale=recipe["res_scale"],
                  res_first=recipe["res_first"], out_fp32=recipe["out_fp32"], tap_mask=recipe.get("tap_mask"))
        if recipe["has_scale"]:
            kw["col_scale"] = torch.rand(Bq, cout, device=dev) + 0.5
        if recipe["has_bias"]:
            kw["bias"] = torch.randn(cout, device=dev)
        if recipe["has_noise"]:
            kw["noise"] = torch.randn(Bq, oh, ow, device=dev)
            kw["noise_strength"] = torch.ones(1, device=dev)
        out_t = torch.empty(Bq, oh, ow, cout, device=dev, dtype=torch.float32 if recipe["out_fp32"] else torch.bfloat16)
        for i in range(3):
            K.conv2d_igemm(xs[i % n_rot], wt, out=out_t, **kw)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        ea.record()
        for i in range(iters):
            K.conv2d_igemm(xs[i % n_rot], wt, out=out_t, **kw)
        eb.record()
        torch.cuda.synchronize()
        per = ea.elapsed_time(eb) * 1e-3 / iters
        iso_secs += per * count
        iso_algo += flops * frac * count
        iso_exec += flops * count
        iso_launches += count
        per_shape.append({"x": list(recipe["x_shape"]), "w": list(recipe["w_shape"]), "taps": list(recipe["taps"]),
                          "stride": list(recipe["stride"]), "up": list(up_), "launches_per_step": count / 2,
                          "us": per * 1e6, "algorithmic_tflops": flops * frac / per / 1e12,
                          "algorithmic_gflop": flops * frac / 1e9})
        del xs
    peaks, which = _peaks()
    burst = float(peaks.get("bf16_tflops", 1590.0))
    sustained = float(peaks.get("bf16_tflops_sustained", burst))
    n, secs, algo, execd = agg.get(("conv_igemm", "modconv"), [0, 1e-9, 0.0, 0.0])
    achieved = iso_algo / max(iso_secs, 1e-12) / 1e12
    per_shape.sort(key=lambda r: -r["algorithmic_gflop"] * r["launches_per_step"])
    return {
        "bound": "tensor",
        "kernel": "tbg_conv2d_igemm entry on the modulated conv2d forward + input-gradient launches of one iteration "
                  "(conv3x3_halo_kernel for the 3x3 stride-1 layers, conv_igemm_kernel for the strided ones)",
        "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
        # the single largest launch of that set (conv3x3_halo_kernel on the top-resolution layer), same timing method
        "dominant_launch": ({"x": per_shape[0]["x"], "w": per_shape[0]["w"], "us": per_shape[0]["us"],
                             "achieved": per_shape[0]["algorithmic_tflops"],
                             "frac": per_shape[0]["algorithmic_tflops"] / burst} if per_shape else None),
        "peak_source": f"{which} bf16_tflops (burst: the launches are timed in isolation), MEASURED_PEAKS.json",
        "method": "every modconv launch configuration of one plain iteration re-issued 20x back-to-back over rotating "
                  "inputs > L2, CUDA events on the launching stream; FLOPs = algorithmic (SURVEY 8d); achieved = sum of "
                  "FLOPs / sum of launch durations over one iteration's launches",
        "executed_tflops": iso_exec / max(iso_secs, 1e-12) / 1e12,
        "launches_per_step": iso_launches / 2, "avg_launch_us": iso_secs / max(iso_launches, 1) * 1e6,
        "in_step": {"achieved": algo / secs / 1e12, "peak": sustained, "frac": algo / secs / 1e12 / sustained,
                    "note": "CUDA events around the same launches inside the (eager) iteration, against "
                            "bf16_tflops_sustained; includes host launch gaps"},
        "traffic": TRAFFIC.get("bytes"), "traffic_note": TRAFFIC.get("note"),
        "per_shape": per_shape[:8],
        "by_kernel_in_step": {f"{k[0]}:{k[1]}": {"launches_per_step": v[0] / 2, "ms_per_step": v[1] * 1e3 / 2,
                                                 "algorithmic_tflops": v[2] / max(v[1], 1e-12) / 1e12,
                                                 "executed_tflops": v[3] / max(v[1], 1e-12) / 1e12}
                              for k, v in sorted(agg.items())},
    }


# DRAM traffic of the dominant modulated-conv launch (conv_1 of the top block: 3x3, 128 -> 128 channels on the
# B x 64 x 256 grid), dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture; filled from
# profiles/ (see profiles/README.md) — None until that capture exists for the current kernel.
TRAFFIC = {"bytes": None, "note": "no ncu --set full capture of the current kernel committed yet"}
try:
    _t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    TRAFFIC = {"bytes": _t["dram_bytes_per_launch"], "note": _t["note"]}
except Exception:
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=208, help="timed iterations (default: 13 schedule cycles)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None,
                    help="BASELINE.json configs index (default: 2 on one GPU, 3 on several)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--plain-only", action="store_true", help="time non-regularised iterations only")
    ap.add_argument("--no-roofline", action="store_true", help="short line: skip the roofline / cpu_baseline passes")
    ap.add_argument("--late-comm", action="store_true",
                    help="experiment: all gradient all-reduces after the last backward pass (TrainingStep.overlap_comm = False)")
    ap.add_argument("--no-scaling-base", action="store_true", help="skip the configs[3]-per-GPU-batch child measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
