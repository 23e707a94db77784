#!/usr/bin/env python
"""bench.py — images/sec of the TextBoxGAN training step (G + D + OCR loss) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1]

N > 1 is launched by the driver under ``torch.distributed.run`` (one rank per GPU, NCCL); the
step is pure data parallel on the batch axis (weak scaling: the per-GPU batch is fixed).

One JSON line on rank 0 with the contract keys plus
  roofline      modulated-conv2d tensor-core roofline of ``conv_igemm_kernel`` measured live with
                CUDA events around every launch of an instrumented step (algorithmic FLOPs per
                SURVEY.md §8d; peak = MEASURED_PEAKS.json bf16_tflops_sustained),
  cpu_baseline  the oracle's restatement of the reference's ``cpu_only`` path timed on this box's
                host cores on a bounded sample (rank 0, N = 1 only),
  e2e           the same metric through the public API with pinned-host inputs copied in and
                the seven loss scalars read back every step,
  gpu_launches  launches of this repo's kernels inside the timed region.
``--impl reference`` times the oracle port of the reference's CPU path (TensorFlow 2.8 is not
installable offline — see DESIGN.md) with all host threads on a bounded sample.
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


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    sample_batch = 4
    value, per_step = _cpu_oracle_rate(args.config, sample_batch, max(1, min(args.steps, 3)), threads)
    from textboxgan_b200.config import baseline_config

    cfg = baseline_config(args.config)
    sample = (f"{max(1, min(args.steps, 3))} plain training steps at batch {sample_batch} of the config-{args.config} "
              f"shape ({cfg.image_width}x{cfg.char_height}, z={cfg.z_dim}) on {threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config_dict(cfg, args, n_gpus=max(1, args.gpus)),     # same workload description as our arm
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference cpu_only path (TensorFlow 2.8 not installable offline)",
    }
    print(json.dumps(line), flush=True)


def _config_dict(cfg, args, n_gpus):
    return {
        "workload": (f"BASELINE.json configs[{args.config}]: TextBoxGAN training step G+D+OCR loss, per-GPU batch "
                     f"{cfg.batch_size_per_gpu}, max_char_number={cfg.max_char_number}, z_dim={cfg.z_dim}, "
                     f"{cfg.image_width}x{cfg.char_height}, plain (non-regularised) step"),
        "global_batch": cfg.batch_size_per_gpu * n_gpus,
        "per_gpu_batch": cfg.batch_size_per_gpu,
        "image": f"{cfg.image_width}x{cfg.char_height}",
        "parallelism": f"dp{n_gpus}",
        "l2": "per-step working set (parameters + Adam state + activations, > 0.5 GB) exceeds the 126 MB L2; no explicit flush",
    }


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from textboxgan_b200 import kernels as K
    from textboxgan_b200 import lib
    from textboxgan_b200.aster_inferer import AsterInferer
    from textboxgan_b200.config import baseline_config
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

    # per-GPU batch fixed (weak scaling): configs 3 and 4 are quoted on 8 GPUs, i.e. batch/8 per GPU
    cfg = baseline_config(args.config, n_gpus=8 if args.config in (3, 4) else 1)
    cfg.attach_strategy(strategy)
    B = cfg.batch_size_per_gpu

    G = Generator(cfg, device=dev, seed=1)
    D = Discriminator(cfg, device=dev, seed=2)
    g_clone = Generator(cfg, device=dev, seed=1)
    aster = AsterInferer(cfg, device=dev)
    go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), cfg.g_opt["reg_interval"], cfg.d_opt["reg_interval"],
                      torch.zeros((), device=dev), cfg)
    ts.use_cuda_graph = not args.eager      # whole iteration replayed from a CUDA graph

    real_h, words_h, labels_h = synthetic_inputs(cfg, B, 4444 + rank)          # reference shuffle_seed, config.py:114
    real_h, words_h, labels_h = real_h.pin_memory(), words_h.pin_memory(), labels_h.pin_memory()
    real, words, labels = real_h.to(dev), words_h.to(dev), labels_h.to(dev)
    zero = torch.zeros((), device=dev)

    def step_resident():
        out = ts.dist_train_step(real, zero, words, labels, False, False, cfg.ocr_loss_weight)
        g_clone.set_as_moving_average_of(G)      # train.py:208 — part of every iteration
        return out

    def step_e2e():
        r = real_h.to(dev, non_blocking=True)
        w = words_h.to(dev, non_blocking=True)
        l = labels_h.to(dev, non_blocking=True)
        out = ts.dist_train_step(r, zero, w, l, False, False, cfg.ocr_loss_weight)
        g_clone.set_as_moving_average_of(G)
        packed = torch.stack([*out[0], *out[1], out[2]]).float()
        return packed.cpu()                       # device -> host read of the step's losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms / steps, wall / steps

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.load().tbg_reset_launch_count()
    ms_step, wall_step = timed(step_resident, args.steps)
    # host-issued launches (EMA, eager mode) + launches replayed from the captured graph
    launches = int(lib.load().tbg_launch_count()) + args.steps * ts.graph_launches(False, False)
    clocks = sampler.stop() if rank == 0 else None
    # the step time that counts is the slower of device time and host wall time per step
    ms_eff = max(ms_step, wall_step * 1e3)

    step_e2e()
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)
    ms_e2e = max(ms_e2e, wall_e2e * 1e3)
    h2d = real_h.numel() * 4 + words_h.numel() * 4 + labels_h.numel() * 4
    d2h = 7 * 4

    # ---- roofline pass: CUDA events around every tensor-core launch of two instrumented steps ----
    roof = None
    # every rank runs the two instrumented eager steps (they contain collectives); rank 0 records
    K.PROFILE = [] if rank == 0 else None
    ts.use_cuda_graph = False                     # events around every launch need eager launches
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    recs = K.PROFILE
    K.PROFILE = None
    if rank == 0:
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
            del xs
        peaks, which = _peaks()
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        key = ("conv_igemm", "modconv")
        n, secs, algo, execd = agg.get(key, [0, 1e-9, 0.0, 0.0])
        achieved = iso_algo / max(iso_secs, 1e-12) / 1e12
        roof = {
            "bound": "tensor", "kernel": "conv_igemm_kernel (modulated conv2d forward + input-gradient launches)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{which} bf16_tflops_sustained (MEASURED_PEAKS.json)",
            "method": "every modconv launch configuration of one step re-issued 20x back-to-back over rotating "
                      "inputs > L2, CUDA events on the launching stream; FLOPs = algorithmic (SURVEY 8d)",
            "executed_tflops": iso_exec / max(iso_secs, 1e-12) / 1e12,
            "launches_per_step": iso_launches / 2, "avg_launch_us": iso_secs / max(iso_launches, 1) * 1e6,
            "in_step_event_tflops": algo / secs / 1e12,
            "traffic": None,
            "traffic_note": "achieved aggregates 24 launch shapes, so no single per-launch DRAM figure applies; ncu --set "
                            "full of the top-layer launches (profiles/r01d_ncu_full_conv_kernels_summary.txt): 33.9 MB DRAM "
                            "read for 33.6 MB of activations (32x128x128ch, batch 32), 604 MB delivered L2->SM",
            "by_kernel_in_step": {f"{k[0]}:{k[1]}": {"launches_per_step": v[0] / 2, "ms_per_step": v[1] * 1e3 / 2,
                                                     "algorithmic_tflops": v[2] / max(v[1], 1e-12) / 1e12,
                                                     "executed_tflops": v[3] / max(v[1], 1e-12) / 1e12}
                                  for k, v in sorted(agg.items())},
        }

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, per = _cpu_oracle_rate(args.config, 4, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"1 timed plain training step (after 1 warm-up) at batch 4 of the config-{args.config} shape, "
                         f"oracle restatement of the reference cpu_only path, {threads} host threads"}
    gb = B * world
    line = {
        "metric": METRIC, "value": gb / (ms_eff * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_eff, "device_ms_per_step": ms_step,
        "host_wall_ms_per_step": wall_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": _config_dict(cfg, args, world),
        "clocks": clocks,
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    # NOTE: no dist.destroy_process_group() here — with NCCL collectives captured inside CUDA graphs it hung the
    # 8-GPU run at exit (round 1); the process group is torn down by interpreter exit.


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not replay the step from a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
