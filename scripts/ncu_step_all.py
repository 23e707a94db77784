#!/usr/bin/env python
"""Per-kernel summary of an ncu capture of ONE eager plain training iteration (every kernel of the step, in-step inputs):
    ncu --metrics <METRICS below> --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_all.csv \
        python scripts/ncu_step_all.py run [cfg]
    python scripts/ncu_step_all.py summarize gpurun_out/step_all.csv > profiles/rNN_ncu_step_all_kernels.txt
`run` warms the iteration up outside the profiled range (cudaProfilerStart/Stop).  `summarize` aggregates by kernel name:
launches, total time, DRAM bytes read + written, achieved DRAM GB/s (bytes / time) against the measured HBM copy
bandwidth, tensor-pipe activity, achieved occupancy, registers.  Per-launch numbers are serialised and cold-ish (ncu
replays each launch once per pass): they characterise the kernels, they are not bench values."""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size")


def run(idx):
    import torch
    from textboxgan_b200.aster_inferer import AsterInferer
    from textboxgan_b200.config import baseline_config
    from textboxgan_b200.discriminator import Discriminator
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.optimizers import Adam, update_optimizer_params
    from textboxgan_b200.training_step import TrainingStep
    from oracle import train_step as OT          # synthetic batch only (test infrastructure, not on the measured path)
    cfg = baseline_config(idx)
    dev = "cuda:0"
    G = Generator(cfg, device=dev, seed=0); D = Discriminator(cfg, device=dev, seed=1)
    aster = AsterInferer(cfg, device=dev, synthetic_weights=True)
    go, do = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    ts = TrainingStep(G, D, aster, mk(go), mk(go), mk(do), 8, 16, torch.zeros((), device=dev), cfg)
    ts.use_cuda_graph = False
    ts.overlap_ocr = False                       # one stream: ncu serialises launches anyway
    g = torch.Generator().manual_seed(4444)
    real, words, labels = OT.synthetic_batch(cfg, cfg.batch_size_per_gpu, g)
    real, words, labels = real.to(dev), words.to(dev), labels.to(dev)
    zero = torch.zeros((), device=dev)
    step = lambda: ts.dist_train_step(real, zero, words, labels, False, False, 1e-4)
    for _ in range(3): step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


def summarize(path):
    peak = 6552.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        kid = row["ID"]
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("at::native::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"<unnamed>::", "", name)[:80]
        v = float(row["Metric Value"].replace(",", "") or 0)
        unit = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        per.setdefault(kid, {"name": name})[m] = v
    agg = collections.OrderedDict()
    for k in per.values():
        a = agg.setdefault(k["name"], {"n": 0, "us": 0.0, "bytes": 0.0, "tensor": 0.0, "warps": 0.0, "regs": 0, "max_us": 0.0})
        us = k.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1; a["us"] += us
        a["bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        a["tensor"] += k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
        a["warps"] += k.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) * us
        a["regs"] = max(a["regs"], int(k.get("launch__registers_per_thread", 0)))
        a["max_us"] = max(a["max_us"], us)
    tot = sum(a["us"] for a in agg.values())
    print(f"one eager plain training iteration, {len(per)} launches, {tot / 1e3:.3f} ms of serialised kernel time under ncu; "
          f"HBM peak = {peak:.0f} GB/s (MEASURED_PEAKS.json)")
    print(f"{'share':>6s} {'launches':>8s} {'total us':>9s} {'longest':>8s} {'DRAM MB':>9s} {'GB/s':>7s} {'of HBM':>6s} {'tensor%':>7s} {'occ%':>5s} {'regs':>4s}  kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = a["bytes"] / a["us"] / 1e3 if a["us"] else 0.0
        print(f"{a['us'] / tot * 100:5.1f}% {a['n']:8d} {a['us']:9.1f} {a['max_us']:8.1f} {a['bytes'] / 1e6:9.1f} {gbs:7.0f} {gbs / peak:6.2f} "
              f"{a['tensor'] / a['us'] if a['us'] else 0:7.1f} {a['warps'] / a['us'] if a['us'] else 0:5.1f} {a['regs']:4d}  {name}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
    else:
        summarize(sys.argv[2])
