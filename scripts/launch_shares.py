#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
    python scripts/launch_shares.py gpurun_out/launches.csv [--slim out.csv] > profiles/rNN_launch_shares.txt
Per-launch times are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    slim = sys.argv[sys.argv.index("--slim") + 1] if "--slim" in sys.argv else None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    slim_rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        full = row["Kernel Name"]
        name = re.sub(r"<.*", "", full)
        name = re.sub(r"\(.*", "", name)[:64]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
        slim_rows.append((row["ID"], name, f"{v:.3f}"))
    tot = sum(v[1] for v in agg.values())
    print(f"# {n} launches, {tot / 1e3:.3f} ms summed kernel time (cold-cache, serialised under ncu)")
    print(f"# {'share':>6} {'count':>6} {'avg_us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:7.2f}% {v[0]:6d} {v[1] / v[0]:9.1f}  {k}")
    if slim:
        with open(slim, "w") as f:
            f.write("id,kernel,duration_us\n")
            for r in slim_rows:
                f.write(",".join(r) + "\n")


if __name__ == "__main__":
    main()
