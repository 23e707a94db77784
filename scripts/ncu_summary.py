#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report (run here, no GPU needed):
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__cycles_elapsed.max", "sm_cyc"),
    ("sm__cycles_active.avg", "sm_cyc_active"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_mem_pipe%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l2_to_sm%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        short = h.split(".", 2)[-1] if h.startswith(("TPC.", "SM_", "GPC.", "FBP.")) else h
        col.setdefault(short, i)
        col.setdefault(h, i)
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        parts = [r[name_i].split("(")[0][:40]]
        vals = {}
        for k, short in KEYS:
            i = col.get(k)
            if i is None:
                continue
            vals[short] = (r[i], units[i])
            parts.append(f"{short}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
