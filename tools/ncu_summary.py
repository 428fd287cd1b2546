#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (reads it with `ncu -i ... --page raw --csv`): the metrics the
roofline in bench.py / DESIGN.md cites.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.csv"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
keys = [k for k in KEYS if k in h]
w.writerow(["id", "kernel"] + [f"{k} [{units[h.index(k)]}]" for k in keys])
for r in rows[2:]:
    w.writerow([r[h.index("ID")], r[h.index("Kernel Name")][:48]] + [r[h.index(k)] for k in keys])
