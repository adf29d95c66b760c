#!/usr/bin/env python
"""One line per kernel launch of an ncu report (`ncu --set full ... -o rep`): duration, DRAM bytes,
SM / memory / tensor-pipe utilisation, occupancy, launch shape, instruction count.

usage: ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(c) if c in hdr else None for c in COLS]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLS)
        for r in rows[1:]:
            w.writerow([r[i] if i is not None else "" for i in idx])


if __name__ == "__main__":
    main()
