#!/usr/bin/env python
"""Per source line: stall samples and executed warp instructions of one kernel of an ncu report
(`--set full --import-source on`).  usage: ncu_hot_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, agg, seen_kernel = "", {}, 0
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            continue
        if len(r) < 8 or r[2] != "-":          # per-line rows carry '-' in the address column
            continue
        try:
            key = (fname, int(r[0]))
            smp, inst = int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += smp; a[1] += inst
    tot = sum(a[0] for a in agg.values()) or 1
    print(f"total samples {tot}, total warp instructions {sum(a[1] for a in agg.values())}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * a[0] / tot:5.1f}%  inst {a[1]:>9}  {f}:{ln}  {a[2][:110]}")


if __name__ == "__main__":
    main()
