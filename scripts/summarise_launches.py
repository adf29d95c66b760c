#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: summarise_launches.py launches.csv [skip_first_n_launches]
Prints kernel, launches, average microseconds and share of the summed GPU time
(per-launch times under ncu are cold-cache and serialised: shares, not absolutes)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'avg_us':>9s} {'share':>7s}")
    for k, a in agg.items():
        print(f"{k:70s} {a[0]:5d} {a[1] / a[0]:9.1f} {a[1] / tot * 100:6.1f}%")
    print(f"{'total':70s} {sum(a[0] for a in agg.values()):5d} {tot:9.1f}")


if __name__ == "__main__":
    main()
