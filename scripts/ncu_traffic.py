#!/usr/bin/env python
"""Extract per-launch DRAM traffic and duration of every kernel in an ncu report
(`ncu --set full ... -o rep`) into profiles/ncu_traffic.json (read by bench.py for roofline.traffic).

usage: ncu_traffic.py report.ncu-rep [out.json]"""
import csv
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_traffic.json"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("frn::", "")
        rd = float(d["dram__bytes_read.sum"]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * scale[unit["dram__bytes_write.sum"]]
        e = res.setdefault(name, {"launches": 0, "dram_bytes_per_launch": 0.0, "us": 0.0})
        e["launches"] += 1
        e["dram_bytes_per_launch"] += rd + wr
        tu = unit["gpu__time_duration.sum"]
        t = float(d["gpu__time_duration.sum"])
        e["us"] += t / 1000 if tu in ("ns", "nsecond") else (t * 1000 if tu in ("ms", "msecond") else t)
    for e in res.values():
        e["dram_bytes_per_launch"] /= e["launches"]
        e["us"] /= e["launches"]
    try:
        with open(out) as f:
            old = json.load(f)
    except Exception:  # noqa: BLE001
        old = {}
    old.update(res)
    with open(out, "w") as f:
        json.dump(old, f, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
