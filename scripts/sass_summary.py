#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which Blackwell units a kernel uses
(B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG (TMA tensor load),
UBLKCP (1-D bulk copy), SYNCS (mbarrier), plus registers / shared memory from cuobjdump -res-usage.

  python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tf-fast-rnnt_b200", "lib", "libfast_rnnt_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.LG2",
             "SHFL", "HMMA", "DFMA", "BAR.SYNC", "UCGABAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (mn.count(".") and op.startswith(mn)):
                    counts[cur][mn] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(order)
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(SO, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UTMALDG = TMA tensor load,")
    print("# UBLKCP = cp.async.bulk (1-D, TMA engine), SYNCS = mbarrier ops, UCGABAR = cluster barrier")
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        short = re.sub(r"\(.*", "", names.get(fn, fn))
        short = re.sub(r"^void ", "", short)
        reg, smem = usage.get(fn, (None, None))
        tags = "  ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
        print(f"{short:<58} instr={c['_total']:<6} regs={reg} static_smem={smem}  {tags}")
        tot.update({k: v for k, v in c.items() if k != "_total"})
    print("# library totals: " + "  ".join(f"{mn}={tot[mn]}" for mn in MNEMONICS if tot[mn]))


if __name__ == "__main__":
    sys.exit(main())
