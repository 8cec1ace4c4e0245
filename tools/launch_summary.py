#!/usr/bin/env python
"""Share of the step per kernel from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv`).
The BVH build kernels (namespace gpubvh, scans) run once at upload and are left out.  usage: launch_summary.py x.csv [out.txt] [title]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, ns = r[4], float(r[14])
        if "gpubvh::" in name and "k_shadeTris" not in name or "cub::" in name or "DeviceScan" in name:
            continue
        short = re.sub(r"\(.*", "", name).replace("eleven::", "").strip()
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns / 1e3
    tot = sum(v[1] for v in agg.values())
    title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
    out = ["# ncu --metrics gpu__time_duration.sum launch list (%s; BVH build kernels filtered out): share per kernel" % title,
           "# kernel, launches, total us, share"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s, %d, %.1f, %.3f" % (k, v[0], v[1], v[1] / tot))
    out.append("# total %.1f us" % tot)
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
