"""Summarise profiles/stalls.sh output: per launch, warp-cycles per issued instruction split by stall reason."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
d = defaultdict(dict)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        name = r[mi].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        d[(int(r[ii]), r[ki].split("(")[0][-28:])][name] = float(r[vi].replace(",", ""))
for k, v in sorted(d.items()):
    tot = sum(x for m, x in v.items() if "inst" not in m and "cycles" not in m)
    print(f"{k[0]:3d} {k[1]:28s} cycles={v.get('sm__cycles_elapsed.max', 0):9.0f} inst={v.get('smsp__inst_executed.sum', 0):11.0f} cyc/inst={tot:5.2f} | " +
          " ".join(f"{m}={x:.2f}" for m, x in sorted(v.items(), key=lambda t: -t[1]) if x >= 0.08 and "inst_exec" not in m and "cycles" not in m))
