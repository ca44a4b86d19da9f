"""Per-kernel DRAM bytes per launch from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [set(), 0.0])
for row in csv.DictReader(lines):
    if not row.get("Metric Name", "").startswith("dram__bytes"):
        continue
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "").replace("kpf::", "").strip()
    agg[name][0].add(row["ID"])
    agg[name][1] += float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
out = {k: round(v / len(ids), -3) for k, (ids, v) in agg.items() if not k.startswith("at::")}
out["_source"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (single pass; --set full hangs on the tcgen05 kernels), "
                  + sys.argv[1] + ", read + write bytes per launch at B=64, cold L2 (ncu flushes caches between kernels)")
print(json.dumps(out, indent=1))
