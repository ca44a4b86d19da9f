"""Top SASS lines by warp-stall samples from `ncu --page source --csv` (profiles/srcstalls.sh): python profiles/read_src.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
col = {k: i for i, k in enumerate(h)}
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    try: s = int(r[col["# Samples"]])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "lines", len(data))
agg = {k: 0 for k in st}
for s, r in data:
    for k in st:
        try: agg[k] += int(r[col[k]])
        except ValueError: pass
print("by reason:", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(tot, 1) > 0))
for s, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((int(r[col[k]] or 0), k[6:]) for k in st), reverse=True)[:2]
    print(f"{s * 100 / tot:5.1f}%  {r[col['Address']][-5:]}  {r[col['Source']][:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
