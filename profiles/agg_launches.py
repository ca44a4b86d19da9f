"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the profiled window)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"<.*", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:64]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
