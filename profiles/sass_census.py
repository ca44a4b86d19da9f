"""Per-kernel census of the Blackwell-native SASS mnemonics in the shipped library (runs on the CPU box: cuobjdump + c++filt only).
   UTCHMMA = tcgen05.mma | LDTM / STTM = tcgen05.ld / st | UTCBAR = tcgen05.commit | UBLKCP = cp.async.bulk (1-D TMA) |
   UTMALDG / UTMASTG = cp.async.bulk.tensor | LDGSTS = cp.async | HMMA = legacy mma.sync (must be 0) | RED.*SYS = system-scope reduction
usage: python profiles/sass_census.py [lib.so] > profiles/sass_census_r2.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "keypointfusion_b200/libkpf_b200.so"
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "RED.SYS"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, fn = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        base = op.split(".")[0]
        if base in WATCH:
            counts[fn][base] += 1
        if base in ("RED", "REDG") and ".SYS" in op:   # system-scope reduction = the arrival counter of the fused exchange (peer memory)
            counts[fn]["RED.SYS"] += 1
names = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {so}: SASS mnemonic counts per kernel (sm_100a)")
print(f"{'kernel':58s} " + " ".join(f"{w:>8s}" for w in WATCH))
tot = collections.Counter()
for mangled, name in sorted(zip(counts, names), key=lambda kv: kv[1]):
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("kpf::", "")[:58]
    print(f"{short:58s} " + " ".join(f"{counts[mangled][w]:8d}" for w in WATCH))
    tot.update(counts[mangled])
print(f"{'TOTAL':58s} " + " ".join(f"{tot[w]:8d}" for w in WATCH))
