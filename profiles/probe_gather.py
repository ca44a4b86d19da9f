"""Profiling aid (not a test): how fast can an SM gather random 512-byte rows (the DESA row gather) out of L2 into shared memory?
kpf_gather_probe, 128 rows per CTA, L2-hot repetition; one CTA alone and one CTA on every SM.  python profiles/probe_gather.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
dev = "cuda"
R = 64 * 1056
table = torch.randint(-30000, 30000, (R, 256), device=dev, dtype=torch.int16)
names = ["TMA gather4", "bulk 512 B", "cp.async 128B req", "LDG.128 + STS", "cp.async row/warp"]
for n_rows in (64, 128, 256):
    for ctas in (1, 148):
        idx = torch.randint(0, R, (ctas, n_rows), device=dev, dtype=torch.int32)
        for mode in range(5):
            out = torch.zeros(ctas, dtype=torch.int64, device=dev)
            ops._call("kpf_gather_probe", ops._p(table), R, ops._p(idx), n_rows, ctas, mode, ops._p(out))
            torch.cuda.synchronize()
            c = out.float()
            print("rows %3d ctas %3d %-18s: %6.0f cycles (max %6.0f) -> %5.1f B/clk/SM" % (n_rows, ctas, names[mode], c.mean(), c.max(), n_rows * 512 / c.mean()))
