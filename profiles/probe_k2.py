"""Profiling aid: K2 (img2pcl_index) alone at the benchmark shape, for profiles/srcstalls.sh nearest_cells <tag> profiles/probe_k2.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
from keypointfusion_b200.utils import synth
dev = "cuda"
inp = synth.make_inputs(64, 128, 21, 128, seed=5)
c = {kk: v.to(dev) for kk, v in inp.items()}
pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
order = ops.spatial_order(pcl, c["center"], c["M"], c["cube"], c["cam"], 128, 32)
for _ in range(10):
    ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True, order=order)
torch.cuda.synchronize()
