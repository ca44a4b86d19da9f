"""Profiling aid: K2 (img2pcl_index) alone at the benchmark shape, graph-timed; also the driver for
profiles/srcstalls.sh nearest_cells <tag> profiles/probe_k2.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
from keypointfusion_b200.utils import synth
dev = "cuda"
inp = synth.make_inputs(64, 128, 21, 128, seed=5)
c = {kk: v.to(dev) for kk, v in inp.items()}
pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
order = ops.spatial_order(pcl, c["center"], c["M"], c["cube"], c["cam"], 128, 32)
fn = lambda: ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True, order=order)
for _ in range(10): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
print("nearest_cells us %.2f" % (e0.elapsed_time(e1) * 1e3 / 50))
