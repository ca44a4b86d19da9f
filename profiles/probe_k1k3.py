"""Profiling aid: K1 (getpcl) and K3 (gather_taps) alone at batch 512, for profiles/srcstalls.sh <kernel> <tag> profiles/probe_k1k3.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
from keypointfusion_b200.utils import synth
dev = "cuda"
B = int(os.environ.get("KPF_PROBE_B", "512"))
inp = synth.make_inputs(64, 128, 21, 128, seed=5)
c = {kk: v.to(dev).repeat(B // 64, *[1] * (v.dim() - 1)) for kk, v in inp.items()}
feat = c["img_feat"].bfloat16()
for _ in range(10):
    pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
    close, _, idx = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True)
    out = ops.gather_taps(feat, idx, close)
torch.cuda.synchronize()
