"""Profiling aid (not a test): do two of the path's kernels overlap when launched on two streams (different SMs)?  Times a 32-sample
token stack (32 CTAs) and a 74-CTA point stage alone and together."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
from keypointfusion_b200.model.model import KPFusion
from keypointfusion_b200.utils import synth
dev = "cuda"
B = 32
net = KPFusion(joint_num=21); synth.fill_state_dict(net, 0); net = net.to(dev).eval()
blk = net.block1; k = blk.kc()
inp = synth.make_inputs(B, 128, 21, 128, seed=5)
c = {kk: v.to(dev) for kk, v in inp.items()}
pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
close, _, idx = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True)
joint = pcl[:, ::48][:, :21].contiguous() + 0.01
featT = ops.repack_features(c["img_feat"].bfloat16(), c["img_feat_rgb"].bfloat16(), c["img_offset"][:, 84:].bfloat16())
part = torch.rand(B, 3, 21, 128, device=dev); jf = torch.rand(B, 21, 128, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def tok(): ops.token_stack(k["tok_init"], desa=part, jf=jf)
def pe():
    with ops.sm_budget(74): ops.point_embed(featT, idx, close, pcl, joint, k["pe_wmat"], k["pe_wvec"], 0.8)
def timed(fns, n=10):
    for f, s in fns:
        with torch.cuda.stream(s): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in (s1, s2): s.wait_event(e0)
    for _ in range(n):
        for f, s in fns:
            with torch.cuda.stream(s): f()
    for s in (s1, s2): torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("token stack alone  us", timed([(tok, s1)]))
print("point stage alone  us", timed([(pe, s2)]))
print("both, two streams  us", timed([(tok, s1), (pe, s2)]))
os.environ["X"] = "1"

# the same pair captured as two parallel branches of ONE CUDA graph
cap = torch.cuda.Stream()
cap.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=cap):
    main = torch.cuda.current_stream()
    s2.wait_stream(main)
    with torch.cuda.stream(s2):
        for _ in range(4): pe()
    for _ in range(4): tok()
    main.wait_stream(s2)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph with two branches (4 + 4 launches) us per pair", e0.elapsed_time(e1) * 1e3 / 20)
# a whole half-batch chain twice, as GraphedFusionPath(chains=2) captures it
from keypointfusion_b200.dataloader.loader import loader
from keypointfusion_b200.runtime import GraphedFusionPath
inp64 = synth.make_inputs(64, 128, 21, 128, seed=9)
ex = {kk: inp64[kk].to(dev) for kk in GraphedFusionPath.KEYS}
for kk in ("img_feat", "img_feat_rgb", "img_offset"): ex[kk] = ex[kk].bfloat16()
for ch in (1, 2):
    gp = GraphedFusionPath(net, loader(img_size=128), ex, chains=ch, bind=True)
    gp(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): gp()
    e1.record(); torch.cuda.synchronize()
    print(f"GraphedFusionPath chains={ch}: {e0.elapsed_time(e1) * 100:.1f} us per step")
