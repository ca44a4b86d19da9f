"""Profiling aid: which stage of the path stops two half-batch chains (two branches of one CUDA graph, 74-SM budget each) from
overlapping?  Captures prefixes of the chain on one branch and on two, prints us per replay."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from keypointfusion_b200 import ops
from keypointfusion_b200 import custom_ops as cops
from keypointfusion_b200.model.model import KPFusion
from keypointfusion_b200.utils import synth
dev = "cuda"
net = KPFusion(joint_num=21); synth.fill_state_dict(net, 0); net = net.to(dev).eval()
J = 21
inp = synth.make_inputs(64, 128, 21, 128, seed=9)
full = {kk: inp[kk].to(dev) for kk in ("img", "img_feat", "img_feat_rgb", "img_offset", "center", "M", "cube", "cam")}
for kk in ("img_feat", "img_feat_rgb", "img_offset"): full[kk] = full[kk].bfloat16()
halves = [{k: v[:32].contiguous() for k, v in full.items()}, {k: v[32:].contiguous() for k, v in full.items()}]
STAGES = ["k1", "k4a", "order_k2", "repack", "pe1", "desa1", "tok1", "k5_1", "tokf1", "pe2", "desa2", "tok2", "k5_2", "tokf2"]

def chain(d, upto):
    st = {}
    blk = net.block1
    def run(name):
        k = (net.block1 if name.endswith("1") else net.block2).kc()
        b = net.block1 if name.endswith("1") else net.block2
        if name == "k1": st["pcl"] = ops.getpcl(d["img"], d["center"], d["cube"], d["M"], d["cam"], 1024, seed=0)[0]
        elif name == "k4a":
            st["juvd"] = ops.offset2joint_weight(d["img_offset"], d["img"], 0.8)
            st["jxyz"] = ops.uvd2xyz(st["juvd"], d["center"], d["M"], d["cube"], d["cam"], 128)
        elif name == "order_k2":
            st["order"] = ops.spatial_order(st["pcl"], d["center"], d["M"], d["cube"], d["cam"], 128, 32)
            st["close"], _, st["idx"] = ops.img2pcl_index(st["pcl"], d["img"], d["center"], d["M"], d["cube"], d["cam"], 128, 4, fs=32, want_i64=False, want_i32=True, order=st["order"])
        elif name == "repack": st["featT"] = ops.repack_features(d["img_feat"], d["img_feat_rgb"], d["img_offset"][:, 84:])
        elif name.startswith("pe"): st["e"] = ops.point_embed(st["featT"], st["idx"], st["close"], st["pcl"], st["jxyz"], k["pe_wmat"], k["pe_wvec"], 0.8, order=st["order"])
        elif name.startswith("desa"): st["part"], st["jf"] = ops.desa_fused(*st["e"], st["pcl"], st["jxyz"], k["ds_wmat"], k["ds_wvec"], b.FA.radius, 64)
        elif name.startswith("tokf"): st["jxyz"] = ops.token_stack(k["tok_final"], x=st["fj"], y=st["tok"], r3d=st["r3d"], want_tokens=False)[1]
        elif name.startswith("tok"): st["tok"], st["r3d"], _ = ops.token_stack(k["tok_init"], desa=st["part"], jf=st["jf"])
        elif name.startswith("k5"):
            _, st["fj"] = ops.spatial_aggregate_tc(d["img_feat_rgb"], st["r3d"], d["img"][:, :, ::4, ::4], d["center"], d["M"], d["cube"], d["cam"], k["wa_packed"],
                                                   b.atten_spatial.bias, b.weight_dis, b.fc_spatial2joint_feature.weight, b.fc_spatial2joint_feature.bias, prev=st.get("fj"))
    for n in STAGES[:upto]:
        run(n)

def timed_graph(nbranch, upto):
    side = torch.cuda.Stream()
    cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    with torch.no_grad(), ops.sm_budget(74):
        with torch.cuda.stream(cap):
            for h in halves[:nbranch]: chain(h, upto)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            main = torch.cuda.current_stream()
            if nbranch == 2:
                side.wait_stream(main)
                with torch.cuda.stream(side): chain(halves[1], upto)
            chain(halves[0], upto)
            if nbranch == 2: main.wait_stream(side)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100

for upto in range(1, len(STAGES) + 1):
    a, b = timed_graph(1, upto), timed_graph(2, upto)
    print(f"prefix ..{STAGES[upto - 1]:9s}: one branch {a:7.1f} us, two branches {b:7.1f} us  (ratio {b / a:.2f})", flush=True)
