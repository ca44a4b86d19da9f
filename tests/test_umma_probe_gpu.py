"""GPU: cycle micro-benchmarks of the tcgen05 building blocks (prints; asserts only sanity)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_probe(capsys):
    from keypointfusion_b200 import ops
    with capsys.disabled():
        print()
        for N, K in ((128, 128), (32, 128), (128, 32), (256, 128)):
            out = torch.zeros(8, dtype=torch.int64, device="cuda")
            ops._call("kpf_umma_probe", ops._p(out), N, K, 50)
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"[probe] N={N} K={K}: gemm+wait {o[0]} cyc | 8 gemms+wait {o[1]} | sync roundtrip {o[2]} | tmem 128-col ld {o[3]} | "
                  f"16 sts.128+fence {o[4]} | syncthreads {o[5]}")
            assert 0 < o[0] < 10_000_000


def test_token_stack_phase_clocks(path_params, capsys):
    """clock64 stamps of CTA 0 through one fused [fusion conv + init_TR] launch (profiling aid, prints only)."""
    import numpy as np
    from keypointfusion_b200 import ops
    p = path_params
    s = p["block1.FA.fusion.1.weight"] / torch.sqrt(p["block1.FA.fusion.1.running_var"] + 1e-5)
    Wfu = p["block1.FA.fusion.0.weight"].squeeze(-1) * s[:, None]
    bfu = (p["block1.FA.fusion.0.bias"] - p["block1.FA.fusion.1.running_mean"]) * s + p["block1.FA.fusion.1.bias"]
    pk = ops.pack_token_program(21, enc=(p, "block1.init_TR."), fusion=(Wfu, bfu)).to("cuda")
    B = 64
    part = torch.rand(B, 3, 21, 128, device="cuda")
    jf = torch.rand(B, 21, 128, device="cuda")
    dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
    for _ in range(3):
        torch.cuda.synchronize()          # an isolated launch: the first stamp then is the kernel's own start, not a predecessor's tail
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.token_stack(pk, desa=part, jf=jf, dbg=dbg)
        e1.record()
    torch.cuda.synchronize()
    with capsys.disabled():
        print(f"\n[token_stack] isolated launch (no successor overlapping on the idle SMs): {e0.elapsed_time(e1) * 1e3:.1f} us")
    t = dbg.cpu().numpy()[:32]
    n = int((t > 0).sum())
    d = np.diff(t[:n])
    with capsys.disabled():
        print("\n[token_stack] stamps:", n, "total cycles", int(t[n - 1] - t[0]))
        print("[token_stack] deltas: start->embed_done", int(d[0]), "| per layer [qkv, attention, o+proj+LN, ffn+LN]:")
        for l in range(4):
            print("   layer", l, [int(x) for x in d[1 + 4 * l: 5 + 4 * l]])
        print("   tail", [int(x) for x in d[17:]])
        f = dbg.cpu().numpy()[32:]
        nf = int((f > 0).sum())
        print("   fine stamps inside encoder layer 0 (deltas):", [int(x) for x in np.diff(f[:nf])])


def test_block_kernel_phase_clocks(path_params, capsys):
    """clock64 stamps of CTA 0 for point_embed / desa_fused / spatial_aggregate_tc at B = 64 (profiling aid, prints only)."""
    import numpy as np
    from keypointfusion_b200 import ops
    from keypointfusion_b200.model.model import Block_KPFusion
    from keypointfusion_b200.utils import synth
    B = 64
    inp = synth.make_inputs(B, 128, 21, 128, seed=5)
    c = {k: v.to("cuda") for k, v in inp.items()}
    pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
    close, _, idx = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True)
    joint = pcl[:, ::48][:, :21].contiguous() + 0.01
    blk = Block_KPFusion(21)
    blk.load_state_dict({k[len("block1."):]: v for k, v in path_params.items() if k.startswith("block1.")})
    blk = blk.to("cuda").eval()
    k = blk.kc()
    featT = ops.repack_features(c["img_feat"].bfloat16(), c["img_feat_rgb"].bfloat16(), c["img_offset"][:, 84:].bfloat16())

    def show(name, dbg, labels):
        t = dbg.cpu().numpy()
        n = int((t > 0).sum())
        d = np.diff(t[:n])
        with capsys.disabled():
            print(f"\n[{name}] total cycles {int(t[n - 1] - t[0])}:", [(labels[i] if i < len(labels) else f"d{i}", int(x)) for i, x in enumerate(d)])
    for _ in range(2):
        d1 = torch.zeros(64, dtype=torch.int64, device="cuda")
        e, acc, ms = ops.point_embed(featT, idx, close, pcl, joint, k["pe_wmat"], k["pe_wvec"], 0.8, dbg=d1)   # featT = (hi, None)
        d2 = torch.zeros(64, dtype=torch.int64, device="cuda")
        part, jf = ops.desa_fused(e, acc, ms, pcl, joint, k["ds_wmat"], k["ds_wvec"], blk.FA.radius, 64, dbg=d2)
        d3 = torch.zeros(64, dtype=torch.int64, device="cuda")
        ops.spatial_aggregate_tc(c["img_feat_rgb"].bfloat16(), joint, c["img"], c["center"], c["M"], c["cube"], c["cam"], k["wa_packed"],
                                 blk.atten_spatial.bias, blk.weight_dis, blk.fc_spatial2joint_feature.weight,
                                 blk.fc_spatial2joint_feature.bias, dbg=d3)
    torch.cuda.synchronize()
    show("point_embed (per tile: setup, gather, offsets, issue+softmax, wait, epilogue, agg-mma+store)", d1,
         ["setup", "gather", "offsets", "stage->issue", "softmax", "mma wait", "epilogue", "agg"] * 5)
    t2 = d2.cpu().numpy()
    with capsys.disabled():
        for nm, a, lab in (("jf role", t2[:8], ["stage", "agg", "mma issue", "mma + drain"]), ("ball-query role", t2[8:16], ["stage", "phase 1", "phase 2"])):
            na = int((a > 0).sum())
            print(f"\n[desa prep / {nm}] total cycles {int(a[na - 1] - a[0])}:", list(zip(lab, [int(x) for x in np.diff(a[:na])])))
        bb = t2[16:]
        nb = int((bb > 0).sum())
        print(f"[desa tiles] total cycles {int(bb[nb - 1] - bb[0])} (CTA 0), stamp deltas:", [int(x) for x in np.diff(bb[:nb])])
    show("spatial_aggregate_tc", d3, ["setup", "tile0", "t1 wait+prefetch", "t1 geometry", "t1 relu copy", "t1 gemmA", "t1 epiA", "t1 gemmB", "rest"])
