"""CPU: pin the oracle (oracle/kpf_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Tolerances: fp32 rel 1e-3 (north_star) with a small abs floor; indices exact
modulo provable distance ties."""
import numpy as np
import torch

from oracle import kpf_oracle as O
from keypointfusion_b200.utils import synth


def close(a, b, rtol=1e-3, atol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b)
    assert np.all(err <= atol + rtol * np.abs(b)), f"max abs err {err.max():.3e} (ref max {np.abs(b).max():.3e})"


def test_getpcl_mask_order_xyz(golden, golden_inputs, golden_meta):
    inp = golden_inputs
    seed = golden_meta["seed"]
    for b in range(golden_meta["B"]):
        pcl, pix = O.getpcl(inp["img"][b, 0].numpy(), inp["center"][b].numpy(), inp["cube"][b].numpy(),
                            inp["M"][b].numpy(), inp["cam"][b].numpy())
        assert pcl.shape[0] == golden["getpcl_counts"][b]          # mask count exact
        assert np.all(np.diff(pix) > 0)                             # row-major order
        if b == 0:
            close(pcl, golden["getpcl_full0"], rtol=1e-6, atol=1e-7)
        s, P = O.getpcl_sample(inp["img"][b, 0].numpy(), inp["center"][b].numpy(), inp["cube"][b].numpy(),
                               inp["M"][b].numpy(), inp["cam"][b].numpy(), ranks=synth.explicit_ranks(pcl.shape[0], 1024, seed + b))
        close(s, golden["pcl_sample"][b], rtol=1e-5, atol=1e-6)


def test_getpcl_planted_edge_cases(golden_inputs):
    inp = golden_inputs
    S = inp["img"].shape[-1]
    c = S // 2
    valid, _ = O.valid_mask(inp["img"][0, 0].numpy(), inp["center"][0].numpy(), inp["cube"][0].numpy())
    assert not valid[c, c]        # 1-5e-6 is inside the isclose band -> background
    assert valid[c, c + 1]        # 1-2e-5 is a valid point
    assert valid[c + 1, c] and valid[c + 4, c + 4]


def test_getpcl_small_and_empty(golden):
    img = synth.make_depth_crops(2, 32, 8)
    img[1] = 1.0
    c, M, cube, cam = synth.make_camera(2, 32, 8)
    for b in range(2):
        pcl, _ = O.getpcl(img[b, 0], c[b], cube[b], M[b], cam[b])
        ref = golden[f"getpcl_small{b}"]
        assert pcl.shape == ref.shape
        if ref.size:
            close(pcl, ref, rtol=1e-6, atol=1e-7)
    s, P = O.getpcl_sample(img[1, 0], c[1], cube[1], M[1], cam[1])
    assert P == 0 and not s.any()                       # loader.py:1176-1177
    s, P = O.getpcl_sample(img[0, 0], c[0], cube[0], M[0], cam[0], seed=3, b=0)
    assert 0 < P < 1024
    # multiset semantics for P < sample_num: every index floor(n/P) or floor(n/P)+1 times
    r = O.resample_ranks(P, 1024, 3, 0)
    cnt = np.bincount(r, minlength=P)
    assert cnt.min() == 1024 // P and cnt.max() <= 1024 // P + 1 and cnt.sum() == 1024


def test_feistel_is_permutation():
    for n in (1, 2, 3, 17, 1024, 1500, 4185):
        key = O.sample_key(5, n)
        vals = [O.feistel_perm(i, n, key) for i in range(n)]
        assert sorted(vals) == list(range(n))
        assert np.array_equal(O.feistel_perm_v(np.arange(n), n, key), np.array(vals))   # vectorised == scalar
    r = O.resample_ranks(5000, 1024, 1, 2)
    assert len(set(r.tolist())) == 1024 and r.max() < 5000


def test_uvd_xyz_transforms(golden, golden_inputs):
    i = golden_inputs
    xyz = O.uvd_nl2xyznl(golden["a5_uvd"], i["center"].numpy(), i["M"].numpy(), i["cube"].numpy(), i["cam"].numpy(), 128)
    close(xyz, golden["a5_xyz"], rtol=1e-4, atol=2e-6)
    back = O.xyz_nl2uvdnl(xyz, i["center"].numpy(), i["M"].numpy(), i["cube"].numpy(), i["cam"].numpy(), 128)
    close(back, golden["a5_uvd_back"], rtol=1e-4, atol=5e-6)
    close(back, golden["a5_uvd"], rtol=1e-4, atol=2e-5)   # round trip


def test_offset2joint_weight(golden, golden_inputs):
    i = golden_inputs
    close(O.offset2joint_weight(i["img_offset"], i["img"], 0.8), golden["a4_joint_uvd"], atol=2e-6)
    close(golden["a4_joint_uvd"], golden["a4_joint_uvd_gfm"], rtol=0, atol=0)  # the two reference copies agree
    close(O.offset2joint_weight(i["img_offset"], i["img"], torch.linspace(0.6, 1.0, 21)), golden["a4_joint_uvd_ktensor"], atol=2e-6)


def _check_topk(close_o, idx_o, d_o, close_r, idx_r, pcl, cells):
    """indices exact, except where the reference picked a cell whose distance ties (<= 2 ulp) the oracle's."""
    B, N, K = idx_o.shape
    bad = 0
    for b in range(B):
        diff = np.flatnonzero((np.sort(idx_o[b], 1) != np.sort(idx_r[b], 1)).any(1))
        for n in diff:
            d = pcl[b, n][None] - cells[b][idx_r[b, n]]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            assert np.allclose(np.sort(d2), np.sort(d_o[b, n]), rtol=3e-7, atol=0), (b, n, d2, d_o[b, n])
            bad += 1
    same = (idx_o == idx_r).all(-1)
    close(close_o[same], close_r[same], rtol=1e-3, atol=1e-6)
    return bad, same.mean()


def test_img2pcl_index(golden, golden_inputs):
    i = golden_inputs
    args = (i["center"].numpy(), i["M"].numpy(), i["cube"].numpy(), i["cam"].numpy(), 128)
    assert np.array_equal(O.nearest_down(i["img"], 32).numpy(), golden["img_down"])
    cells = O.cell_xyz(golden["img_down"], *args)
    c, idx, d = O.img2pcl_index(golden["pcl_sample"], golden["img_down"], *args, select_num=4)
    bad, frac_same = _check_topk(c, idx, d, golden["a6_closeness"], golden["a6_index"].astype(np.int64), golden["pcl_sample"], cells)
    assert frac_same > 0.98, frac_same
    c9, idx9, d9 = O.img2pcl_index(golden["pcl_sample"][:, :64], golden["img_down"], *args, select_num=9)
    _check_topk(c9, idx9, d9, golden["a6_closeness9"], golden["a6_index9"].astype(np.int64), golden["pcl_sample"][:, :64], cells)


def test_pcl_joint2offset(golden):
    out = O.pcl_joint2offset(torch.from_numpy(golden["joint_xyz0"]), torch.from_numpy(golden["pcl_sample"]), 0.8)
    close(out[:, :128], golden["a7_pcl_offset"], atol=2e-6)
    close(golden["a7_pcl_offset"], golden["a7_pcl_offset_gfm"], rtol=0, atol=0)


def test_gathers(golden, golden_inputs):
    i = golden_inputs
    idx = torch.from_numpy(golden["a6_index"].astype(np.int64))
    cl = torch.from_numpy(golden["a6_closeness"])
    close(O.gather_taps(i["img_feat"], idx, cl)[:, :64], golden["a8_pcl_feat"], atol=2e-6)
    close(O.gather_taps(i["img_feat_rgb"], idx, cl)[:, :64], golden["a8_pcl_feat_rgb"], atol=2e-6)
    close(O.gather_taps(i["img_offset"][:, 84:], idx, cl)[:, :64], golden["a8_pcl_weight"], atol=2e-6)


def test_heatmap_and_gam(golden, golden_inputs):
    i = golden_inputs
    j3 = torch.from_numpy(golden["a10_joint"])
    close(O.joint2heatmap(j3[:, :, :2], 0.8, 32, sigma=1)[:1], golden["a10_hm_s1"], atol=1e-7)
    close(O.joint2heatmap(j3, 0.8, 32)[:1], golden["a10_hm_default"], atol=1e-7)
    gam = O.img2anchor_dis(j3, torch.from_numpy(golden["img_down"]), i["center"].numpy(), i["M"].numpy(), i["cube"].numpy(),
                           i["cam"].numpy(), 128)
    close(gam[:1], golden["a11_gam"], rtol=1e-4, atol=1e-7)


def test_joint2offset(golden, golden_inputs):
    i = golden_inputs
    j3 = torch.from_numpy(golden["a10_joint"])
    close(O.joint2offset(j3, i["img"], 0.8, 32)[:1], golden["a16_joint2offset_gfm"], atol=2e-6)
    close(O.joint2offset(j3, i["img"], 0.8, 32, eps=0.0)[:1], golden["a16_joint2offset_model"], atol=2e-6)
    close(golden["a16_joint2feature"], golden["a16_joint2offset_gfm"], rtol=0, atol=0)  # dispatcher == joint2offset
    pix = torch.cat([O.joint2offset(j3, i["img"], 0.8, 32), torch.from_numpy(golden["a16_feature2joint_in_w"])], 1)
    close(O.offset2joint_weight(pix, i["img"], 0.8), golden["a16_feature2joint"], atol=5e-6)


def _dec_params(golden_meta):
    sd = {k: torch.zeros(s) for k, s in golden_meta["updatedDecoder_keys"].items()}
    return synth.fill_state_dict(sd, golden_meta["seed"])


def test_updated_decoder(golden, golden_meta):
    p = _dec_params(golden_meta)
    out = O.updated_decoder(p, "", torch.from_numpy(golden["a13_anchor"]), torch.from_numpy(golden["a13_key"]))
    close(out, golden["a13_out"], atol=5e-6)


def test_fusion_layers(golden, golden_meta):
    for name, fn, cls in (("rgbd", O.rgbd_fusion, "RGBDFusion"), ("ac", O.ac_fusion, "ACFusion")):
        p = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta[f"{cls}_keys"].items()}, golden_meta["seed"])
        (ro, do), mg = fn(p, torch.from_numpy(golden[f"a14_{name}_rgb"]), torch.from_numpy(golden[f"a14_{name}_depth"]))
        close(ro, golden[f"a14_{name}_rgb_out"], atol=2e-6)
        close(do, golden[f"a14_{name}_depth_out"], atol=2e-6)
        close(mg, golden[f"a14_{name}_merge"], atol=2e-6)
    p = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["FSP_keys"].items()}, golden_meta["seed"])
    close(O.fsp(p, torch.from_numpy(golden["a14_rgbd_rgb"]), torch.from_numpy(golden["a14_rgbd_depth"])), golden["a15_fsp_out"], atol=2e-6)


def test_whole_path(golden, golden_inputs, path_params):
    """a9/a12/a17 + DESA + BERT encoders: the oracle's fusion path vs the reference's two blocks, fed the
    reference's own pcl/index so differences are float-only.  Final joints: <= 0.05 mm mean (north_star)."""
    i = golden_inputs
    p = path_params
    pcl = torch.from_numpy(golden["pcl_sample"])
    idx = torch.from_numpy(golden["a6_index"].astype(np.int64))
    cl = torch.from_numpy(golden["a6_closeness"])
    jx = torch.from_numpy(golden["joint_xyz0"])
    img_down = torch.from_numpy(golden["img_down"])
    prev = None
    for blk in (1, 2):
        (r3d, r2d, prev, sw, _), t = O.block_kpfusion(p, f"block{blk}.", i["img_feat"], i["img_feat_rgb"], pcl, jx, cl, idx,
                                                      i["img_offset"], prev, img_down, i["center"].numpy(), i["M"].numpy(),
                                                      i["cube"].numpy(), i["cam"].numpy(), 128)
        close(t["joint_feat_desa"], golden[f"b{blk}_desa"], atol=2e-5)
        close(t["tok_init"], golden[f"b{blk}_tok_init"], atol=2e-5)
        close(t["cross"], golden[f"b{blk}_cross"].transpose(0, 2, 1), atol=2e-5)
        close(prev, golden[f"b{blk}_img_feat_j"], atol=2e-5)
        close(sw[:1], golden[f"b{blk}_sw0"], atol=2e-6)
        for name, v in (("r3d", r3d), ("r2d", r2d)):
            ref = golden[f"b{blk}_{name}"]
            mm = np.linalg.norm((v.numpy() - ref) * 125.0, axis=-1).mean()   # cube/2 = 125 mm
            assert mm <= 0.05, (blk, name, mm)
            close(v, ref, atol=2e-5)
        jx = r2d
    # the oracle's own front end reproduces the reference's inputs to the blocks
    res, sws, ex = O.fusion_path(p, i["img"], pcl, i["img_offset"], i["img_feat"], i["img_feat_rgb"], i["center"].numpy(),
                                 i["M"].numpy(), i["cube"].numpy(), i["cam"].numpy())
    mm = np.linalg.norm((res[3].numpy() - golden["b2_r2d"]) * 125.0, axis=-1).mean()
    assert mm <= 0.05, mm


def test_eval_tail_rigid_align(golden):
    """oracle rigid_align vs the reference's GFM.rigid_align (incl. the reflection branch)."""
    for i in range(golden["f4_A"].shape[0]):
        a2 = O.rigid_align(golden["f4_A"][i].astype(np.float64), golden["f4_B"][i].astype(np.float64))
        close(a2, golden["f4_aligned"][i], rtol=1e-5, atol=1e-6)   # the reference ran on float32 joints
