"""GPU parity tests: every kernel behind the C ABI vs the CPU oracle on the same seeded inputs, and vs the
reference-generated golden fixtures.  Bars (north_star): indices / masks / order bit-exact; fp32 features
<= 1e-3 rel; bf16 <= 1e-2 rel."""
import numpy as np
import pytest
import torch

from keypointfusion_b200.utils import synth
from oracle import kpf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from keypointfusion_b200 import ops as _ops
    return _ops


def close(a, b, rtol=1e-3, atol=1e-5):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a.astype(np.float64) - b.astype(np.float64))
    assert np.all(err <= atol + rtol * np.abs(b)), f"max abs err {err.max():.3e} (ref max {np.abs(b).max():.3e})"


def cu(x):
    return (x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))).to(DEV)


def geo(inp):
    return [inp[k].numpy() for k in ("center", "M", "cube", "cam")]


# ------------------------------------------------------------------------------------------------ a1-a3
@pytest.mark.parametrize("S", [32, 64, 128, 256])
def test_backproject_all_exact(ops, S):
    B = 3
    img = synth.make_depth_crops(B, S, seed=S)
    img[B - 1] = 1.0  # empty sample
    center, M, cube, cam = synth.make_camera(B, S, seed=S)
    xyz, pix, cnt = ops.backproject_all(cu(img), cu(center), cu(cube), cu(M), cu(cam))
    xyz, pix, cnt = xyz.cpu().numpy(), pix.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        ref, rpix = O.getpcl(img[b, 0], center[b], cube[b], M[b], cam[b])
        P = ref.shape[0]
        assert cnt[b] == P                                     # valid mask: exact count
        assert np.array_equal(pix[b, :P], rpix)                # pixel -> point order: bit-exact
        assert np.all(pix[b, P:] == -1)
        assert np.array_equal(xyz[b, :P], ref.astype(np.float32)), np.abs(xyz[b, :P] - ref).max()  # fp64 math, fp32 rounding: exact
    assert cnt[B - 1] == 0


def test_getpcl_sample_vs_oracle_and_golden(ops, golden, golden_inputs, golden_meta):
    inp = golden_inputs
    B, seed = golden_meta["B"], golden_meta["seed"]
    ranks = np.stack([synth.explicit_ranks(int(golden["getpcl_counts"][b]), 1024, seed + b) for b in range(B)])
    pcl, cnt = ops.getpcl(cu(inp["img"]), cu(inp["center"]), cu(inp["cube"]), cu(inp["M"]), cu(inp["cam"]), ranks=cu(ranks))
    assert np.array_equal(cnt.cpu().numpy(), golden["getpcl_counts"])
    close(pcl, golden["pcl_sample"], rtol=1e-5, atol=1e-6)      # vs the reference's numpy getpcl
    # built-in counter-based selection == oracle's, bit for bit
    pcl2, _ = ops.getpcl(cu(inp["img"]), cu(inp["center"]), cu(inp["cube"]), cu(inp["M"]), cu(inp["cam"]), seed=11)
    for b in range(B):
        ref, P = O.getpcl_sample(inp["img"][b, 0].numpy(), inp["center"][b].numpy(), inp["cube"][b].numpy(), inp["M"][b].numpy(),
                                 inp["cam"][b].numpy(), seed=11, b=b)
        assert np.array_equal(pcl2[b].cpu().numpy(), ref)


def test_getpcl_sparse_empty_clamp(ops):
    img = synth.make_depth_crops(3, 32, 8)
    img[1] = 1.0
    img[2] = 1.0
    img[2, 0, 5, 7] = 0.25  # a single valid pixel
    center, M, cube, cam = synth.make_camera(3, 32, 8)
    pcl, cnt = ops.getpcl(cu(img), cu(center), cu(cube), cu(M), cu(cam), seed=3, clamp=True)
    pcl, cnt = pcl.cpu().numpy(), cnt.cpu().numpy()
    assert 0 < cnt[0] < 1024 and cnt[1] == 0 and cnt[2] == 1
    assert not pcl[1].any()
    for b in (0, 2):
        ref, P = O.getpcl_sample(img[b, 0], center[b], cube[b], M[b], cam[b], seed=3, b=b, clamp=True)
        assert np.array_equal(pcl[b], ref)
    assert np.all(pcl[2] == pcl[2][0])


# ------------------------------------------------------------------------------------------------ a5
def test_uvd_xyz_exact(ops, golden, golden_inputs):
    inp = golden_inputs
    g = geo(inp)
    uvd = golden["a5_uvd"]
    xyz = ops.uvd2xyz(cu(uvd), *[cu(x) for x in g], 128)
    ref = O.uvd_nl2xyznl(uvd, *g, 128)
    assert np.array_equal(xyz.cpu().numpy(), ref)               # same fp32 op order -> bit-exact vs oracle
    close(xyz, golden["a5_xyz"], rtol=1e-4, atol=2e-6)          # vs reference
    back = ops.xyz2uvd(xyz, *[cu(x) for x in g], 128)
    assert np.array_equal(back.cpu().numpy(), O.xyz_nl2uvdnl(ref, *g, 128))
    close(back, golden["a5_uvd_back"], rtol=1e-4, atol=5e-6)


# ------------------------------------------------------------------------------------------------ a6
@pytest.mark.parametrize("K", [4, 9])
def test_img2pcl_index_exact(ops, golden, golden_inputs, K):
    inp = golden_inputs
    g = geo(inp)
    pcl = golden["pcl_sample"]
    # full-resolution crop passed with strides (fused nearest down-sample) and the explicit down-sampled map agree
    c1, i64, i32 = ops.img2pcl_index(cu(pcl), cu(inp["img"]), *[cu(x) for x in g], 128, select_num=K, fs=32, want_i32=True)
    c2, j64, _ = ops.img2pcl_index(cu(pcl), cu(golden["img_down"]), *[cu(x) for x in g], 128, select_num=K)
    assert torch.equal(i64, j64) and torch.equal(c1, c2) and torch.equal(i64.int(), i32)
    rc, ri, rd = O.img2pcl_index(pcl, golden["img_down"], *g, 128, select_num=K)
    assert np.array_equal(i64.cpu().numpy(), ri)                # sampled-cell indices: bit-exact
    assert np.array_equal(c1.cpu().numpy(), rc)                 # weights: same op order -> bit-exact
    if K == 4:  # vs reference (ties aside)
        same = (ri == golden["a6_index"]).all(-1)
        assert same.mean() > 0.98
        close(c1.cpu().numpy()[same], golden["a6_closeness"][same], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("S,J", [(64, 21), (96, 42), (192, 21), (256, 42)])
def test_kernel_sweep_backproject_index_gather(ops, S, J):
    """BASELINE config 4: K1 + K2 + K3 over crop sizes 64-256 and 21-42 joints (fs = S/4)."""
    B, fs = 2, S // 4
    inp = synth.make_inputs(B, S, J, 128, seed=S + J)
    g = geo(inp)
    pcl, cnt = ops.getpcl(cu(inp["img"]), cu(inp["center"]), cu(inp["cube"]), cu(inp["M"]), cu(inp["cam"]), seed=5)
    for b in range(B):
        ref, P = O.getpcl_sample(inp["img"][b, 0].numpy(), g[0][b], g[2][b], g[1][b], g[3][b], seed=5, b=b)
        assert int(cnt[b]) == P and np.array_equal(pcl[b].cpu().numpy(), ref)
    c, i64, _ = ops.img2pcl_index(pcl, cu(inp["img"]), *[cu(x) for x in g], S, select_num=4, fs=fs)
    rc, ri, _ = O.img2pcl_index(pcl.cpu().numpy(), O.nearest_down(inp["img"], fs).numpy(), *g, S, select_num=4)
    assert np.array_equal(i64.cpu().numpy(), ri) and np.array_equal(c.cpu().numpy(), rc)
    for name, ch in (("img_feat", slice(None)), ("img_offset", slice(4 * J, None))):
        out = ops.gather_taps(cu(inp[name])[:, ch], i64, c)
        close(out, O.gather_taps(inp[name][:, ch], torch.from_numpy(ri), torch.from_numpy(rc)), atol=2e-6)


# ------------------------------------------------------------------------------------------------ a4, a7, a8, a10, a11, a16
def test_offset2joint_weight(ops, golden, golden_inputs):
    inp = golden_inputs
    out = ops.offset2joint_weight(cu(inp["img_offset"]), cu(inp["img"]), 0.8)
    close(out, O.offset2joint_weight(inp["img_offset"], inp["img"], 0.8), atol=2e-6)
    close(out, golden["a4_joint_uvd"], atol=2e-6)
    ks = torch.linspace(0.6, 1.0, 21)
    close(ops.offset2joint_weight(cu(inp["img_offset"]), cu(inp["img"]), ks), golden["a4_joint_uvd_ktensor"], atol=2e-6)
    ob = inp["img_offset"].bfloat16()
    close(ops.offset2joint_weight(cu(ob), cu(inp["img"]), 0.8), O.offset2joint_weight(ob.float(), inp["img"], 0.8), atol=2e-6)
    # all-background sample: softmax over uniformly masked weights (mean of cell coords)
    img = torch.ones(1, 1, 128, 128)
    close(ops.offset2joint_weight(cu(inp["img_offset"][:1]), cu(img), 0.8), O.offset2joint_weight(inp["img_offset"][:1], img, 0.8),
          atol=2e-6)


def test_pcl_joint2offset(ops, golden):
    out = ops.pcl_joint2offset(cu(golden["joint_xyz0"]), cu(golden["pcl_sample"]), 0.8)
    close(out, O.pcl_joint2offset(torch.from_numpy(golden["joint_xyz0"]), torch.from_numpy(golden["pcl_sample"]), 0.8), atol=2e-6)
    close(out[:, :128], golden["a7_pcl_offset"], atol=2e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_taps(ops, golden, golden_inputs, dtype):
    inp = golden_inputs
    idx = torch.from_numpy(golden["a6_index"].astype(np.int64))
    cl = torch.from_numpy(golden["a6_closeness"])
    tol = dict(rtol=1e-3, atol=2e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    for name, key, ch in (("img_feat", "a8_pcl_feat", slice(None)), ("img_feat_rgb", "a8_pcl_feat_rgb", slice(None)),
                          ("img_offset", "a8_pcl_weight", slice(84, None))):
        f = inp[name].to(dtype)
        out = ops.gather_taps(cu(f)[:, ch], cu(idx), cu(cl))
        close(out, O.gather_taps(f.float()[:, ch], idx, cl), **tol)
        close(out[:, :64], golden[key], **tol)
        out32 = ops.gather_taps(cu(f)[:, ch], cu(idx.int()), cu(cl))   # int32 indices: same result
        assert torch.equal(out, out32)


def test_heatmap_gam_joint2offset(ops, golden, golden_inputs):
    inp = golden_inputs
    g = geo(inp)
    j3 = torch.from_numpy(golden["a10_joint"])
    close(ops.joint2heatmap(cu(j3[:, :, :2]), 0.8, 32, sigma=1)[:1], golden["a10_hm_s1"], atol=1e-6)
    close(ops.joint2heatmap(cu(j3), 0.8, 32)[:1], golden["a10_hm_default"], atol=1e-6)
    gam = ops.img2anchor_dis(cu(j3), cu(inp["img"]), *[cu(x) for x in g], 128, fs=32)
    close(gam[:1], golden["a11_gam"], rtol=1e-3, atol=1e-6)
    close(gam, O.img2anchor_dis(j3, torch.from_numpy(golden["img_down"]), *g, 128), rtol=1e-3, atol=1e-6)
    close(ops.joint2offset(cu(j3), cu(inp["img"]), 0.8, 32)[:1], golden["a16_joint2offset_gfm"], atol=2e-6)
    close(ops.joint2offset(cu(j3), cu(inp["img"]), 0.8, 32, eps=0.0)[:1], golden["a16_joint2offset_model"], atol=2e-6)


# ------------------------------------------------------------------------------------------------ a12, a13
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_spatial_aggregate(ops, golden, golden_inputs, path_params, dtype):
    inp = golden_inputs
    g = geo(inp)
    p = path_params
    j3 = torch.from_numpy(golden["a10_joint"])
    f = inp["img_feat_rgb"].to(dtype)
    prev = torch.from_numpy(np.random.RandomState(0).standard_normal((2, 21, 128)).astype(np.float32))
    hm = O.joint2heatmap(j3[:, :, :2], 0.8, 32, sigma=1)
    gam = O.img2anchor_dis(j3, torch.from_numpy(golden["img_down"]), *g, 128)
    tol = dict(rtol=1e-3, atol=2e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=2e-3)
    for pv in (None, prev):
        rsw, rfj = O.spatial_aggregate(p, "block1.", f.float(), hm, gam, pv)
        sw, fj, hmo, gamo = ops.spatial_aggregate(cu(f), cu(j3), cu(inp["img"]), *[cu(x) for x in g], cu(p["block1.atten_spatial.weight"]),
                                                  cu(p["block1.atten_spatial.bias"]), cu(p["block1.weight_dis"]),
                                                  cu(p["block1.fc_spatial2joint_feature.weight"]),
                                                  cu(p["block1.fc_spatial2joint_feature.bias"]), prev=None if pv is None else cu(pv),
                                                  want_maps=True)
        close(hmo, hm, atol=1e-6)
        close(gamo, gam, rtol=1e-3, atol=1e-6)
        close(sw, rsw, **tol)
        close(fj, rfj, **tol)


@pytest.mark.parametrize("maps", ["bf16", "fp32"])
def test_spatial_aggregate_tensor_core(ops, golden, golden_inputs, path_params, maps):
    """tcgen05 version, split precision: fp32-class vs the fp32 oracle on the same features -- bf16 maps (one exact plane) and fp32 maps
    (two bf16 planes from kpf_split_planes)."""
    inp = golden_inputs
    g = geo(inp)
    p = path_params
    j3 = torch.from_numpy(golden["a10_joint"])
    f = inp["img_feat_rgb"].bfloat16().float() if maps == "bf16" else inp["img_feat_rgb"]
    prev = torch.from_numpy(np.random.RandomState(0).standard_normal((2, 21, 128)).astype(np.float32))
    hm = O.joint2heatmap(j3[:, :, :2], 0.8, 32, sigma=1)
    gam = O.img2anchor_dis(j3, torch.from_numpy(golden["img_down"]), *g, 128)
    wa = ops.pack_spatial_wa(p["block1.atten_spatial.weight"], 21)
    if maps == "bf16":
        planes = cu(f).bfloat16()
    else:
        planes = ops.split_map(cu(f))
        assert float((planes[0].float() + planes[1].float() - cu(f)).abs().max()) <= 2.0 ** -15 * float(f.abs().max())
    for pv in (None, prev):
        rsw, rfj = O.spatial_aggregate(p, "block1.", f, hm, gam, pv)
        sw, fj = ops.spatial_aggregate_tc(planes, cu(j3), cu(inp["img"]), *[cu(x) for x in g], cu(wa), cu(p["block1.atten_spatial.bias"]),
                                          cu(p["block1.weight_dis"]), cu(p["block1.fc_spatial2joint_feature.weight"]),
                                          cu(p["block1.fc_spatial2joint_feature.bias"]), prev=None if pv is None else cu(pv))
        for name, a, b in (("sw", sw, rsw), ("fj", fj, rfj)):
            a = a.float().cpu()
            rel, worst = float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())
            print(f"[K5 tc {maps}] {name}: rms rel {rel:.2e} worst {worst:.2e}")
            assert rel < 3e-5 and worst < 3e-4, (name, rel, worst)


def test_cross_decoder_layer(ops, golden, golden_meta):
    sd = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["updatedDecoder_keys"].items()}, golden_meta["seed"])
    wp = ops.pack_decoder_layer(sd, "decoder.3.", 21, 128)
    out = ops.cross_decoder_layer(cu(golden["a13_anchor"]), cu(golden["a13_key"]), cu(wp), heads=4, ffn=128)
    close(out, golden["a13_out"], atol=2e-5)                     # vs the reference's updatedDecoder
    close(out, O.updated_decoder(sd, "", torch.from_numpy(golden["a13_anchor"]), torch.from_numpy(golden["a13_key"])), atol=2e-5)


# ------------------------------------------------------------------------------------------------ a14, a15
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fusion_layers(ops, golden, golden_meta, dtype):
    tol = dict(rtol=1e-3, atol=2e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    seed = golden_meta["seed"]
    r, d = torch.from_numpy(golden["a14_rgbd_rgb"]).to(dtype), torch.from_numpy(golden["a14_rgbd_depth"]).to(dtype)
    p = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["RGBDFusion_keys"].items()}, seed)
    gw = torch.cat([p["gate_rgb.weight"].reshape(1, -1), p["gate_depth.weight"].reshape(1, -1)], 0)
    gb = torch.cat([p["gate_rgb.bias"], p["gate_depth.bias"]])
    ro, do, mg, am = ops.rgbd_fusion(cu(r), cu(d), cu(gw), cu(gb), want_attn_mean=True)
    (rro, rdo), rmg = O.rgbd_fusion(p, r.float(), d.float())
    close(ro, rro, **tol), close(do, rdo, **tol), close(mg, rmg, **tol)
    assert abs(float(am.sum()) - 1.0) < 1e-4
    if dtype == torch.float32:
        close(ro, golden["a14_rgbd_rgb_out"], **tol), close(mg, golden["a14_rgbd_merge"], **tol)
    r, d = torch.from_numpy(golden["a14_ac_rgb"]).to(dtype), torch.from_numpy(golden["a14_ac_depth"]).to(dtype)
    p = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["ACFusion_keys"].items()}, seed)
    ro, do, mg = ops.ac_fusion(cu(r), cu(d), cu(p["cam_rgb.weight"]), cu(p["cam_rgb.bias"]), cu(p["cam_depth.weight"]), cu(p["cam_depth.bias"]))
    (rro, rdo), rmg = O.ac_fusion(p, r.float(), d.float())
    close(ro, rro, **tol), close(do, rdo, **tol), close(mg, rmg, **tol)
    if dtype == torch.float32:
        close(do, golden["a14_ac_depth_out"], **tol)
    p = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["FSP_keys"].items()}, seed)
    gdd, mn = torch.from_numpy(golden["a14_rgbd_rgb"]).to(dtype), torch.from_numpy(golden["a14_rgbd_depth"]).to(dtype)
    out = ops.fsp(cu(gdd), cu(mn), cu(p["filter.fc.0.weight"]), cu(p["filter.fc.0.bias"]), cu(p["filter.fc.2.weight"]), cu(p["filter.fc.2.bias"]))
    close(out, O.fsp(p, gdd.float(), mn.float()), **tol)
    if dtype == torch.float32:
        close(out, golden["a15_fsp_out"], **tol)


# ------------------------------------------------------------------------------------------------ 8f-4 evaluation tail
def test_eval_errors(ops, golden):
    A, Bt = golden["f4_A"], golden["f4_B"]
    n = A.shape[0]
    cube = np.full((n, 3), 250.0, np.float32)
    center = np.zeros((n, 3), np.float32)
    err, pa = ops.eval_errors(cu(A), cu(Bt), cu(cube))
    close(err, O.xyz2error(A, Bt, center, cube), rtol=1e-5, atol=1e-4)
    ref_pa = O.xyz2error(golden["f4_aligned"], Bt, center, cube)       # the reference's own aligned joints
    close(pa, ref_pa, rtol=1e-4, atol=1e-3)
    assert float(pa[1].mean()) > 1.0                                    # the reflected sample cannot be aligned by a rotation


@pytest.mark.parametrize("N", [37, 300, 1024, 1500])
def test_spatial_order_is_a_permutation(ops, golden_inputs, N):
    """Scheduling aid (no reference counterpart): for any N the order must be a permutation of the point ids; both sort paths
    (one key per thread with warp shuffles for N <= 1024, shared-memory bitonic above) are exercised."""
    inp = golden_inputs
    B = inp["img"].shape[0]
    pcl = torch.rand(B, N, 3, generator=torch.Generator().manual_seed(N)) * 1.6 - 0.8
    order = ops.spatial_order(cu(pcl), cu(inp["center"]), cu(inp["M"]), cu(inp["cube"]), cu(inp["cam"]), 128, 32)
    assert order.dtype == torch.int32 and tuple(order.shape) == (B, N)
    assert torch.equal(torch.sort(order.long().cpu(), dim=1)[0], torch.arange(N).expand(B, -1))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fs,C,K", [(15, 12, 9), (32, 37, 4), (48, 128, 4), (80, 12, 9)])
def test_gather_taps_shapes_and_fallback(ops, dtype, fs, C, K):
    """K3 beyond the benchmark shape: odd map sizes (scalar loader), channel counts that are no multiple of the 16-byte chunk, 9 taps,
    a channel slice of a wider map, and an 80 x 80 map whose slab does not fit shared memory (two-kernel workspace path)."""
    g = torch.Generator().manual_seed(fs * 131 + C)
    B, N, HW = 2, 300, fs * fs
    wide = torch.randn(B, C + 5, fs, fs, generator=g).to(dtype)
    feat = wide[:, 3:3 + C]
    idx = torch.randint(0, HW, (B, N, K), generator=g)
    cl = torch.rand(B, N, K, generator=g)
    ref = torch.einsum("bnkc,bnk->bnc", feat.float().reshape(B, C, HW).transpose(1, 2)[torch.arange(B)[:, None, None], idx], cl)
    out = ops.gather_taps(cu(wide)[:, 3:3 + C], cu(idx), cu(cl))
    assert out.dtype == dtype and out.shape == (B, N, C)
    tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=2e-2)
    close(out.float(), ref, **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,h,w", [(64, 32, 32), (128, 16, 16), (256, 8, 8), (512, 4, 4), (96, 24, 24), (40, 7, 7), (8, 2, 2)])
def test_rgbd_fusion_stage_shapes(ops, dtype, C, h, w):
    """K7 at the four ResNet-18 stage shapes of a 128 crop (model/resnet.py:439-442: the chunked kernel with 32 / 8 / 2 / 2 chunk
    columns per CTA), a 24 x 24 map (partial last CTA), and two maps whose HW is no multiple of the 16-byte chunk (generic kernel)."""
    g = torch.Generator().manual_seed(C + h)
    B = 3
    r, d = torch.randn(B, C, h, w, generator=g).to(dtype), torch.randn(B, C, h, w, generator=g).to(dtype)
    p = {"gate_rgb.weight": torch.randn(1, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5, "gate_rgb.bias": torch.randn(1, generator=g) * 0.1,
         "gate_depth.weight": torch.randn(1, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5, "gate_depth.bias": torch.randn(1, generator=g) * 0.1}
    gw = torch.cat([p["gate_rgb.weight"].reshape(1, -1), p["gate_depth.weight"].reshape(1, -1)], 0)
    gb = torch.cat([p["gate_rgb.bias"], p["gate_depth.bias"]])
    ro, do, mg, am = ops.rgbd_fusion(cu(r), cu(d), cu(gw), cu(gb), want_attn_mean=True)
    (rro, rdo), rmg = O.rgbd_fusion(p, r.float(), d.float())
    tol = dict(rtol=1e-4, atol=2e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    close(ro, rro, **tol), close(do, rdo, **tol), close(mg, rmg, **tol)
    assert abs(float(am.sum()) - 1.0) < 1e-4     # the two attention maps sum to one at every pixel
