"""GPU: crop + normalise front end vs fixtures produced by the reference's demo_RGBD.py methods: bit-exact crops."""
import os

import numpy as np
import pytest
import torch

from oracle import kpf_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_crop.npz")
DEV = "cuda"


@pytest.fixture(scope="module")
def gc():
    return dict(np.load(G))


def u16(a):
    return torch.from_numpy(a.astype(np.int32)).to(DEV).to(torch.uint16)


def test_demo_frame_exact(gc):
    from keypointfusion_b200 import ops
    d = u16(gc["box_depth"][None])
    c = ops.center_from_bbox(d, torch.from_numpy(gc["box_bbox"][None]))
    assert np.allclose(c.cpu().numpy()[0], gc["box_center"], rtol=0, atol=1e-9)
    img, M, c3, _ = ops.crop_depth(d, torch.from_numpy(gc["box_center"][None]), [250, 250, 250], gc["box_cam"])
    assert np.array_equal(img[0, 0].cpu().numpy(), gc["box_crop_d"])                       # normalised depth crop: bit-exact
    assert np.allclose(M[0].cpu().numpy(), gc["box_M"], rtol=1e-6, atol=1e-6)
    assert np.allclose(c3[0].cpu().numpy(), gc["box_com3d"], rtol=1e-6)
    rgb = ops.crop_rgb(torch.from_numpy(gc["box_rgb"][None]).to(DEV), torch.from_numpy(gc["box_center"][None]), [250, 250, 250], gc["box_cam"])
    assert np.array_equal(rgb[0].cpu().numpy(), gc["box_crop_rgb"])                        # BGR crop: bit-exact


def test_synthetic_640x480_batch_exact(gc):
    from keypointfusion_b200 import ops
    B = gc["syn_depth"].shape[0]
    d = u16(gc["syn_depth"])
    c = ops.center_from_bbox(d, torch.from_numpy(gc["syn_bbox"]))
    assert np.allclose(c.cpu().numpy(), gc["syn_center"], rtol=0, atol=1e-9)
    img, M, c3, _ = ops.crop_depth(d, torch.from_numpy(gc["syn_center"]), [250, 250, 250], gc["syn_cam"])
    assert np.array_equal(img[:, 0].cpu().numpy(), gc["syn_crop_d"])
    assert np.allclose(M.cpu().numpy(), gc["syn_M"], rtol=1e-6, atol=1e-6)
    rgb = ops.crop_rgb(torch.from_numpy(gc["syn_rgb"]).to(DEV), torch.from_numpy(gc["syn_center"]), [250, 250, 250], gc["syn_cam"])
    assert np.array_equal(rgb.cpu().numpy(), gc["syn_crop_rgb"])
    # centre computed on the GPU (not the golden one) gives the same crops: the fp64 means round to the same integer bounds
    img2, _, _, _ = ops.crop_depth(d, c, [250, 250, 250], gc["syn_cam"])
    assert torch.equal(img, img2)


def test_frontend_feeds_the_fusion_path(gc, path_params):
    """config-5 style: frames -> crop -> back-projection -> fusion path (feature maps synthetic), vs the oracle end to end."""
    from keypointfusion_b200.demo_RGBD import Model_RGBD
    from keypointfusion_b200.model.model import KPFusion
    from keypointfusion_b200.utils import synth
    fe = Model_RGBD(None, cam_para=tuple(gc["syn_cam"]), seed=3)
    B = gc["syn_depth"].shape[0]
    b = fe.prepare_batch(torch.from_numpy(gc["syn_rgb"]).to(DEV), u16(gc["syn_depth"]), torch.from_numpy(gc["syn_bbox"]))
    assert b["img"].shape == (B, 1, 128, 128) and b["pcl"].shape == (B, 1024, 3) and b["img_rgb"].shape == (B, 3, 128, 128)
    assert float(b["pcl"].abs().max()) <= 1.0                                               # clamp (demo_RGBD.py:332)
    for i in range(B):  # back-projection of the cropped frame == oracle on the same crop (bit-exact, clamped)
        ref, P = O.getpcl_sample(b["img"][i, 0].cpu().numpy(), b["center"][i].cpu().numpy(), b["cube"][i].cpu().numpy(),
                                 b["M"][i].cpu().numpy(), b["cam_para"][i].cpu().numpy(), seed=3, b=i, clamp=True)
        assert np.array_equal(b["pcl"][i].cpu().numpy(), ref)
    net = KPFusion(joint_num=21)
    net.load_state_dict(path_params)
    net = net.to(DEV).eval()
    fm = synth.make_feature_maps(B, seed=9)
    f = [torch.from_numpy(x).to(DEV) for x in fm]
    with torch.no_grad():
        res, sw, _ = net.forward_path(f[2], f[0], None, f[1], b["img"], b["pcl"], fe.depthloader, b["center"], b["M"], b["cube"], b["cam_para"], 0.8)
    ores, _, _ = O.fusion_path(path_params, b["img"].cpu(), b["pcl"].cpu(), torch.from_numpy(fm[2]), torch.from_numpy(fm[0]), torch.from_numpy(fm[1]),
                               b["center"].cpu().numpy(), b["M"].cpu().numpy(), b["cube"].cpu().numpy(), b["cam_para"].cpu().numpy())
    err = np.linalg.norm((res[-1].cpu().numpy() - ores[-1].numpy()) * 125.0, axis=-1).mean()
    assert err <= 0.05, err
