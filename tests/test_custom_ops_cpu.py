"""CPU: the torch.library layer (keypointfusion_b200/custom_ops.py).  Every operator the drop-in modules call on the hot path is a
registered `torch.ops.kpf.*` custom op with a fake (meta) implementation, so the whole post-backbone path can be shape-propagated
with FakeTensorMode -- no GPU, no kernel launch -- exactly as torch.compile / torch.export do when they trace it."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from keypointfusion_b200 import custom_ops
from keypointfusion_b200.dataloader.loader import loader
from keypointfusion_b200.model.model import KPFusion
from keypointfusion_b200.utils import synth


def test_ops_are_registered_with_schemas():
    need = {"getpcl", "offset2joint_weight", "uvd2xyz", "xyz2uvd", "spatial_order", "img2pcl_index", "repack_features", "split_map",
            "point_embed", "desa_fused", "token_stack", "spatial_aggregate_tc", "gather_taps", "joint2heatmap", "img2anchor_dis",
            "joint2offset", "pcl_joint2offset", "rgbd_fusion", "ac_fusion", "fsp", "cross_decoder_layer", "ball_query", "eval_errors"}
    assert need <= set(custom_ops.REGISTERED), need - set(custom_ops.REGISTERED)
    for n in need:
        op = getattr(torch.ops.kpf, n).default
        assert "Tensor" in str(op._schema)
    # desa_fused declares the one in-place effect of the path (the joints' rows appended behind the points of `e`)
    assert "Tensor(a0!) e" in str(torch.ops.kpf.desa_fused.default._schema)


@pytest.mark.parametrize("maps", [torch.bfloat16, torch.float32])
def test_fusion_path_shape_propagates_under_fake_tensors(maps):
    net = KPFusion(joint_num=21).eval()
    synth.fill_state_dict(net, 0)
    with torch.no_grad():
        net.block1.kc(), net.block2.kc()       # packing the weights needs no device
        with FakeTensorMode(allow_non_fake_inputs=True):
            B, dev = 3, "cuda"
            img = torch.empty(B, 1, 128, 128, device=dev)
            f_d = torch.empty(B, 128, 32, 32, device=dev, dtype=maps)
            f_rgb = torch.empty(B, 128, 32, 32, device=dev, dtype=maps)
            off = torch.empty(B, 105, 32, 32, device=dev, dtype=maps)
            center, cube, M, cam = (torch.empty(B, 3, device=dev), torch.empty(B, 3, device=dev), torch.empty(B, 3, 3, device=dev),
                                    torch.empty(B, 4, device=dev))
            pcl, count = torch.ops.kpf.getpcl(img, center, cube, M, cam, 1024, 0, False, 1.0, None)
            res, sw, _ = net.forward_path(off, f_d, None, f_rgb, img, pcl, loader(img_size=128), center, M, cube, cam, 0.8)
    assert [tuple(r.shape) for r in res[2:]] == [(B, 21, 3)] * 4 and all(r.dtype == torch.float32 for r in res[2:])
    assert [tuple(w.shape) for w in sw] == [(B, 21, 32, 32)] * 2
    assert res[2].device.type == "cuda"


def test_general_attention_heads_shape_propagate_under_fake_tensors():
    """The exports of transfusion_head.py off the live path reach their kernels through torch.ops.kpf.* too: detrDecoder,
    spatial_aggregate_TR, MultiheadAttention and the position embeddings trace with FakeTensors (no GPU, no launch)."""
    from keypointfusion_b200.model import transfusion_head as T
    need = {"linear_rows", "mha_core", "add_layernorm_rows", "sine_posembed"}
    assert need <= set(custom_ops.REGISTERED)
    det, sat, mha = T.detrDecoder(num_decoder_layers=2).eval(), T.spatial_aggregate_TR(num_decoder_layers=2).eval(), T.MultiheadAttention(128, 4).eval()
    pel, sine = T.PositionEmbeddingLearned(3, 32).eval(), T.DetrSinePositionEmbedding(64, normalize=True)
    with torch.no_grad(), FakeTensorMode(allow_non_fake_inputs=True):
        dev = "cuda"
        anchors, img = torch.empty(3, 21, 128, device=dev), torch.empty(3, 128, 32, 32, device=dev)
        assert tuple(det.to(dev)(anchors, img).shape) == (3, 128, 21)
        assert tuple(sat.to(dev)(img, anchors).shape) == (3, 128, 1024)
        o, w = mha.to(dev)(torch.empty(7, 2, 128, device=dev), torch.empty(45, 2, 128, device=dev), torch.empty(45, 2, 128, device=dev))
        assert tuple(o.shape) == (7, 2, 128) and tuple(w.shape) == (2, 7, 45)
        assert tuple(pel.to(dev)(torch.empty(2, 50, 3, device=dev)).shape) == (2, 32, 50)
        assert tuple(sine(img, torch.empty(3, 32, 32, device=dev)).shape) == (3, 128, 32, 32)
