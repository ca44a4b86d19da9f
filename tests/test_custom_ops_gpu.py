"""GPU: the torch.library layer on the device: `torch.library.opcheck` (schema / fake-vs-real consistency / mutation declaration) for
the operators of the hot path, and `torch.compile(net.forward_path, fullgraph=True)` -- one traced graph with no graph break, whose
result is bit-identical to eager (the compiled graph calls the same kernels; torch.compile is not the hot path, it is the proof
that the drop-ins can live inside a compiled / exported model next to the PyTorch backbones)."""
import numpy as np
import pytest
import torch

from keypointfusion_b200.utils import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def net(path_params):
    from keypointfusion_b200.model.model import KPFusion
    n = KPFusion(joint_num=21)
    n.load_state_dict(path_params)
    return n.to(DEV).eval()


def _inputs(B, seed, bf16=True):
    inp = synth.make_inputs(B, 128, 21, 128, seed=seed, bf16_round=bf16)
    c = {k: v.to(DEV) for k, v in inp.items()}
    if bf16:
        for k in ("img_feat", "img_feat_rgb", "img_offset"):
            c[k] = c[k].bfloat16()
    return c


def test_opcheck_hot_path_ops(net):
    from torch.library import opcheck
    from keypointfusion_b200 import ops
    K = torch.ops.kpf
    c = _inputs(2, 5)
    tests = ("test_schema", "test_faketensor")   # (no autograd registration: inference-only operators)
    pcl, cnt = K.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], 1024, 3, False, 1.0, None)
    opcheck(K.getpcl, (c["img"], c["center"], c["cube"], c["M"], c["cam"], 1024, 3, False, 1.0, None), test_utils=tests)
    opcheck(K.offset2joint_weight, (c["img_offset"], c["img"], 0.8), test_utils=tests)
    juvd = K.offset2joint_weight(c["img_offset"], c["img"], 0.8)
    opcheck(K.uvd2xyz, (juvd, c["center"], c["M"], c["cube"], c["cam"], 128.0, 1.0), test_utils=tests)
    jxyz = K.uvd2xyz(juvd, c["center"], c["M"], c["cube"], c["cam"], 128.0, 1.0)
    img_down = c["img"][:, :, ::4, ::4]
    order = K.spatial_order(pcl, c["center"], c["M"], c["cube"], c["cam"], 128.0, 32, 1.0)
    opcheck(K.img2pcl_index, (pcl, img_down, c["center"], c["M"], c["cube"], c["cam"], 128.0, 4, 1.0, False, order), test_utils=tests)
    close, idx = K.img2pcl_index(pcl, img_down, c["center"], c["M"], c["cube"], c["cam"], 128.0, 4, 1.0, False, order)
    opcheck(K.repack_features, (c["img_feat"], c["img_feat_rgb"], c["img_offset"][:, 84:]), test_utils=tests)
    opcheck(K.repack_features, (c["img_feat"].float(), c["img_feat_rgb"].float(), c["img_offset"][:, 84:].float()), test_utils=tests)
    hi, lo = K.repack_features(c["img_feat"], c["img_feat_rgb"], c["img_offset"][:, 84:])
    k = net.block1.kc()
    args = (hi, lo, idx, close, pcl, jxyz, k["pe_wmat"], k["pe_wvec"], 0.8, ops.SPLIT_FMT, order)
    opcheck(K.point_embed, args, test_utils=tests)
    e, acc, ms = K.point_embed(*args)
    dargs = (e, acc, ms, pcl, jxyz, k["ds_wmat"], k["ds_wvec"], 0.1, 0.2, 0.4, 64, ops.SPLIT_FMT)
    opcheck(K.desa_fused, dargs, test_utils=tests)
    part, jf = K.desa_fused(*dargs)
    pk = k["tok_init"]
    targs = (pk.wmat, pk.wseq, pk.wvec, pk.cross, pk.pre, pk.D, pk.L, pk.F, pk.Fc, pk.J, pk.fmt, True, None, None, None, part, jf)
    opcheck(K.token_stack, targs, test_utils=tests)
    tok, r3d = K.token_stack(*targs)
    blk = net.block1
    sargs = (c["img_feat_rgb"], c["img_feat_rgb"].new_empty(0), r3d, img_down, c["center"], c["M"], c["cube"], c["cam"], k["wa_packed"],
             blk.atten_spatial.bias.detach(), blk.weight_dis.detach(), blk.fc_spatial2joint_feature.weight.detach(),
             blk.fc_spatial2joint_feature.bias.detach(), 128.0, 1.0, 0.8, 1.0, 10.0, ops.SPLIT_FMT, None)
    opcheck(K.spatial_aggregate_tc, sargs, test_utils=tests)
    opcheck(K.split_map, (c["img_feat_rgb"].float(),), test_utils=tests)


@pytest.mark.parametrize("bf16", [True, False])
def test_forward_path_compiles_fullgraph(net, bf16):
    from keypointfusion_b200.dataloader.loader import loader
    c = _inputs(3, 9, bf16)
    L = loader(img_size=128)

    def path(img_offset, img_feat, img_feat_rgb, img, center, M, cube, cam):
        pcl, _ = torch.ops.kpf.getpcl(img, center, cube, M, cam, 1024, 4, False, 1.0, None)
        res, sw, _ = net.forward_path(img_offset, img_feat, None, img_feat_rgb, img, pcl, L, center, M, cube, cam, 0.8)
        return res[2], res[3], res[4], res[5], sw[0], sw[1]
    args = (c["img_offset"], c["img_feat"], c["img_feat_rgb"], c["img"], c["center"], c["M"], c["cube"], c["cam"])
    with torch.no_grad():
        eager = path(*args)                      # also builds the packed-weight caches the traced graph treats as constants
        torch._dynamo.reset()
        compiled = torch.compile(path, fullgraph=True, backend="aot_eager")   # fullgraph: ANY graph break is an error
        got = compiled(*args)
    for a, b in zip(eager, got):
        assert torch.equal(a, b)
