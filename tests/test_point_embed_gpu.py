"""GPU: fused point stage (repack + gather + embeddings + softmax-aggregation partials on tcgen05) vs the fp32 oracle."""
import numpy as np
import pytest
import torch

from keypointfusion_b200.utils import synth
from oracle import kpf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rms_rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


def test_repack(golden_inputs):
    from keypointfusion_b200 import ops
    i = golden_inputs
    for dt in (torch.float32, torch.bfloat16):
        out = ops.repack_features(i["img_feat"].to(DEV, dt), i["img_feat_rgb"].to(DEV, dt), i["img_offset"][:, 84:].to(DEV, dt))
        ref = torch.cat([i["img_feat"], i["img_feat_rgb"], i["img_offset"][:, 84:], torch.zeros(2, 11, 32, 32)], 1)
        ref = ref.reshape(2, 288, 1024).permute(0, 2, 1).to(dt).bfloat16()
        assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("B", [2, 5])
def test_point_embed(golden, golden_inputs, path_params, B):
    from keypointfusion_b200 import ops
    from keypointfusion_b200.model.model import Block_KPFusion
    inp = synth.make_inputs(B, 128, 21, 128, seed=50 + B, bf16_round=True)
    g = [inp[k].numpy() for k in ("center", "M", "cube", "cam")]
    c = {k: v.to(DEV) for k, v in inp.items()}
    pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
    close, _, idx = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True)
    joint = torch.from_numpy(np.random.RandomState(B).uniform(-0.6, 0.6, (B, 21, 3)).astype(np.float32))
    blk = Block_KPFusion(21)
    blk.load_state_dict({k[len("block1."):]: v for k, v in path_params.items() if k.startswith("block1.")})
    blk = blk.to(DEV).eval()
    k = blk.kc()
    featT = ops.repack_features(c["img_feat"].bfloat16(), c["img_feat_rgb"].bfloat16(), c["img_offset"][:, 84:].bfloat16())
    e, acc, ms = ops.point_embed(featT, idx, close, pcl, joint.to(DEV), k["pe_wmat"], k["pe_wvec"], 0.8)
    agg = ops.combine_point_partials(acc, ms, 21)
    # oracle on the same (bf16-rounded) maps
    p = path_params
    pc, ix, cl = pcl.cpu(), idx.cpu().long(), close.cpu()
    off = O.pcl_joint2offset(joint, pc, 0.8)
    pf, pr = O.gather_taps(inp["img_feat"], ix, cl), O.gather_taps(inp["img_feat_rgb"], ix, cl)
    pw = O.gather_taps(inp["img_offset"][:, 84:], ix, cl)
    ee = torch.relu(O.conv_bn(p, "block1.pcl_feat_emb.", pf) + O.conv_bn(p, "block1.pcl_xyz_emb.", pc) +
                    O.conv_bn(p, "block1.pcl_pose_emb.", torch.cat([pw, off], -1)))
    ee = torch.relu(ee + O.conv_bn(p, "block1.pcl_feat_emb_RGB.", pr))
    ragg = torch.softmax(pw.permute(0, 2, 1), -1) @ ee
    assert rms_rel(e, ee) < 1e-2, rms_rel(e, ee)
    assert rms_rel(agg, ragg) < 1e-2, rms_rel(agg, ragg)
    # spatial processing order (scheduling aid): K2 and the point features are bit-identical, the aggregation only regroups
    order = ops.spatial_order(pcl, c["center"], c["M"], c["cube"], c["cam"], 128, 32)
    assert torch.equal(torch.sort(order.long(), dim=1)[0].cpu(), torch.arange(pcl.shape[1]).expand(B, -1))
    close2, _, idx2 = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True,
                                        order=order)
    assert torch.equal(idx2, idx) and torch.equal(close2, close)
    e2, acc2, ms2 = ops.point_embed(featT, idx, close, pcl, joint.to(DEV), k["pe_wmat"], k["pe_wvec"], 0.8, order=order)
    assert torch.equal(e2, e)
    agg2 = ops.combine_point_partials(acc2, ms2, 21)
    assert rms_rel(agg2, ragg) < 1e-2 and rms_rel(agg2, agg) < 2e-3, (rms_rel(agg2, ragg), rms_rel(agg2, agg))


@pytest.mark.parametrize("B", [2, 3])
def test_desa_fused(path_params, B):
    """point stage -> DESA kernel vs the oracle's joint embeddings + DESA (same ball-query membership: centres are inputs)."""
    from keypointfusion_b200 import ops
    from keypointfusion_b200.model.model import Block_KPFusion
    import torch.nn.functional as F
    inp = synth.make_inputs(B, 128, 21, 128, seed=70 + B, bf16_round=True)
    c = {k: v.to(DEV) for k, v in inp.items()}
    pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=2)
    close, _, idx = ops.img2pcl_index(pcl, c["img"], c["center"], c["M"], c["cube"], c["cam"], 128, 4, fs=32, want_i64=False, want_i32=True)
    # joints placed on points of the cloud so that every ball is well populated
    joint = pcl[:, ::48][:, :21].contiguous() + 0.01
    blk = Block_KPFusion(21)
    blk.load_state_dict({k[len("block1."):]: v for k, v in path_params.items() if k.startswith("block1.")})
    blk = blk.to(DEV).eval()
    k = blk.kc()
    featT = ops.repack_features(c["img_feat"].bfloat16(), c["img_feat_rgb"].bfloat16(), c["img_offset"][:, 84:].bfloat16())
    e, acc, ms = ops.point_embed(featT, idx, close, pcl, joint, k["pe_wmat"], k["pe_wvec"], 0.8)
    part, jf = ops.desa_fused(e, acc, ms, pcl, joint, k["ds_wmat"], k["ds_wvec"], blk.FA.radius, blk.FA.S[0])
    fu = blk.FA.kc()["fusion"]
    out = F.relu(F.linear(torch.cat([part.permute(0, 2, 1, 3).reshape(B, 21, -1), jf], -1), *fu))
    # oracle
    p = path_params
    pc, ix, cl, jt = pcl.cpu(), idx.cpu().long(), close.cpu(), joint.cpu()
    off = O.pcl_joint2offset(jt, pc, 0.8)
    pf, pr = O.gather_taps(inp["img_feat"], ix, cl), O.gather_taps(inp["img_feat_rgb"], ix, cl)
    pw = O.gather_taps(inp["img_offset"][:, 84:], ix, cl)
    ee = torch.relu(O.conv_bn(p, "block1.pcl_feat_emb.", pf) + O.conv_bn(p, "block1.pcl_xyz_emb.", pc) +
                    O.conv_bn(p, "block1.pcl_pose_emb.", torch.cat([pw, off], -1)))
    ee = torch.relu(ee + O.conv_bn(p, "block1.pcl_feat_emb_RGB.", pr))
    rjf = torch.relu(O.conv_bn(p, "block1.joint_feat_emb.", torch.softmax(pw.permute(0, 2, 1), -1) @ ee) + O.conv_bn(p, "block1.joint_xyz_emb.", jt))
    rout = O.desa(p, "block1.FA.", ee, rjf, pc, jt)
    assert rms_rel(jf, rjf) < 1e-2, rms_rel(jf, rjf)
    assert rms_rel(out, rout) < 1e-2, rms_rel(out, rout)
