"""Eager-PyTorch composition of the fusion path, device agnostic: the op chain a user of the reference executes on a B200 (BASELINE.md
section 3, SURVEY.md section 2: "the bar is: beat the PyTorch-eager composition on the same B200").

It follows the reference's PyTorch code op for op -- including what makes it slow on a GPU: the [B,N,HW,3] broadcast difference and
the [B,N,HW] distance matrix of img2pcl_index (dataloader/loader.py:956-957), the int64 index `repeat` in front of torch.gather
(model/model.py:297-306), torch.linalg.inv per helper call (loader.py:781), the [B,J,C,HW] product of the spatial aggregation
(model.py:337-340) -- with stock torch operators only (cuBLAS / cuDNN / ATen); pointnet2_ops' ball query (a CUDA extension the image
does not have) is restated with torch.sort.  Used ONLY by bench.py's `cuda_eager_baseline` leg and validated against the oracle in
tests/test_eager_baseline_cpu.py; nothing in keypointfusion_b200/ imports it.
"""
import math

import torch
import torch.nn.functional as F


def _coords(fs, device):
    t = 2.0 * (torch.arange(fs, device=device, dtype=torch.float32) + 0.5) / fs - 1.0
    return t.view(1, fs).expand(fs, fs).reshape(-1), t.view(fs, 1).expand(fs, fs).reshape(-1)   # u = column, v = row


def uvd2xyz(uvd, center, M, cube, cam, img_size, flip=1.0):        # loader.py:775-789
    B = uvd.shape[0]
    Mi = torch.linalg.inv(M)
    uv = (uvd[..., :2] + 1) * img_size / 2
    d = uvd[..., 2] * cube[:, 2:3] / 2.0 + center[:, 2:3]
    xw = Mi[:, 0, 0:1] * uv[..., 0] + Mi[:, 0, 1:2] * uv[..., 1] + Mi[:, 0, 2:3]
    yw = Mi[:, 1, 0:1] * uv[..., 0] + Mi[:, 1, 1:2] * uv[..., 1] + Mi[:, 1, 2:3]
    X = (xw - cam[:, 2:3]) * d / cam[:, 0:1]
    Y = flip * (yw - cam[:, 3:4]) * d / cam[:, 1:2]
    xyz = torch.stack([X, Y, d], -1)
    return (xyz - center.view(B, 1, 3)) / (cube.view(B, 1, 3) / 2.0)


def offset2joint_weight(offset, depth, kernel):                     # model.py:466-500
    offset = offset.float()
    B, C5, fs, _ = offset.shape
    J = C5 // 5
    d = F.interpolate(depth, [fs, fs]).reshape(B, 1, fs * fs)
    u, v = _coords(fs, offset.device)
    coords = torch.cat([u.view(1, 1, -1).expand(B, 1, -1), v.view(1, 1, -1).expand(B, 1, -1), d], 1)
    unit = offset[:, :3 * J].reshape(B, J, 3, -1)
    heat = offset[:, 3 * J:4 * J].reshape(B, J, -1)
    wgt = offset[:, 4 * J:].reshape(B, J, -1)
    m = d.lt(0.99).float()
    w = torch.softmax(wgt.masked_fill(d.gt(0.99).expand(B, J, -1), -1e8), -1)
    dist = kernel - (heat * m) * kernel
    return (((unit * m.unsqueeze(1)) * dist.unsqueeze(2) + coords.unsqueeze(1)) * w.unsqueeze(2)).sum(-1)


def img2pcl_index(pcl, img_down, center, M, cube, cam, img_size, K=4):   # loader.py:936-967
    B, _, H, W = img_down.shape
    u, v = _coords(W, pcl.device)
    uvd = torch.stack([u.expand(B, -1), v.expand(B, -1), img_down.reshape(B, H * W)], -1)
    cells = uvd2xyz(uvd, center, M, cube, cam, img_size)
    dist = torch.sum(torch.pow(pcl.unsqueeze(2) - cells.unsqueeze(1), 2), dim=-1)      # [B,N,HW,3] -> [B,N,HW]
    val, idx = torch.topk(dist, K, largest=False)
    c = 1 / (val + 1e-8)
    return c / (c.sum(-1, keepdim=True) + 1e-8), idx


def gather_taps(feat, index, closeness):                            # model.py:297-306
    B, C = feat.shape[:2]
    N, K = index.shape[1:]
    g = torch.gather(feat.reshape(B, C, -1).float(), -1, index.reshape(B, 1, N * K).repeat(1, C, 1)).view(B, C, N, K)
    return (g * closeness.unsqueeze(1)).sum(-1).permute(0, 2, 1)


def pcl_joint2offset(joint, pcl, kernel):                           # model.py:503-525
    B, J, _ = joint.shape
    off = joint.unsqueeze(2) - pcl.unsqueeze(1)
    dis = off.pow(2).sum(-1).sqrt()
    unit = off / (dis.unsqueeze(-1) + 1e-8)
    heat = (kernel - dis) / kernel
    m = heat.ge(0).float() * pcl[:, :, 2].lt(0.99).float().unsqueeze(1)
    unit = (unit * m.unsqueeze(-1)).permute(0, 2, 1, 3).reshape(B, -1, 3 * J)
    return torch.cat([unit, (heat * m).permute(0, 2, 1)], -1)


def conv_bn(p, prefix, x):                                           # Conv1d(k=1) + BatchNorm1d (eval), model.py:254-259
    y = F.linear(x, p[prefix + "0.weight"].squeeze(-1), p[prefix + "0.bias"])
    return F.batch_norm(y.transpose(1, 2), p[prefix + "1.running_mean"], p[prefix + "1.running_var"], p[prefix + "1.weight"],
                        p[prefix + "1.bias"], False, 0.0, 1e-5).transpose(1, 2)


def _bn_last(p, prefix, x):
    s = p[prefix + "weight"] / torch.sqrt(p[prefix + "running_var"] + 1e-5)
    return (x - p[prefix + "running_mean"]) * s + p[prefix + "bias"]


def ball_query(xyz, centers, radius, nsample):                      # pointnet2_ops 3.0.0 semantics (model.py:158, :174)
    B, N, _ = xyz.shape
    d2 = (centers.unsqueeze(2) - xyz.unsqueeze(1)).pow(2).sum(-1)                          # B J N
    ar = torch.arange(N, device=xyz.device).view(1, 1, N).expand_as(d2)
    key = torch.where(d2 < radius * radius, ar, torch.full_like(ar, N))
    srt = torch.sort(key, dim=-1)[0][..., :nsample]
    first = srt[..., :1]
    first = torch.where(first == N, torch.zeros_like(first), first)
    return torch.where(srt == N, first.expand_as(srt), srt)


def desa(p, prefix, pcl_feat, node_feat, pcl_xyz, node_xyz, radius=(0.1, 0.2, 0.4), nsample=64):   # model.py:129-204
    B, J, C = node_feat.shape
    xyz, feat = torch.cat([pcl_xyz, node_xyz], 1), torch.cat([pcl_feat, node_feat], 1)
    bi = torch.arange(B, device=xyz.device).view(B, 1, 1)
    outs = []
    for i, r in enumerate(radius):
        idx = ball_query(xyz, node_xyz, r, nsample)
        gx = (xyz[bi, idx] - node_xyz.unsqueeze(2)) / r
        gf = feat[bi, idx] - node_feat.unsqueeze(2)
        g = torch.relu(_bn_last(p, f"{prefix}bn_l0_blocks.{i}.", F.linear(gx, p[f"{prefix}conv_l0_blocks.{i}.weight"].view(-1, 3),
                                                                       p[f"{prefix}conv_l0_blocks.{i}.bias"])) +
                       _bn_last(p, f"{prefix}bn_f0_blocks.{i}.", F.linear(gf, p[f"{prefix}conv_f0_blocks.{i}.weight"].view(-1, C),
                                                                       p[f"{prefix}conv_f0_blocks.{i}.bias"])))
        w = p[f"{prefix}conv_blocks.{i}.0.weight"]
        g = torch.relu(_bn_last(p, f"{prefix}bn_blocks.{i}.0.", F.linear(g, w.view(w.shape[0], -1), p[f"{prefix}conv_blocks.{i}.0.bias"])))
        outs.append(g.max(2)[0])
    outs.append(node_feat)
    y = F.linear(torch.cat(outs, -1), p[prefix + "fusion.0.weight"].squeeze(-1), p[prefix + "fusion.0.bias"])
    return torch.relu(_bn_last(p, prefix + "fusion.1.", y))


def bert_encoder(p, prefix, x, layers=4, heads=4):                  # model.py:30-126 (transformers BertLayer)
    B, L, _ = x.shape
    h = p[prefix + "bert.position_embeddings.weight"][:L].unsqueeze(0) + F.linear(x, p[prefix + "bert.img_embedding.weight"],
                                                                                  p[prefix + "bert.img_embedding.bias"])
    C = h.shape[-1]
    hd = C // heads
    for i in range(layers):
        lp = f"{prefix}bert.encoder.layer.{i}."
        lin = lambda n, t: F.linear(t, p[lp + n + ".weight"], p[lp + n + ".bias"])
        q = lin("attention.self.query", h).view(B, L, heads, hd).transpose(1, 2)
        k = lin("attention.self.key", h).view(B, L, heads, hd).transpose(1, 2)
        v = lin("attention.self.value", h).view(B, L, heads, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1)
        ctx = (a @ v).transpose(1, 2).reshape(B, L, C)
        h = F.layer_norm(lin("attention.output.dense", ctx) + h, (C,), p[lp + "attention.output.LayerNorm.weight"],
                         p[lp + "attention.output.LayerNorm.bias"], 1e-12)
        h = F.layer_norm(lin("output.dense", F.gelu(lin("intermediate.dense", h))) + h, (C,), p[lp + "output.LayerNorm.weight"],
                         p[lp + "output.LayerNorm.bias"], 1e-12)
    return h, F.linear(h, p[prefix + "cls_head.weight"], p[prefix + "cls_head.bias"]) + F.linear(x, p[prefix + "residual.weight"],
                                                                                                 p[prefix + "residual.bias"])


def decoder_layer(p, prefix, anchor, tokens, heads=4):              # transfusion_head.py:132-173, :303-556
    B, J, C = anchor.shape
    hd = C // heads
    q_in = anchor + p[prefix + "self_posembed.weight"][:J]
    k_in = tokens + p[prefix + "cross_posembed.weight"][:J]
    Wi, bi = p[prefix + "multihead_attn.in_proj_weight"], p[prefix + "multihead_attn.in_proj_bias"]
    q = (F.linear(q_in, Wi[:C], bi[:C]) * hd ** -0.5).view(B, J, heads, hd).transpose(1, 2)
    kv = F.linear(k_in, Wi[C:], bi[C:])
    k, v = (t.reshape(B, J, heads, hd).transpose(1, 2) for t in (kv[..., :C], kv[..., C:]))
    o = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(B, J, C)
    o = F.linear(o, p[prefix + "multihead_attn.out_proj.weight"], p[prefix + "multihead_attn.out_proj.bias"])
    x = F.layer_norm(anchor + o, (C,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)
    y = F.linear(torch.relu(F.linear(x, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"])), p[prefix + "linear2.weight"],
                 p[prefix + "linear2.bias"])
    return F.layer_norm(x + y, (C,), p[prefix + "norm3.weight"], p[prefix + "norm3.bias"], 1e-5)


def block(p, pf, img_feat, img_feat_rgb, pcl, joint_xyz, close, index, img_offset, prev, img_down, center, M, cube, cam, img_size, J):
    B = pcl.shape[0]
    H = img_feat.shape[2]
    off = pcl_joint2offset(joint_xyz, pcl, 0.8)
    pf_d, pf_rgb, pw = gather_taps(img_feat, index, close), gather_taps(img_feat_rgb, index, close), gather_taps(img_offset[:, 4 * J:], index, close)
    e = torch.relu(conv_bn(p, pf + "pcl_feat_emb.", pf_d) + conv_bn(p, pf + "pcl_xyz_emb.", pcl) +
                   conv_bn(p, pf + "pcl_pose_emb.", torch.cat([pw, off], -1)))
    e = torch.relu(e + conv_bn(p, pf + "pcl_feat_emb_RGB.", pf_rgb))
    jf = torch.softmax(pw.permute(0, 2, 1), -1) @ e
    jf = torch.relu(conv_bn(p, pf + "joint_feat_emb.", jf) + conv_bn(p, pf + "joint_xyz_emb.", joint_xyz))
    jf = desa(p, pf + "FA.", e, jf, pcl, joint_xyz)
    tok, r3d = bert_encoder(p, pf + "init_TR.", jf)
    # joint2heatmap (generateFeature.py:584-600) and img2anchor_dis (loader.py:791-819)
    ar = torch.arange(H, device=pcl.device, dtype=torch.float32) + 0.5
    jx, jy = (r3d[..., 0:1] + 1) / 2 * H, (r3d[..., 1:2] + 1) / 2 * H
    hm = torch.exp(-(((ar.view(1, 1, 1, H) - jx.unsqueeze(-1)) / 0.8) ** 2 + ((ar.view(1, 1, H, 1) - jy.unsqueeze(-1)) / 0.8) ** 2) / 2.0)
    u, v = _coords(H, pcl.device)
    cells = uvd2xyz(torch.stack([u.expand(B, -1), v.expand(B, -1), img_down.reshape(B, -1)], -1), center, M, cube, cam, img_size)
    jx3 = uvd2xyz(r3d, center, M, cube, cam, img_size)
    gam = (1 / (10 * (cells.unsqueeze(1) - jx3.unsqueeze(2)).pow(2).sum(-1) + 1)).view(B, J, H, H)
    sw = torch.sigmoid(F.conv2d(torch.cat([img_feat_rgb.float(), hm], 1), p[pf + "atten_spatial.weight"], p[pf + "atten_spatial.bias"]))
    s = torch.sigmoid(p[pf + "weight_dis"])
    w = s * gam + (1 - s) * sw
    prod = torch.relu(w.unsqueeze(2) * img_feat_rgb.float().unsqueeze(1)).view(B, J, -1, H * H)           # [B,J,C,HW], model.py:337-340
    fj = F.linear(prod, p[pf + "fc_spatial2joint_feature.weight"], p[pf + "fc_spatial2joint_feature.bias"]).view(B, J, -1)
    if prev is not None:
        fj = torch.relu((fj + prev) / 2)
    rj = decoder_layer(p, pf + "crossTR.decoder.3.", fj, tok)
    _, r2d = bert_encoder(p, pf + "final_TR.", torch.cat([r3d, rj], 2))
    return r3d, r2d, fj, sw


@torch.no_grad()
def fusion_path(p, img, pcl, img_offset, img_feat, img_feat_rgb, center, M, cube, cam, img_size=128, kernel=0.8, J=21):
    """model.py:399-426 after the backbones.  p: state_dict on the tensors' device.  -> ([r3d_1, r2d_1, r3d_2, r2d_2], [sw_1, sw_2])."""
    H = img_feat.shape[2]
    img_feat, img_feat_rgb, img_offset = img_feat.float(), img_feat_rgb.float(), img_offset.float()
    joint_uvd = offset2joint_weight(img_offset, img, kernel)
    img_down = F.interpolate(img, [H, H])
    joint_xyz = uvd2xyz(joint_uvd, center, M, cube, cam, img_size)
    close, index = img2pcl_index(pcl, img_down, center, M, cube, cam, img_size, 4)
    res, sws, prev = [], [], None
    for i in range(2):
        r3d, r2d, prev, sw = block(p, f"block{i + 1}.", img_feat, img_feat_rgb, pcl, joint_xyz, close, index, img_offset, prev, img_down,
                                   center, M, cube, cam, img_size, J)
        res += [r3d, r2d]
        sws.append(sw)
        joint_xyz = r2d
    return res, sws
