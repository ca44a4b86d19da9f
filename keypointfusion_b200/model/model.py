"""Drop-in for the hot-path part of the reference's model/model.py: module functions (model.py:429-585),
TR_Encoder / KP_Interaction_TR (:30-126), DESA (:129-204), Block_KPFusion (:207-351) and the KPFusion glue (:354-426),
with the reference's names, forward signatures, tensor layouts and state_dict keys (tests/golden/golden_meta.json).

Everything after the two backbones runs on the B200 kernels behind include/kpf_b200.h.  Inference only: BatchNorm uses
running statistics (folded into the preceding 1x1 conv), dropout is identity, no autograd through the kernels.
"""
import math

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import custom_ops as cops   # registers torch.ops.kpf.*
from .. import ops
from ..util.generateFeature import GFM
from ..util.img2pcl import Pcl_utils
from .transfusion_head import updatedDecoder

BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------------------------------
# module-level functions (model/model.py:429-585)
# ----------------------------------------------------------------------------------------------------------------------
def img2pcl(img):  # model.py:429-437
    B, _, W, H = img.size()
    t = 2.0 * (torch.arange(W, device=img.device, dtype=torch.float32) + 0.5) / W - 1.0
    u = t.view(1, 1, 1, W).expand(B, 1, W, W)
    v = t.view(1, 1, W, 1).expand(B, 1, W, W)
    return torch.cat((u, v, img), dim=1).view(B, 3, H * W).permute(0, 2, 1)


def _is_scalar(k):
    return not torch.is_tensor(k)


def joint2offset(joint, img, kernel_size, feature_size):  # model.py:440-463 (no +1e-8 under the sqrt, :455)
    if _is_scalar(kernel_size):
        return torch.ops.kpf.joint2offset(joint, img, float(kernel_size), feature_size, 0.0)
    return ops.joint2offset(joint, img, kernel_size, feature_size, eps=0.0)   # per-joint kernel tensor (generateFeature.py:76-80)


def offset2joint_weight(offset, depth, kernel_size):  # model.py:466-500
    if _is_scalar(kernel_size):
        return torch.ops.kpf.offset2joint_weight(offset, depth, float(kernel_size))
    return ops.offset2joint_weight(offset, depth, kernel_size)


def pcl_joint2offset(joint, pcl, kernel_size):  # model.py:503-525
    if _is_scalar(kernel_size):
        return torch.ops.kpf.pcl_joint2offset(joint, pcl, float(kernel_size))
    return ops.pcl_joint2offset(joint, pcl, kernel_size)


def pcl_offset2joint_weight(pcl_result, pcl, kernel_size):  # model.py:528-555 (no caller in the reference)
    B, N, C5 = pcl_result.shape
    J = C5 // 5
    r = pcl_result.permute(0, 2, 1)
    offset, heat, weight = r[:, :J * 3].reshape(B, J, 3, N), r[:, J * 3:J * 4].reshape(B, J, 1, N), r[:, J * 4:].reshape(B, J, 1, N)
    w = F.softmax(weight.masked_fill(pcl[:, :, 2].gt(0.99).view(B, 1, 1, N), -1e8), dim=-1)
    k = ops.kernel_vec(kernel_size, J, pcl.device).view(1, J, 1, 1)
    return torch.sum((offset * (k - heat * k) + pcl.permute(0, 2, 1).reshape(B, 1, 3, N)) * w, dim=-1)


def _fold_bn(conv_w, conv_b, bn):
    """eval-mode BatchNorm folded into the preceding 1x1 conv: y = s*(Wx+b-mean)+beta."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    W = conv_w.reshape(conv_w.shape[0], -1) * s[:, None]
    b = (conv_b - bn.running_mean) * s + bn.bias
    return W.contiguous(), b.contiguous()


def _param_fingerprint(module):
    """(data_ptr, version) of every parameter and buffer below `module`: changes on .to(), load_state_dict (a copy_ bumps the
    version counter), in-place edits and optimiser steps."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


class _KernelCache:
    """Mixin: lazily built, device-resident folded/packed weights.  The cache is keyed on the parameters' storage pointers and
    version counters, so it is rebuilt after ANY change of the weights -- a parent-level `net.load_state_dict(ckpt)` (which
    recurses through `_load_from_state_dict` and never calls a child's `load_state_dict`), `.to()`, an in-place edit -- and a
    DataParallel replica (whose parameters are different tensors) never reuses the source module's device-0 cache."""

    def invalidate(self):
        self.__dict__["_kc"] = None
        self.__dict__["_kc_key"] = None
        for m in self.children():
            if hasattr(m, "invalidate"):
                m.invalidate()

    def kc(self):
        c = self.__dict__.get("_kc")
        if torch.compiler.is_compiling():   # tracing: the packed weights are graph constants; build them with one eager call first
            if c is None:
                raise RuntimeError(f"{type(self).__name__}: run one eager forward before torch.compile so the packed weights exist")
            return c
        key = _param_fingerprint(self)
        if c is None or self.__dict__.get("_kc_key") != key:
            with torch.no_grad():
                c = self._build_kc()
            self.__dict__["_kc"] = c
            self.__dict__["_kc_key"] = key
        return c


def _inference_only(module, *tensors):
    """The kernels have no backward and fold BatchNorm with running statistics: refuse to run where that would silently drop
    gradients or use the wrong statistics (a model swapped into the reference's train.py must not 'train')."""
    if module.training:
        raise RuntimeError(f"{type(module).__name__}: inference only (BatchNorm folded with running statistics, dropout = identity, no "
                           "autograd through the B200 kernels); call .eval()")
    if torch.is_grad_enabled() and (any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors)):
        raise RuntimeError(f"{type(module).__name__}: an input requires grad but the B200 kernels have no backward; wrap the call in "
                           "torch.no_grad() or detach the inputs")


# ----------------------------------------------------------------------------------------------------------------------
# BERT-style token encoder (model/model.py:30-126; transformers 4.25.1 BertEncoder).  Own minimal module tree with the
# reference's state_dict keys -- including the BertEmbeddings / BertPooler tables the reference constructs but never
# calls (model.py:35, :37) -- so checkpoints load unchanged and `transformers` is not needed on the path.
# ----------------------------------------------------------------------------------------------------------------------
class _Cfg:
    def __init__(self, **kw):
        self.vocab_size, self.max_position_embeddings, self.type_vocab_size = 30522, 512, 2  # config/config.json
        self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.intermediate_size = 128, 4, 4, 16
        self.layer_norm_eps, self.initializer_range, self.hidden_dropout_prob = 1e-12, 0.02, 0.1
        self.img_feature_dim, self.output_feature_dim = 128, 3
        self.__dict__.update(kw)


class _BertEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        # transformers 4.25.1 (the reference's pin) registers this as a PERSISTENT buffer, so the released checkpoints carry
        # `...bert.embeddings.position_ids`; newer transformers dropped it from the state_dict.  Registered persistent here and
        # tolerated when missing (see _load_from_state_dict) so that both key lists load with strict=True.
        self.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)).clone(), persistent=True)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if prefix + "position_ids" not in state_dict:
            state_dict = dict(state_dict)
            state_dict[prefix + "position_ids"] = self.position_ids
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class _BertSelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query, self.key, self.value = (nn.Linear(c.hidden_size, c.hidden_size) for _ in range(3))


class _BertSelfOutput(nn.Module):
    def __init__(self, c, d_in):
        super().__init__()
        self.dense = nn.Linear(d_in, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _BertAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _BertSelfAttention(c)
        self.output = _BertSelfOutput(c, c.hidden_size)


class _BertIntermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _BertAttention(c)
        self.intermediate = _BertIntermediate(c)
        self.output = _BertSelfOutput(c, c.intermediate_size)


class _BertEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(c) for _ in range(c.num_hidden_layers)])


class _BertPooler(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)


def _bert_init(module, std):
    for m in module.modules():  # BertPreTrainedModel._init_weights
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)


class TR_Encoder(nn.Module):  # model.py:30-103
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = _BertEmbeddings(config)
        self.encoder = _BertEncoder(config)
        self.pooler = _BertPooler(config)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.img_dim = config.img_feature_dim
        self.img_embedding = nn.Linear(self.img_dim, config.hidden_size, bias=True)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        _bert_init(self, config.initializer_range)


class KP_Interaction_TR(_KernelCache, nn.Module):  # model.py:106-126
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = TR_Encoder(config)
        self.cls_head = nn.Linear(config.hidden_size, config.output_feature_dim)
        self.residual = nn.Linear(config.img_feature_dim, config.output_feature_dim)
        _bert_init(self, config.initializer_range)

    def _build_kc(self):
        return {}

    def forward(self, img_feats, *unused, want_tokens=True, **unused_kw):
        """img_feats [B,J,D] -> (tokens [B,J,hidden], pred [B,J,3]).  model.py:45-103, :116-126 -- one launch of the
        split-precision tcgen05 token-stack kernel (csrc/token_stack.cu)."""
        _inference_only(self, img_feats)
        c = self.config
        J, D = img_feats.shape[1], img_feats.shape[2]
        if (c.hidden_size, c.num_attention_heads) != (128, 4) or J > 32 or not (D == 128 or 128 < D <= 144):
            raise NotImplementedError("token-stack kernel: hidden 128, 4 heads, <= 32 tokens, 128 <= input dim <= 144 (no fallback path)")
        k = self.kc()
        if J not in k:  # the position-embedding slice depends on the token count
            k[J] = ops.pack_token_program(J, enc=(self.state_dict(), "")).to(img_feats.device)
        return cops.run_token_program(k[J], x=img_feats, want_tokens=want_tokens)

    forward_tc = forward


# ----------------------------------------------------------------------------------------------------------------------
# DESA (model/model.py:129-204) -- pointnet2_ops' QueryAndGroup replaced by the kpf_ball_query kernel
# ----------------------------------------------------------------------------------------------------------------------
class DESA(_KernelCache, nn.Module):
    def __init__(self, in_channel, mlp, S=[64, 64, 64], radius=[0.1, 0.2, 0, 4]):
        super(DESA, self).__init__()
        self.S, self.radius, self.scale_num = S, radius, len(radius)
        self.groupers = nn.ModuleList()  # parameter-free in the reference (QueryAndGroup)
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        self.conv_l0_blocks, self.conv_f0_blocks = nn.ModuleList(), nn.ModuleList()
        self.bn_l0_blocks, self.bn_f0_blocks = nn.ModuleList(), nn.ModuleList()
        for i in range(self.scale_num):
            self.conv_l0_blocks.append(nn.Conv2d(3, mlp[0], 1))
            self.conv_f0_blocks.append(nn.Conv2d(in_channel, mlp[0], 1))
            self.bn_l0_blocks.append(nn.BatchNorm2d(mlp[0]))
            self.bn_f0_blocks.append(nn.BatchNorm2d(mlp[0]))
            last_channel = mlp[0]
            convs, bns = nn.ModuleList(), nn.ModuleList()
            for out_channel in mlp[1:]:
                convs.append(nn.Conv2d(last_channel, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.fusion = nn.Sequential(nn.Conv1d(in_channel + mlp[-1] * self.scale_num, in_channel, 1), nn.BatchNorm1d(in_channel),
                                    nn.ReLU())

    def _build_kc(self):
        k = {"l0": [], "f0": [], "mlp": []}
        for i in range(self.scale_num):
            k["l0"].append(_fold_bn(self.conv_l0_blocks[i].weight, self.conv_l0_blocks[i].bias, self.bn_l0_blocks[i]))
            k["f0"].append(_fold_bn(self.conv_f0_blocks[i].weight, self.conv_f0_blocks[i].bias, self.bn_f0_blocks[i]))
            k["mlp"].append([_fold_bn(c.weight, c.bias, b) for c, b in zip(self.conv_blocks[i], self.bn_blocks[i])])
        k["fusion"] = _fold_bn(self.fusion[0].weight, self.fusion[0].bias, self.fusion[1])
        k["scales"] = [(k["f0"][i][0], k["f0"][i][1], k["l0"][i][0], k["l0"][i][1], k["mlp"][i][0][0], k["mlp"][i][0][1])
                       for i in range(self.scale_num)] if all(len(m) == 1 for m in k["mlp"]) else None
        return k

    def forward(self, pcl_feat, node_feat, pcl_xyz, node_xyz):
        """pcl_feat [B,N,C], node_feat [B,J,C], pcl_xyz [B,N,3], node_xyz [B,J,3] -> [B,J,C].  model.py:166-204.
        Stand-alone call of the fused kernels: ball query + grouped two-layer MLP + max-pool (csrc/desa_fused.cu, the joints'
        features given instead of embedded there) and the 512 -> 128 fusion conv (csrc/token_stack.cu, prologue-only program).
        Inside Block_KPFusion the same kernels run with the joint embedding and init_TR fused in."""
        _inference_only(self, pcl_feat, node_feat)
        k = self.kc()
        B, J, C = node_feat.shape
        N = pcl_feat.shape[1]
        if k["scales"] is None or C != 128 or N % 64 or J > 32 or len(set(self.S)) != 1 or self.scale_num > 4:
            raise NotImplementedError("DESA kernels: 128 channels, one nsample for all scales, N % 64 == 0, single-conv MLPs (no fallback path)")
        dev = node_feat.device
        if "ds" not in k:
            eye, z = torch.eye(128, device=dev), torch.zeros(128, device=dev)
            k["ds"] = ops.pack_desa(eye, z, torch.zeros(128, 3, device=dev), z, k["scales"])
            k["fu"] = ops.pack_token_program(J, fusion=k["fusion"]).to(dev)
        e = ops.e_from_float(pcl_feat.float())
        part, _ = ops.desa_fused(e, None, None, pcl_xyz, node_xyz, k["ds"][0], k["ds"][1], self.radius, self.S[0], jf_in=node_feat)
        return cops.run_token_program(k["fu"], desa=part, jf=node_feat)[0]


# ----------------------------------------------------------------------------------------------------------------------
# Block_KPFusion (model/model.py:207-351)
# ----------------------------------------------------------------------------------------------------------------------
class Block_KPFusion(_KernelCache, nn.Module):
    def __init__(self, joint_num=21, feature_size=128, num_points=4):
        super(Block_KPFusion, self).__init__()
        self.joint_num = joint_num
        self.dim = 128
        self.feature_size = feature_size
        self.num_points = num_points
        # unused by forward in the reference too, kept for the state_dict contract (model.py:217-219)
        self.sampling_offsets = nn.Linear(self.feature_size, 2 * self.num_points, bias=True)
        self.attention_weights = nn.Linear(self.feature_size, self.num_points, bias=True)
        self.sampling_feature_embding = nn.Linear(self.dim, self.dim, bias=True)

        self.FA = DESA(128, [128, 128], [64, 64, 64], [0.1, 0.2, 0.4])
        self.init_TR = KP_Interaction_TR(_Cfg(img_feature_dim=128))     # model.py:222-233
        self.final_TR = KP_Interaction_TR(_Cfg(img_feature_dim=131))    # model.py:235-245
        self.crossTR = updatedDecoder(joint_num=joint_num, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1,
                                      num_decoder_layers=4, activation='relu')  # model.py:246-252

        def cb(i, o):
            return nn.Sequential(nn.Conv1d(i, o, 1), nn.BatchNorm1d(o))
        self.pcl_feat_emb = cb(self.dim, self.dim)
        self.pcl_xyz_emb = cb(3, self.dim)
        self.pcl_pose_emb = cb(self.joint_num * 5, self.dim)
        self.joint_feat_emb = cb(self.dim, self.dim)
        self.joint_xyz_emb = cb(3, self.dim)
        self.pcl_feat_emb_RGB = cb(self.dim, self.dim)

        self.sigmoid = nn.Sigmoid()
        self.atten_spatial = nn.Conv2d(feature_size + joint_num, joint_num, kernel_size=1, stride=1, bias=True)
        self.fc_spatial2joint_feature = nn.Linear(32 * 32, 1, bias=True)
        self.reduction_joint_feature = nn.Linear(self.dim * 2, self.dim, bias=True)
        self.reduction_joint_feature_update = nn.Conv1d(joint_num * 3, joint_num, kernel_size=1, stride=1)
        self.softmax = nn.Softmax(dim=-1)
        self.apply(self._init_weights)                                  # model.py:269
        self.cls_head = nn.Linear(128, 3)                               # model.py:270 (after apply)
        self.weight_dis = nn.Parameter(torch.zeros([1]))
        self.GFM_ = GFM()

    def _init_weights(self, m):  # model.py:275-285
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2. / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 0.001)
        elif isinstance(m, nn.ConvTranspose2d):
            nn.init.normal_(m.weight, std=0.001)

    def _build_kc(self):
        Wf, bf = _fold_bn(self.pcl_feat_emb[0].weight, self.pcl_feat_emb[0].bias, self.pcl_feat_emb[1])
        Wx, bx = _fold_bn(self.pcl_xyz_emb[0].weight, self.pcl_xyz_emb[0].bias, self.pcl_xyz_emb[1])
        Wp, bp = _fold_bn(self.pcl_pose_emb[0].weight, self.pcl_pose_emb[0].bias, self.pcl_pose_emb[1])
        Wr, br = _fold_bn(self.pcl_feat_emb_RGB[0].weight, self.pcl_feat_emb_RGB[0].bias, self.pcl_feat_emb_RGB[1])
        Wj, bj = _fold_bn(self.joint_feat_emb[0].weight, self.joint_feat_emb[0].bias, self.joint_feat_emb[1])
        Wjx, bjx = _fold_bn(self.joint_xyz_emb[0].weight, self.joint_xyz_emb[0].bias, self.joint_xyz_emb[1])
        pe_wmat, pe_wvec = ops.pack_point_embed(Wf, bf, Wx, bx, Wp, bp, Wr, br, self.joint_num)
        ds_wmat, ds_wvec = ops.pack_desa(Wj, bj, Wjx, bjx, self.FA.kc()["scales"])
        wa_packed = ops.pack_spatial_wa(self.atten_spatial.weight, self.joint_num, self.dim)
        dev = self.weight_dis.device
        tok_init = ops.pack_token_program(self.joint_num, enc=(self.init_TR.state_dict(), ""), fusion=self.FA.kc()["fusion"]).to(dev)
        tok_final = ops.pack_token_program(self.joint_num, cross=(self.crossTR.decoder[-1].state_dict(), ""),
                                           enc=(self.final_TR.state_dict(), "")).to(dev)
        return dict(tok_init=tok_init, tok_final=tok_final, wa_packed=wa_packed, ds_wmat=ds_wmat, ds_wvec=ds_wvec, pe_wmat=pe_wmat,
                    pe_wvec=pe_wvec)

    def forward(self, img_feat, img_feature_rgb, pcl, joint_xyz, pcl_closeness, pcl_index, img_offset, updated_2d_feature, loader,
                img_down, center, M, cube, cam_para, writer=None, ii=0, featT=None, rgb_planes=None, point_order=None, exchange=None,
                pe_stage=None, pe_stage_read=False):
        """model.py:287-351.  Five launches, all hand-written split-precision tcgen05 kernels (fp32-class results for bf16 AND
        fp32 feature maps -- an fp32 map is carried as two bf16 planes):
            point stage (K4b + K3 + embeddings + softmax partials) -> DESA -> [fusion conv + init_TR] -> K5 -> [crossTR + final_TR]
        `featT` / `rgb_planes`: the repacked maps / the rgb map's planes when the caller (KPFusion.forward_path) already made them
        for both blocks.  `pe_stage` (ops.point_embed_stage): the point stage stores its gathered, joint-independent operand tiles
        there (pe_stage_read False) or loads them from there instead of gathering again (True: a block that runs on the same
        maps, taps and point order as the one that filled it -- block 2 after block 1)."""
        _inference_only(self, img_feat, img_feature_rgb, joint_xyz, img_offset)
        k = self.kc()
        B, N, _ = pcl.shape
        C, H = img_feat.shape[1], img_feat.shape[2]
        J = self.joint_num
        if not (N % 64 == 0 and J <= 21 and self.FA.kc()["scales"] is not None and len(set(self.FA.S)) == 1 and self.FA.scale_num == 3
                and C == 128 and (H * H) % 128 == 0):
            raise NotImplementedError("Block_KPFusion kernels: N % 64 == 0 points, J <= 21 joints, 128-channel maps with H*H % 128 == 0, "
                                      "three DESA scales sharing nsample (no fallback path)")
        pcl = pcl.float().contiguous()
        joint_xyz = joint_xyz.detach().float().contiguous()
        K = torch.ops.kpf
        fmt = ops.SPLIT_FMT
        if featT is None:
            featT = K.repack_features(img_feat, img_feature_rgb, img_offset[:, J * 4:])
        if rgb_planes is None:
            rgb_planes = _rgb_planes(img_feature_rgb)
        if pcl_index.dtype != torch.int32:
            pcl_index = pcl_index.to(torch.int32)
        if pe_stage is None:
            e, p_acc, p_ms = K.point_embed(featT[0], featT[1], pcl_index, pcl_closeness, pcl, joint_xyz, k["pe_wmat"], k["pe_wvec"], 0.8, fmt,
                                           point_order)                                                   # model.py:295-320
        else:
            e, p_acc, p_ms = K.point_embed_staged(featT[0], featT[1], pcl_index, pcl_closeness, pcl, joint_xyz, k["pe_wmat"], k["pe_wvec"], 0.8,
                                                  fmt, pe_stage, pe_stage_read, point_order)
        r = self.FA.radius
        part, jf = K.desa_fused(e, p_acc, p_ms, pcl, joint_xyz, k["ds_wmat"], k["ds_wvec"], float(r[0]), float(r[1]), float(r[2]),
                                self.FA.S[0], fmt)                                                        # model.py:323-327
        outfeature_init_TR, refined_3d_joints = cops.run_token_program(k["tok_init"], desa=part, jf=jf)   # model.py:203, :330
        spatial_weight_loss, img_feat_j = K.spatial_aggregate_tc(
            rgb_planes[0], rgb_planes[1], refined_3d_joints, img_down, center, M, cube, cam_para, k["wa_packed"], self.atten_spatial.bias,
            self.weight_dis, self.fc_spatial2joint_feature.weight, self.fc_spatial2joint_feature.bias, float(loader.img_size),
            float(loader.flip), 0.8, 1.0, 10.0, fmt, updated_2d_feature)                                  # model.py:334-344
        _, refined_2d_joints = cops.run_token_program(k["tok_final"], x=img_feat_j, y=outfeature_init_TR, r3d=refined_3d_joints,
                                                      want_tokens=False, exchange=exchange)               # model.py:347-349
        return refined_3d_joints, refined_2d_joints, img_feat_j, spatial_weight_loss, None


def _rgb_planes(img_feat_rgb):
    """The rgb-branch map as K5's operand planes: a bf16 map is exact in one plane (second = empty tensor); an fp32 map is split."""
    if img_feat_rgb.dtype == torch.bfloat16:
        return img_feat_rgb.contiguous(), img_feat_rgb.new_empty(0)
    return torch.ops.kpf.split_map(img_feat_rgb.float().contiguous())


# ----------------------------------------------------------------------------------------------------------------------
# KPFusion glue (model/model.py:354-426)
# ----------------------------------------------------------------------------------------------------------------------
class KPFusion(nn.Module):
    """The reference builds its two backbones by name (model.py:363-373); they stay stock PyTorch and are passed in here
    (`backbone_rgb`, `backbone_d`: img -> (img_offset [B,5J,H,W], img_feat [B,128,H,W])).  With no backbones the module
    is the fusion path only (BASELINE.json config 2) and is driven through `forward_path`."""

    def __init__(self, net='KPFusion', pretrain='', joint_num=21, dataset='dexycb', mano_dir='', kernel_size=1, backbone_rgb=None,
                 backbone_d=None):
        super(KPFusion, self).__init__()
        self.joint_num, self.kernel_size, self.dim, self.classify_out, self.num_stages, self.net = joint_num, kernel_size, 128, 3, 2, net
        if backbone_rgb is not None:
            self.backbone_rgb = backbone_rgb
        if backbone_d is not None:
            self.backbone_d = backbone_d
        self.sigmoid = nn.Sigmoid()
        self.softmax = nn.Softmax(dim=-1)
        self.pcl_utils = Pcl_utils()
        for i in range(self.num_stages):
            setattr(self, f"block{i + 1}", Block_KPFusion(joint_num=joint_num))

    def forward_path(self, img_offset, img_feat, img_offset_rgb, img_feat_rgb, img, pcl, loader, center, M, cube, cam_para, kernel=0.8,
                     writer=None, ii=0, exchange=None):
        """model.py:399-426: everything after the backbones.  exchange: a runtime.PeerExchange -- the last block's final kernel then
        also writes its joints into every rank's gathered tensor (the path's one exchange step, fused)."""
        J = self.joint_num
        H = img_feat.shape[2]
        K = torch.ops.kpf
        img_size, flip = float(loader.img_size), float(loader.flip)
        joint_uvd = offset2joint_weight(img_offset, img, kernel)                                         # :399
        result = [img_offset, img_offset_rgb]
        S = img.shape[-1]
        img_down = img[:, :, ::S // H, ::S // H] if S % H == 0 else F.interpolate(img, [H, H])           # :409, zero-copy view
        joint_xyz = K.uvd2xyz(joint_uvd, center, M, cube, cam_para, img_size, flip)                      # :410
        # processing order of the points by feature-map cell: warps / point tiles then touch neighbouring cells (K2's insertions
        # coincide, the point stage's gathers hit the same lines); the results do not depend on it
        order = K.spatial_order(pcl, center, M, cube, cam_para, img_size, H, flip) if pcl.shape[1] <= 8192 else None
        pcl_closeness, pcl_index = K.img2pcl_index(pcl, img_down, center, M, cube, cam_para, img_size, 4, flip, False, order)   # :411
        updated_2d_feature = [None] * (self.num_stages + 1)
        spatial_weight = [None] * self.num_stages
        # shared by both blocks: the channels-last repack of the three maps (one plane for bf16 maps, two for fp32 maps) and the
        # rgb map's NCHW plane(s) for K5
        featT = K.repack_features(img_feat, img_feat_rgb, img_offset[:, J * 4:])
        rgb_planes = _rgb_planes(img_feat_rgb)
        # ... and the gathered point-stage operands: both blocks gather the same taps (model.py:297-306 per block); block 1 stores
        # its tiles, the later blocks load them with the TMA engine
        # (KPF_PE_STAGE=0: every block gathers for itself -- 175 MB less DRAM traffic per 64-sample step for ~6 us more kernel time)
        staged = self.num_stages > 1 and pcl.shape[1] % 64 == 0 and os.environ.get("KPF_PE_STAGE", "1") != "0"
        pe_stage = ops.point_embed_stage(pcl.shape[0], pcl.shape[1], pcl.device) if staged else None
        for i in range(self.num_stages):                                                                 # :417-424
            block = getattr(self, f"block{i + 1}")
            r3d, r2d, updated_2d_feature[i + 1], spatial_weight[i], _ = block(
                img_feat, img_feat_rgb, pcl, joint_xyz, pcl_closeness, pcl_index, img_offset, updated_2d_feature[i], loader, img_down,
                center, M, cube, cam_para, writer, ii, featT=featT, rgb_planes=rgb_planes, point_order=order,
                exchange=exchange if i == self.num_stages - 1 else None, pe_stage=pe_stage, pe_stage_read=i > 0)
            result.append(r3d)
            result.append(r2d)
            joint_xyz = r2d
        return result, spatial_weight, None

    @staticmethod
    def _as_backbone_input(x, backbone):
        """the crops arrive fp32; a backbone kept in bf16 / channels_last (stock PyTorch, the caller's choice) gets them in its own dtype"""
        p = next(backbone.parameters(), None)
        return x if p is None or p.dtype == x.dtype else x.to(p.dtype)

    def forward(self, img_rgb, img, pcl, loader, center, M, cube, cam_para, kernel=0.8, writer=None, ii=0):
        img_offset, img_feat = self.backbone_d(self._as_backbone_input(img, self.backbone_d))                  # model.py:397 (stock PyTorch)
        img_offset_rgb, img_feat_rgb = self.backbone_rgb(self._as_backbone_input(img_rgb, self.backbone_rgb))  # model.py:398
        return self.forward_path(img_offset.detach(), img_feat, img_offset_rgb, img_feat_rgb, img, pcl, loader, center, M, cube,
                                 cam_para, kernel, writer, ii)
