"""Drop-in for the reference's model/transfusion_head.py with its constructor arguments, init and state_dict keys.
Live part (KPFusion uses it): TransformerDecoderLayer(cross_only) + updatedDecoder -> the fused decoder-layer kernels (K6:
csrc/cross_attn.cu, csrc/token_stack.cu).  The rest of the file's exports -- MultiheadAttention for any (L, S) with
masks and averaged weights, TransformerDecoderLayer with self-attention / tensor position embeddings, detrDecoder,
spatial_aggregate_TR and the three position-embedding classes -- run on the general-shape fp32 kernels of csrc/attn_general.cu.
No torch matmul / softmax / LayerNorm anywhere.  Inference only."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Linear
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_
from torch.nn.parameter import Parameter

from .. import custom_ops as _cops   # noqa: F401  (registers torch.ops.kpf.*)
from .. import ops

_K = torch.ops.kpf   # the general-shape kernels are reached through their custom ops, like the hot path (traceable by torch.compile)


def _linear(x, weight, bias=None, pos=None, pos_index=None, scale=1.0, relu=False, out_layout=ops.ROWS_BPO):
    return _K.linear_rows(x, weight, bias, pos, pos_index, float(scale), bool(relu), int(out_layout))


def _attend(q, k, v, heads, attn_mask=None, key_padding_mask=None, need_weights=False):
    out, w = _K.mha_core(q, k, v, int(heads), attn_mask, key_padding_mask, bool(need_weights))
    return out, (w if need_weights else None)


def _add_ln(x, r, norm, channel_major=False):
    return _K.add_layernorm_rows(x, r, norm.weight, norm.bias, float(norm.eps), bool(channel_major))


class PositionEmbeddingLearned(nn.Module):
    """model/transfusion_head.py:16-32: Conv1d(k=1) -> BatchNorm1d -> ReLU -> Conv1d(k=1) over [B,P,in] -> [B,F,P].
    Two launches of the rows GEMM (csrc/attn_general.cu); BatchNorm folded from its running statistics (inference)."""

    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
            nn.BatchNorm1d(num_pos_feats),
            nn.ReLU(inplace=True),
            nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward(self, xyz):
        if self.training:
            raise RuntimeError("PositionEmbeddingLearned: inference only (BatchNorm is folded from its running statistics)")
        c1, bn, _, c2 = self.position_embedding_head
        g = bn.weight / torch.sqrt(bn.running_var + bn.eps)          # fold: y = g (W x + b - mean) + beta
        w1 = c1.weight[:, :, 0] * g[:, None]
        b1 = (c1.bias - bn.running_mean) * g + bn.bias
        h = _linear(xyz, w1, b1, relu=True)                                            # [B,P,F]
        return _linear(h, c2.weight[:, :, 0], c2.bias, out_layout=ops.ROWS_BOP)        # written channel-major: [B,F,P]  (:31)


class DetrLearnedPositionEmbedding(nn.Module):
    """model/transfusion_head.py:34-54: row / column embedding tables broadcast over the map -> [B, 2*dim, H, W].
    Pure indexing (two table look-ups, broadcast, concatenation): no arithmetic, so no kernel."""

    def __init__(self, embedding_dim=256):
        super().__init__()
        self.row_embeddings = nn.Embedding(50, embedding_dim)
        self.column_embeddings = nn.Embedding(50, embedding_dim)

    def forward(self, pixel_values, pixel_mask=None):
        height, width = pixel_values.shape[-2:]
        x_emb = self.column_embeddings.weight[:width]      # [W,D]
        y_emb = self.row_embeddings.weight[:height]        # [H,D]
        pos = torch.cat([x_emb.unsqueeze(0).expand(height, -1, -1), y_emb.unsqueeze(1).expand(-1, width, -1)], dim=-1)
        return pos.permute(2, 0, 1).unsqueeze(0).repeat(pixel_values.shape[0], 1, 1, 1)


class DetrSinePositionEmbedding(nn.Module):
    """model/transfusion_head.py:57-91 -> [B, 2*embedding_dim, H, W] (kpf_sine_posembed)."""

    def __init__(self, embedding_dim=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if scale is None:
            scale = 2 * math.pi
        self.scale = scale
        self._dim_t = None

    def dim_t(self, device):   # :84-85, evaluated once with the reference's own expression (no state_dict entry)
        if self._dim_t is None or self._dim_t.device != device:
            d = torch.arange(self.embedding_dim, dtype=torch.float32, device=device)
            self._dim_t = self.temperature ** (2 * torch.div(d, 2, rounding_mode="floor") / self.embedding_dim)
        return self._dim_t

    def forward(self, pixel_values, pixel_mask):
        if pixel_mask is None:
            raise ValueError("No pixel mask provided")
        B, H, W = pixel_mask.shape
        return _K.sine_posembed(self.dim_t(pixel_values.device), B, H, W, pixel_mask, bool(self.normalize), float(self.scale))

    def full_mask(self, H, W, device):
        """The embedding of an all-ones mask (what detrDecoder / spatial_aggregate_TR pass, :612, :770): [1, 2*dim, H, W], the same
        for every sample of a batch."""
        return _K.sine_posembed(self.dim_t(device), 1, H, W, None, bool(self.normalize), float(self.scale))


class MultiheadAttention(nn.Module):
    """model/transfusion_head.py:176-300.  forward(query[L,N,E], key[S,N,E], value) -> (out[L,N,E], weights[N,L,S])."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None):
        super(MultiheadAttention, self).__init__()
        if add_bias_kv or add_zero_attn or (kdim not in (None, embed_dim)) or (vdim not in (None, embed_dim)):
            raise NotImplementedError("only the configuration the reference instantiates (transfusion_head.py:100-101)")
        self.embed_dim, self.kdim, self.vdim = embed_dim, embed_dim, embed_dim
        self._qkv_same_embed_dim = True
        self.num_heads, self.dropout = num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.in_proj_weight = Parameter(torch.empty(3 * embed_dim, embed_dim))
        if bias:
            self.in_proj_bias = Parameter(torch.empty(3 * embed_dim))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = Linear(embed_dim, embed_dim, bias=bias)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self._reset_parameters()

    def _reset_parameters(self):  # transfusion_head.py:236-250
        xavier_uniform_(self.in_proj_weight)
        if self.in_proj_bias is not None:
            constant_(self.in_proj_bias, 0.)
            constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None):
        """multi_head_attention_forward (transfusion_head.py:303-556) for any L, S on the kernels of csrc/attn_general.cu, without the
        two torch.equal host syncs (:373-374): the q / k / v projections always use their own slices of in_proj_weight, which is what
        every branch of the reference computes.  The [L,N,E] / [S,N,E] inputs and the [L,N,E] output are addressed in place through
        strides (no transposed copies).  attn_mask: additive float [L,S] (or bool, True = masked); key_padding_mask: bool [N,S]."""
        if self.training and self.dropout > 0:
            raise RuntimeError("MultiheadAttention: inference only (attention dropout is not implemented in the B200 kernels)")
        E, H, hd = self.embed_dim, self.num_heads, self.head_dim
        L, N, _ = query.shape
        W, b = self.in_proj_weight, self.in_proj_bias
        bq, bk, bv = (None, None, None) if b is None else (b[:E], b[E:2 * E], b[2 * E:])
        q = _linear(query.transpose(0, 1), W[:E], bq, scale=float(hd) ** -0.5)             # [N,L,E]   :403, :468
        if key is value:
            kv = _linear(key.transpose(0, 1), W[E:], None if b is None else b[E:])          # [N,S,2E]  :418
            k, v = kv[..., :E], kv[..., E:]
        else:
            k = _linear(key.transpose(0, 1), W[E:2 * E], bk)
            v = _linear(value.transpose(0, 1), W[2 * E:], bv)
        a, w = _attend(q, k, v, H, attn_mask=attn_mask, key_padding_mask=key_padding_mask, need_weights=need_weights)
        out = _linear(a, self.out_proj.weight, self.out_proj.bias, out_layout=ops.ROWS_PBO)   # written sequence-first: [L,N,E]  :546
        return out, w


class TransformerDecoderLayer(nn.Module):
    """model/transfusion_head.py:94-173 (parameter container; the fused kernel consumes its packed weights)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", self_posembed=None,
                 cross_posembed=None, cross_only=False):
        super().__init__()
        self.cross_only = cross_only
        if not self.cross_only:
            self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)  # exists though unused when cross_only (state_dict contract)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1, self.dropout2, self.dropout3 = nn.Dropout(dropout), nn.Dropout(dropout), nn.Dropout(dropout)
        if activation != "relu":
            raise NotImplementedError("the reference builds these layers with activation='relu' (model.py:252)")
        self.activation = F.relu
        self.self_posembed = self_posembed
        self.cross_posembed = cross_posembed
        self.d_model, self.nhead, self.dim_feedforward = d_model, nhead, dim_feedforward
        self._wpack = None
        self._tc = None

    def _key(self, J):   # rebuilt after ANY weight change, however it was made (see model._KernelCache)
        return (J,) + tuple((t.data_ptr(), t._version) for t in self.parameters())

    def packed_tc(self, J):
        key = self._key(J)
        if self._tc is None or self._tc[0] != key:
            self._tc = (key, ops.pack_token_program(J, cross=(self.state_dict(), ""), C=self.d_model).to(self.linear1.weight.device))
        return self._tc[1]

    def packed(self, J):
        key = self._key(J)
        if self._wpack is None or self._wpack[0] != key:
            self._wpack = (key, ops.pack_decoder_layer(dict(self.state_dict()), "", J, self.d_model))
        return self._wpack[1]

    def _pos_args(self, embed, pos, P):
        """-> kwargs for ops.linear_rows: the position term of with_pos_embed (transfusion_head.py:141-155).  `embed` an nn.Embedding:
        `pos` are indices [B,P] (None = arange, what every decoder of the reference passes) looked up inside the kernel; `embed` None:
        `pos` is the embedding itself, [B or 1,P,C] with any strides; any other module is evaluated and must return [B,P,C]."""
        if isinstance(embed, nn.Embedding):
            if pos is None:
                return dict(pos=embed.weight[:P].unsqueeze(0))
            return dict(pos=embed.weight, pos_index=pos)
        if embed is not None:
            pos = embed(pos)
        return dict(pos=pos) if pos is not None else {}

    def forward(self, query, key, query_pos=None, key_pos=None, attn_mask=None, out_jc=None, out_jc_c0=0, want_cj=True,
                precision="fp32"):
        """query [B,Pq,C], key [B,Pk,C] (any strides) -> [B,C,Pq] (transfusion_head.py:132-173).
        The configuration updatedDecoder builds (cross_only, two nn.Embedding position tables indexed by the token number, Pq = Pk =
        J <= 32) runs as ONE fused kernel: precision "fp32" = CUDA cores (csrc/cross_attn.cu), "tc" = split-precision tcgen05
        (csrc/token_stack.cu, the one Block_KPFusion fuses with final_TR).  Everything else -- self-attention (cross_only=False), H*W
        keys (detrDecoder) or H*W queries (spatial_aggregate_TR), tensor position embeddings, attn_mask -- runs on the general-shape
        fp32 kernels of csrc/attn_general.cu.  All hand-written, all fp32-class."""
        if self.training or (torch.is_grad_enabled() and (query.requires_grad or key.requires_grad)):
            raise RuntimeError("TransformerDecoderLayer: inference only (no autograd through the B200 kernels); use .eval() and torch.no_grad()")
        B, Pq, C = query.shape
        Pk = key.shape[1]
        fused = (self.cross_only and attn_mask is None and query_pos is None and key_pos is None and Pq == Pk and Pq <= 32
                 and isinstance(self.self_posembed, nn.Embedding) and isinstance(self.cross_posembed, nn.Embedding)
                 and self.self_posembed.num_embeddings == Pq and self.cross_posembed.num_embeddings == Pk)
        if fused:
            J = Pq
            if precision in ("tc", "bf16") and self.d_model == 128 and self.nhead == 4 and self.dim_feedforward in (16, 128):
                return ops.token_stack(self.packed_tc(J), x=query, y=key, out_jc=out_jc, out_jc_c0=out_jc_c0, want_cj=want_cj)[2]
            if C <= 256 and self.dim_feedforward <= 256:
                return ops.cross_decoder_layer(query, key, self.packed(J), self.nhead, self.dim_feedforward, out_jc, out_jc_c0, want_cj)
        if out_jc is not None:
            raise NotImplementedError("out_jc is an extension of the fused 21-token path only")
        H, hd = self.nhead, C // self.nhead
        qpos = self._pos_args(self.self_posembed, query_pos, Pq)
        kpos = self._pos_args(self.cross_posembed, key_pos, Pk)
        x = query
        if not self.cross_only:   # :157-161  q = k = v = query + pos
            m = self.self_attn
            W, b = m.in_proj_weight, m.in_proj_bias
            q = _linear(x, W[:C], None if b is None else b[:C], scale=float(hd) ** -0.5, **qpos)
            kv = _linear(x, W[C:], None if b is None else b[C:], **qpos)             # [B,Pq,2C]
            a, _ = _attend(q, kv[..., :C], kv[..., C:], H)
            a = _linear(a, m.out_proj.weight, m.out_proj.bias)
            x = _add_ln(x, a, self.norm1)
        m = self.multihead_attn   # :163-167  value = key + key_pos as well
        W, b = m.in_proj_weight, m.in_proj_bias
        q = _linear(x, W[:C], None if b is None else b[:C], scale=float(hd) ** -0.5, **qpos)
        kv = _linear(key, W[C:], None if b is None else b[C:], **kpos)               # [B,Pk,2C]
        a, _ = _attend(q, kv[..., :C], kv[..., C:], H, attn_mask=attn_mask)
        a = _linear(a, m.out_proj.weight, m.out_proj.bias)
        x = _add_ln(x, a, self.norm2)
        h = _linear(x, self.linear1.weight, self.linear1.bias, relu=True)            # :169-171
        y = _linear(h, self.linear2.weight, self.linear2.bias)
        return _add_ln(x, y, self.norm3, channel_major=True)                         # :172 [B,C,Pq]


class updatedDecoder(nn.Module):
    """model/transfusion_head.py:635-708."""

    def __init__(self, joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=3,
                 activation='relu', bn_momentum=0.1, img_feature_seq_length=1024):
        super(updatedDecoder, self).__init__()
        self.decoder = nn.ModuleList()
        self.bn_momentum = bn_momentum
        self.num_decoder_layers = num_decoder_layers
        for i in range(self.num_decoder_layers):
            self.decoder.append(TransformerDecoderLayer(hidden_channel, num_heads, ffn_channel, dropout, activation,
                                                        self_posembed=nn.Embedding(joint_num, hidden_channel),
                                                        cross_posembed=nn.Embedding(joint_num, hidden_channel),
                                                        cross_only=True))
        self.joint_num = joint_num
        self.init_weights()

    def init_weights(self):  # transfusion_head.py:668-680
        for m in self.decoder.parameters():
            if m.dim() > 1:
                nn.init.xavier_uniform_(m)
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.momentum = self.bn_momentum

    def invalidate(self):
        for layer in self.decoder:
            layer._wpack = None
            layer._tc = None

    def forward(self, anchor_feats, img_feats, out_jc=None, out_jc_c0=0, want_cj=True, precision="fp32"):
        """anchor_feats [B,J,C] (queries), img_feats [B,J,C] (keys) -> [B,C,J].  Every layer of the reference gets the
        same inputs and only the last output is returned (transfusion_head.py:705-708): layers 0..n-2 are dead compute
        and are skipped; their parameters stay in the state_dict."""
        B, J, C = img_feats.shape
        assert anchor_feats.shape[1] == self.joint_num and J == self.joint_num
        return self.decoder[-1](anchor_feats, img_feats, out_jc=out_jc, out_jc_c0=out_jc_c0, want_cj=want_cj, precision=precision)


class _SineDecoder(nn.Module):
    """Shared constructor body of detrDecoder / spatial_aggregate_TR (transfusion_head.py:561-604, :712-755)."""

    def _build(self, joint_num, hidden_channel, num_heads, ffn_channel, dropout, num_decoder_layers, activation, bn_momentum, self_pos,
               cross_pos, sine_name):
        self.decoder = nn.ModuleList()
        self.bn_momentum = bn_momentum
        self.num_decoder_layers = num_decoder_layers
        for i in range(self.num_decoder_layers):
            self.decoder.append(TransformerDecoderLayer(
                hidden_channel, num_heads, ffn_channel, dropout, activation,
                self_posembed=nn.Embedding(joint_num, hidden_channel) if self_pos else None,
                cross_posembed=nn.Embedding(joint_num, hidden_channel) if cross_pos else None, cross_only=True))
        self.joint_num = joint_num
        setattr(self, sine_name, DetrSinePositionEmbedding(hidden_channel // 2, normalize=True))
        self.init_weights()

    def init_weights(self):
        for m in self.decoder.parameters():
            if m.dim() > 1:
                nn.init.xavier_uniform_(m)
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.momentum = self.bn_momentum


class detrDecoder(_SineDecoder):
    """model/transfusion_head.py:560-632: the J joint tokens attend over the H*W cells of an image feature map (keys = values =
    cells + sine position embedding; queries = tokens + a learned per-joint embedding)."""

    def __init__(self, joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=3,
                 activation='relu', bn_momentum=0.1, img_feature_seq_length=1024):
        super(detrDecoder, self).__init__()
        self._build(joint_num, hidden_channel, num_heads, ffn_channel, dropout, num_decoder_layers, activation, bn_momentum,
                    self_pos=True, cross_pos=False, sine_name="key_position_embedding")

    def forward(self, anchor_feats, img_feats):
        """anchor_feats [B,J,C], img_feats [B,C,W,H] -> [B,C,J].  Every layer of the reference gets the same inputs and only the last
        output is returned (:627-631): layers 0..n-2 are dead compute and are skipped.  The feature map and its position embedding
        are read in place as [B,H*W,C] rows through strides (the reference's flatten + permute, :614-615, without the copies)."""
        B, C, W, H = img_feats.shape
        assert anchor_feats.shape[1] == self.joint_num
        key_pos = self.key_position_embedding.full_mask(W, H, img_feats.device).flatten(2).permute(0, 2, 1)   # [1,WH,C]
        keys = img_feats.flatten(2).permute(0, 2, 1)                                                           # [B,WH,C] view
        return self.decoder[-1](anchor_feats, keys, None, key_pos)


class spatial_aggregate_TR(_SineDecoder):
    """model/transfusion_head.py:711-783: the H*W cells of an image feature map (queries, + sine position embedding) attend over the
    J joint tokens (keys = values = tokens + a learned per-joint embedding)."""

    def __init__(self, joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=3,
                 activation='relu', bn_momentum=0.1, img_feature_seq_length=1024):
        super(spatial_aggregate_TR, self).__init__()
        self._build(joint_num, hidden_channel, num_heads, ffn_channel, dropout, num_decoder_layers, activation, bn_momentum,
                    self_pos=False, cross_pos=True, sine_name="query_position_embedding")

    def forward(self, img_feats, anchor_feats):
        """img_feats [B,C,W,H], anchor_feats [B,J,C] -> [B,C,W*H]  (:758-783; last layer only, like the reference's return value)."""
        B, C, W, H = img_feats.shape
        assert anchor_feats.shape[1] == self.joint_num
        query_pos = self.query_position_embedding.full_mask(W, H, img_feats.device).flatten(2).permute(0, 2, 1)
        queries = img_feats.flatten(2).permute(0, 2, 1)
        return self.decoder[-1](queries, anchor_feats, query_pos, None)
