"""Drop-in for the live part of the reference's model/transfusion_head.py: MultiheadAttention, TransformerDecoderLayer
and updatedDecoder, with the reference's constructor arguments, init and state_dict keys.  updatedDecoder.forward runs
the fused B200 decoder-layer kernel (K6).  Inference only."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Linear
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_
from torch.nn.parameter import Parameter

from .. import ops


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
        """General (any L, S) path of multi_head_attention_forward (transfusion_head.py:303-556) without the two
        torch.equal host syncs (:373-374): q/k/v projections are always applied with their own weight slices, which
        is what every branch computes.  Runs as device library calls; the fused kernel path is updatedDecoder."""
        E, H, hd = self.embed_dim, self.num_heads, self.head_dim
        L, N, _ = query.shape
        S = key.shape[0]
        W, b = self.in_proj_weight, self.in_proj_bias
        q = F.linear(query, W[:E], None if b is None else b[:E]) * (float(hd) ** -0.5)
        k = F.linear(key, W[E:2 * E], None if b is None else b[E:2 * E])
        v = F.linear(value, W[2 * E:], None if b is None else b[2 * E:])
        q = q.contiguous().view(L, N * H, hd).transpose(0, 1)
        k = k.contiguous().view(S, N * H, hd).transpose(0, 1)
        v = v.contiguous().view(S, N * H, hd).transpose(0, 1)
        w = torch.bmm(q, k.transpose(1, 2))
        if attn_mask is not None:
            w = w + attn_mask.unsqueeze(0)
        if key_padding_mask is not None:
            w = w.view(N, H, L, S).masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2), float('-inf')).view(N * H, L, S)
        w = F.softmax(w, dim=-1)
        w = F.dropout(w, p=self.dropout, training=self.training)
        o = torch.bmm(w, v).transpose(0, 1).contiguous().view(L, N, E)
        o = F.linear(o, self.out_proj.weight, self.out_proj.bias)
        return o, (w.view(N, H, L, S).sum(dim=1) / H if need_weights else None)


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

    def forward(self, query, key, query_pos=None, key_pos=None, attn_mask=None, out_jc=None, out_jc_c0=0, want_cj=True,
                precision="fp32"):
        """query [B,J,C], key [B,J,C] -> [B,C,J] (transfusion_head.py:132-173, cross_only, index position embeddings).
        precision "fp32": CUDA-core kernel (csrc/cross_attn.cu); "tc": split-precision tcgen05 kernel (csrc/token_stack.cu, the one
        Block_KPFusion fuses with final_TR) -- both hand-written, both fp32-class."""
        if self.training or (torch.is_grad_enabled() and (query.requires_grad or key.requires_grad)):
            raise RuntimeError("TransformerDecoderLayer: inference only (no autograd through the B200 kernels); use .eval() and torch.no_grad()")
        if not self.cross_only or attn_mask is not None or self.self_posembed is None or self.cross_posembed is None:
            raise NotImplementedError("only the cross_only configuration updatedDecoder builds (transfusion_head.py:652-661)")
        J = query.shape[1]
        if precision in ("tc", "bf16") and self.d_model == 128 and self.nhead == 4 and J <= 32 and self.dim_feedforward in (16, 128):
            return ops.token_stack(self.packed_tc(J), x=query, y=key, out_jc=out_jc, out_jc_c0=out_jc_c0, want_cj=want_cj)[2]
        return ops.cross_decoder_layer(query, key, self.packed(J), self.nhead, self.dim_feedforward, out_jc, out_jc_c0, want_cj)


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
