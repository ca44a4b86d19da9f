"""GPU: the general-shape attention kernels (csrc/attn_general.cu) behind the exports of model/transfusion_head.py that KPFusion does not
run itself.  Checked against the unmodified reference's outputs (tests/golden/golden_heads.npz) and, at the reference's
real sizes (32 x 32 map = 1024 cells, 21 joints), against the oracle.  Tolerance: fp32 kernels vs fp32 reference, 2e-5 absolute on
O(1) activations (the bar of the fused decoder-layer test)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import kpf_oracle as O
from keypointfusion_b200 import ops
from keypointfusion_b200.model import transfusion_head as T
from keypointfusion_b200.utils import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


@pytest.fixture(scope="module")
def heads():
    return dict(np.load(os.path.join(GOLDEN, "golden_heads.npz"))), json.load(open(os.path.join(GOLDEN, "golden_heads_meta.json")))


def params(meta, name):
    return synth.fill_state_dict({k: torch.zeros(s) for k, s in meta[name].items()}, meta["seed"])


def close(a, b, atol, what=""):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(b) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max()
    print(f"[heads] {what}: max abs err {err:.2e}")
    assert err <= atol, f"{what}: max abs err {err:.3e}"


def d(a):
    return torch.from_numpy(a).to(DEV)


def test_linear_rows_strided_operands():
    """Every operand strided: a [P,B,K] input read as [B,P,K], an embedding-table position term, a channel-major output."""
    g = torch.Generator().manual_seed(3)
    for (B, P, K, Oc) in ((3, 37, 50, 70), (2, 130, 128, 256), (1, 5, 3, 9)):
        x, W, b = torch.randn(P, B, K, generator=g), torch.randn(Oc, K, generator=g), torch.randn(Oc, generator=g)
        table, idx = torch.randn(P + 4, K, generator=g), torch.randint(0, P + 4, (B, P), generator=g)
        ref = torch.relu(((x.transpose(0, 1) + table[idx]).double() @ W.double().t() + b.double()) * 0.37)
        out = torch.empty(B, Oc, P, device=DEV)
        ops.linear_rows(x.to(DEV).transpose(0, 1), W.to(DEV), b.to(DEV), pos=table.to(DEV), pos_index=idx.to(DEV), scale=0.37, relu=True,
                        out=out.transpose(1, 2))
        close(out.transpose(1, 2), ref, 2e-5 * max(1.0, float(ref.abs().max())), f"linear_rows {B}x{P}x{K}->{Oc}")
        pos = torch.randn(1, P, K, generator=g)
        ref = (x.transpose(0, 1) + pos).double() @ W.double().t()
        close(ops.linear_rows(x.to(DEV).transpose(0, 1), W.to(DEV), None, pos=pos.to(DEV)), ref, 2e-5 * max(1.0, float(ref.abs().max())), "no bias")


def test_position_embeddings_vs_reference(heads):
    g, meta = heads
    m = T.DetrSinePositionEmbedding(16, normalize=True)
    close(m(torch.zeros(2, 4, 6, 5, device=DEV), d(g["sine_mask"])), g["sine_norm"], 2e-6, "sine normalised, padded mask")
    m = T.DetrSinePositionEmbedding(8, temperature=100)
    close(m(torch.zeros(2, 4, 6, 5, device=DEV), d(g["sine_mask"])), g["sine_raw"], 2e-6, "sine raw")
    m = T.DetrSinePositionEmbedding(64, normalize=True)
    close(m(torch.zeros(1, 4, 10, 12, device=DEV), torch.ones(1, 10, 12, device=DEV)), g["sine_ones"], 2e-6, "sine ones")
    close(m.full_mask(10, 12, torch.device(DEV)), g["sine_ones"], 2e-6, "sine full mask")
    pe = T.PositionEmbeddingLearned(3, 32)
    pe.load_state_dict(params(meta, "pel_keys"))
    pe = pe.to(DEV).eval()
    with torch.no_grad():
        close(pe(d(g["pel_xyz"])), g["pel_out"], 2e-5, "PositionEmbeddingLearned")
    dl = T.DetrLearnedPositionEmbedding(16)
    dl.load_state_dict(params(meta, "dlearn_keys"))
    dl = dl.to(DEV)
    with torch.no_grad():
        assert np.array_equal(dl(torch.zeros(2, 4, 5, 7, device=DEV)).cpu().numpy(), g["dlearn_out"])


def test_multihead_attention_general_vs_reference(heads):
    """Any (L, S), head_dim 16 and 64, additive attn_mask + key_padding_mask, head-averaged weights; key is value and key is not value."""
    g, meta = heads
    for tag, (E, H) in {"mha_a": (64, 4), "mha_b": (128, 2)}.items():
        m = T.MultiheadAttention(E, H)
        synth.fill_state_dict(m, meta["seed"])
        m = m.to(DEV).eval()
        with torch.no_grad():
            o, w = m(d(g[tag + "_q"]), d(g[tag + "_k"]), d(g[tag + "_v"]), key_padding_mask=d(g[tag + "_kpm"]), need_weights=True,
                     attn_mask=d(g[tag + "_am"]))
            close(o, g[tag + "_out"], 2e-5, tag + " out")
            close(w, g[tag + "_w"], 2e-6, tag + " weights")
            k = d(g[tag + "_k"])
            o2, w2 = m(d(g[tag + "_q"]), k, k, need_weights=False)
            assert w2 is None
            close(o2, g[tag + "_out_kk"], 2e-5, tag + " key is value")


def test_fully_masked_row_is_nan_like_the_reference():
    m = T.MultiheadAttention(32, 2).to(DEV).eval()
    q, k = torch.randn(3, 1, 32, device=DEV), torch.randn(5, 1, 32, device=DEV)
    with torch.no_grad():
        o, _ = m(q, k, k, key_padding_mask=torch.ones(1, 5, dtype=torch.bool, device=DEV))
    assert torch.isnan(o).all()


def test_decoder_layer_with_self_attention_vs_reference(heads):
    g, meta = heads
    lay = T.TransformerDecoderLayer(128, 4, 64, 0.1, "relu", self_posembed=None, cross_posembed=None, cross_only=False)
    lay.load_state_dict(params(meta, "lay_keys"))
    lay = lay.to(DEV).eval()
    with torch.no_grad():
        out = lay(d(g["lay_q"]), d(g["lay_k"]), d(g["lay_qp"]), d(g["lay_kp"]))
    close(out, g["lay_out"], 2e-5, "decoder layer, self + cross attention")


def test_detr_decoder_and_spatial_aggregate_tr_vs_reference(heads):
    g, meta = heads
    for name, cls, args in (("detr", T.detrDecoder, ("dec_anchor", "dec_img")), ("satr", T.spatial_aggregate_TR, ("dec_img", "dec_anchor"))):
        m = cls(joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=2)
        m.load_state_dict(params(meta, name + "_keys"))
        m = m.to(DEV).eval()
        with torch.no_grad():
            out = m(d(g[args[0]]), d(g[args[1]]))
        close(out, g[name + "_out"], 2e-5, name)


def test_decoders_at_the_reference_sizes_vs_oracle():
    """32 x 32 feature map (1024 cells), 21 joints, batch 4, the reference's own constructor call (transfusion_head.py:788-796)."""
    gen = torch.Generator().manual_seed(5)
    anchors, img = torch.randn(4, 21, 128, generator=gen), torch.randn(4, 128, 32, 32, generator=gen)
    for name, cls in (("detr", T.detrDecoder), ("satr", T.spatial_aggregate_TR)):
        m = cls(joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=4, activation='relu')
        synth.fill_state_dict(m, 9)
        p = {k: v.clone() for k, v in m.state_dict().items()}
        m = m.to(DEV).eval()
        with torch.no_grad():
            if name == "detr":
                out, ref = m(anchors.to(DEV), img.to(DEV)), O.detr_decoder(p, "", anchors, img, 4)
            else:
                out, ref = m(img.to(DEV), anchors.to(DEV)), O.spatial_aggregate_tr(p, "", img, anchors, 4)
        close(out, ref, 2e-5, name + " 32x32")


def test_updated_decoder_general_route_matches_the_fused_kernels(golden, golden_meta):
    """The live configuration through the GENERAL kernels (explicit position indices force that route) == the reference golden."""
    dec = T.updatedDecoder(joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=4)
    dec.load_state_dict(synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["updatedDecoder_keys"].items()}, golden_meta["seed"]))
    dec = dec.to(DEV).eval()
    a, k = d(golden["a13_anchor"]), d(golden["a13_key"])
    idx = torch.arange(21, device=DEV).unsqueeze(0).expand(a.shape[0], -1)
    with torch.no_grad():
        close(dec.decoder[-1](a, k, idx, idx), golden["a13_out"], 2e-5, "updatedDecoder layer, general route")
        close(dec(a, k), golden["a13_out"], 2e-5, "updatedDecoder, fused route")


def test_mha_matches_torch_functional_and_the_live_golden(golden, golden_meta):
    """transfusion_head.py:303-556 was derived from torch's multi_head_attention_forward: same numbers for an (L, S) = (5, 9) call; and
    the reference's own output for the live layer's attention module (tests/golden/make_golden.py, a13_mha_*)."""
    torch.manual_seed(0)
    m = T.MultiheadAttention(64, 4)
    synth.fill_state_dict(m, 3)
    q, k = torch.randn(5, 2, 64), torch.randn(9, 2, 64)
    ro, rw = torch.nn.functional.multi_head_attention_forward(q, k, k, 64, 4, m.in_proj_weight, m.in_proj_bias, None, None, False, 0.0,
                                                              m.out_proj.weight, m.out_proj.bias, training=False)
    m = m.to(DEV).eval()
    with torch.no_grad():
        o, w = m(q.to(DEV), k.to(DEV), k.to(DEV))
    close(o, ro.detach(), 1e-5, "vs torch functional: out")
    close(w, rw.detach(), 1e-6, "vs torch functional: weights")
    sd = synth.fill_state_dict({k_: torch.zeros(s) for k_, s in golden_meta["updatedDecoder_keys"].items()}, golden_meta["seed"])
    m = T.MultiheadAttention(128, 4)
    m.load_state_dict({k_[len("decoder.3.multihead_attn."):]: v for k_, v in sd.items() if k_.startswith("decoder.3.multihead_attn.")})
    m = m.to(DEV).eval()
    with torch.no_grad():
        kk = d(golden["a13_mha_k"])
        o, w = m(d(golden["a13_mha_q"]), kk, kk)
    close(o, golden["a13_mha_out"], 2e-5, "live layer's attention: out")
    close(w, golden["a13_mha_w"], 1e-6, "live layer's attention: weights")


def test_general_ops_opcheck_and_compile():
    """torch.library.opcheck for the four general-shape custom ops (strided inputs, every optional argument) and a fullgraph
    torch.compile of detrDecoder: the decoders trace as opaque kpf ops, like the hot path."""
    from torch.library import opcheck
    K = torch.ops.kpf
    tests = ("test_schema", "test_faketensor")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(9, 2, 40, generator=g).to(DEV).transpose(0, 1)                       # [B,P,K] view of a [P,B,K] tensor
    W, b = torch.randn(24, 40, generator=g).to(DEV), torch.randn(24, generator=g).to(DEV)
    table, idx = torch.randn(12, 40, generator=g).to(DEV), torch.randint(0, 12, (2, 9), generator=g).to(DEV)
    for layout in (0, 1, 2):
        opcheck(K.linear_rows, (x, W, b, table, idx, 0.5, True, layout), test_utils=tests)
    opcheck(K.linear_rows, (x, W, None, torch.randn(1, 9, 40, generator=g).to(DEV), None, 1.0, False, 0), test_utils=tests)
    q, kv = torch.randn(2, 9, 64, generator=g).to(DEV), torch.randn(2, 33, 128, generator=g).to(DEV)
    am, kpm = torch.randn(9, 33, generator=g).to(DEV), (torch.rand(2, 33, generator=g) > 0.8).to(DEV)
    opcheck(K.mha_core, (q, kv[..., :64], kv[..., 64:], 4, am, kpm, True), test_utils=tests)
    opcheck(K.mha_core, (q, kv[..., :64], kv[..., 64:], 2, None, None, False), test_utils=tests)
    gam, bet = torch.rand(64, generator=g).to(DEV), torch.randn(64, generator=g).to(DEV)
    for cm in (False, True):
        opcheck(K.add_layernorm_rows, (q, torch.randn(2, 9, 64, generator=g).to(DEV), gam, bet, 1e-5, cm), test_utils=tests)
    opcheck(K.add_layernorm_rows, (torch.randn(2, 64, 9, generator=g).to(DEV).transpose(1, 2), None, gam, bet, 1e-5, True), test_utils=tests)
    opcheck(K.sine_posembed, (torch.rand(16, generator=g).to(DEV) + 1, 2, 6, 5, torch.ones(2, 6, 5, device=DEV), True, 6.283), test_utils=tests)
    opcheck(K.sine_posembed, (torch.rand(16, generator=g).to(DEV) + 1, 1, 6, 5, None, False, 6.283), test_utils=tests)

    det = T.detrDecoder(joint_num=21, num_decoder_layers=2)
    synth.fill_state_dict(det, 4)
    det = det.to(DEV).eval()
    anchors, img = torch.randn(3, 21, 128, generator=g).to(DEV), torch.randn(3, 128, 16, 16, generator=g).to(DEV)
    with torch.no_grad():
        eager = det(anchors, img)
        torch._dynamo.reset()
        got = torch.compile(det, fullgraph=True, backend="aot_eager")(anchors, img)
    assert torch.equal(eager, got)
