"""CPU: the exports of model/transfusion_head.py off KPFusion's live path (position embeddings, general MultiheadAttention,
TransformerDecoderLayer with self-attention, detrDecoder, spatial_aggregate_TR).  Pins the oracle's restatements against
tests/golden/golden_heads.npz (the unmodified reference, tests/golden/make_golden_heads.py) and checks the drop-in classes' state_dict
contract.  The kernels themselves are checked in tests/test_heads_gpu.py."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import kpf_oracle as O
from keypointfusion_b200.model import transfusion_head as T
from keypointfusion_b200.utils import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def heads():
    return dict(np.load(os.path.join(GOLDEN, "golden_heads.npz"))), json.load(open(os.path.join(GOLDEN, "golden_heads_meta.json")))


def params(meta, name):
    return synth.fill_state_dict({k: torch.zeros(s) for k, s in meta[name].items()}, meta["seed"])


def close(a, b, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= atol, f"max abs err {np.abs(a - b).max():.3e}"


def t(a):
    return torch.from_numpy(a)


def test_oracle_position_embeddings(heads):
    g, meta = heads
    close(O.sine_position_embedding(t(g["sine_mask"]), 16, normalize=True), g["sine_norm"], 1e-6)
    close(O.sine_position_embedding(t(g["sine_mask"]), 8, temperature=100), g["sine_raw"], 1e-6)
    close(O.sine_position_embedding(torch.ones(1, 10, 12), 64, normalize=True), g["sine_ones"], 1e-6)
    close(O.position_embedding_learned(params(meta, "pel_keys"), "position_embedding_head.", t(g["pel_xyz"])), g["pel_out"], 1e-5)


def test_oracle_mha_general(heads):
    g, meta = heads
    for tag, (E, H) in {"mha_a": (64, 4), "mha_b": (128, 2)}.items():
        sd = {"in_proj_weight": torch.zeros(3 * E, E), "in_proj_bias": torch.zeros(3 * E), "out_proj.weight": torch.zeros(E, E),
              "out_proj.bias": torch.zeros(E)}
        p = synth.fill_state_dict(sd, meta["seed"])
        o, w = O.mha_forward(p, "", t(g[tag + "_q"]), t(g[tag + "_k"]), t(g[tag + "_v"]), H, key_padding_mask=t(g[tag + "_kpm"]),
                             attn_mask=t(g[tag + "_am"]))
        close(o, g[tag + "_out"], 5e-6)
        close(w, g[tag + "_w"], 1e-6)
        close(O.mha_forward(p, "", t(g[tag + "_q"]), t(g[tag + "_k"]), t(g[tag + "_k"]), H)[0], g[tag + "_out_kk"], 5e-6)


def test_oracle_decoder_layer_and_decoders(heads):
    g, meta = heads
    out = O.decoder_layer(params(meta, "lay_keys"), "", t(g["lay_q"]), t(g["lay_k"]), t(g["lay_qp"]), t(g["lay_kp"]), cross_only=False)
    close(out, g["lay_out"], 1e-5)
    close(O.detr_decoder(params(meta, "detr_keys"), "", t(g["dec_anchor"]), t(g["dec_img"]), 2), g["detr_out"], 1e-5)
    close(O.spatial_aggregate_tr(params(meta, "satr_keys"), "", t(g["dec_img"]), t(g["dec_anchor"]), 2), g["satr_out"], 1e-5)


def test_state_dict_contract(heads):
    """Same keys and shapes as the reference's modules, so its checkpoints load strictly."""
    _, meta = heads
    mods = {"pel_keys": T.PositionEmbeddingLearned(3, 32), "dlearn_keys": T.DetrLearnedPositionEmbedding(16),
            "lay_keys": T.TransformerDecoderLayer(128, 4, 64, 0.1, "relu", self_posembed=None, cross_posembed=None, cross_only=False),
            "detr_keys": T.detrDecoder(joint_num=21, num_decoder_layers=2), "satr_keys": T.spatial_aggregate_TR(joint_num=21, num_decoder_layers=2)}
    for name, m in mods.items():
        have = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert have == meta[name], name
        m.load_state_dict(params(meta, name), strict=True)


def test_learned_position_embedding_is_pure_indexing(heads):
    g, meta = heads
    m = T.DetrLearnedPositionEmbedding(16)
    m.load_state_dict(params(meta, "dlearn_keys"))
    with torch.no_grad():
        assert np.array_equal(m(torch.zeros(2, 4, 5, 7)).numpy(), g["dlearn_out"])


def test_no_library_math_in_transfusion_head():
    """The drop-in file launches kernels only: no torch matmul / softmax / LayerNorm call on any forward path."""
    src = open(T.__file__).read()
    for name in ("F.linear", "torch.bmm", "F.softmax", "torch.matmul", "scaled_dot_product", "F.layer_norm", "torch.softmax"):
        assert name not in src, name
