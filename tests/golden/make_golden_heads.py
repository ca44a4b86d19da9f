"""Golden vectors for the exports of model/transfusion_head.py that KPFusion itself does not run: the three
position-embedding classes, MultiheadAttention for general shapes with masks, TransformerDecoderLayer with self-attention,
detrDecoder and spatial_aggregate_TR.  Runs the UNMODIFIED reference on CPU through ref_shims.py (build container only) and
writes golden_heads.npz + golden_heads_meta.json, which are committed.

    python tests/golden/make_golden_heads.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

ref_shims.install()
from keypointfusion_b200.utils import synth  # noqa: E402
from model import transfusion_head as T  # noqa: E402

torch.set_grad_enabled(False)
SEED = 11


def npy(t):
    return t.detach().cpu().numpy()


def main():
    rs = np.random.RandomState(SEED)
    rnd = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    out, meta = {}, {"seed": SEED}

    # ---- position embeddings (:16-91)
    pe = T.PositionEmbeddingLearned(3, 32).eval()
    synth.fill_state_dict(pe, SEED)
    out["pel_xyz"] = npy(rnd(2, 50, 3))
    out["pel_out"] = npy(pe(torch.from_numpy(out["pel_xyz"])))
    meta["pel_keys"] = {k: list(v.shape) for k, v in pe.state_dict().items()}
    dl = T.DetrLearnedPositionEmbedding(16).eval()
    synth.fill_state_dict(dl, SEED)
    out["dlearn_out"] = npy(dl(torch.zeros(2, 4, 5, 7)))
    meta["dlearn_keys"] = {k: list(v.shape) for k, v in dl.state_dict().items()}
    mask = torch.ones(2, 6, 5)
    mask[0, 4:, :] = 0          # padded rows / columns, as DETR batches have them
    mask[1, :, 3:] = 0
    out["sine_mask"] = npy(mask)
    out["sine_norm"] = npy(T.DetrSinePositionEmbedding(16, normalize=True)(torch.zeros(2, 4, 6, 5), mask))
    out["sine_raw"] = npy(T.DetrSinePositionEmbedding(8, temperature=100)(torch.zeros(2, 4, 6, 5), mask))
    out["sine_ones"] = npy(T.DetrSinePositionEmbedding(64, normalize=True)(torch.zeros(1, 4, 10, 12), torch.ones(1, 10, 12)))

    # ---- MultiheadAttention, general shapes + masks (:176-300, :303-556)
    for tag, (E, H, L, S) in {"mha_a": (64, 4, 7, 45), "mha_b": (128, 2, 33, 70)}.items():
        m = T.MultiheadAttention(E, H).eval()
        synth.fill_state_dict(m, SEED)
        q, k, v = rnd(L, 2, E), rnd(S, 2, E), rnd(S, 2, E)
        am = rnd(L, S)
        kpm = torch.zeros(2, S, dtype=torch.bool)
        kpm[0, S - 5:] = True
        kpm[1, ::7] = True
        o, w = m(q, k, v, key_padding_mask=kpm, need_weights=True, attn_mask=am)
        o2, _ = m(q, k, k, need_weights=False)
        for n, t in (("q", q), ("k", k), ("v", v), ("am", am), ("kpm", kpm), ("out", o), ("w", w), ("out_kk", o2)):
            out[f"{tag}_{n}"] = npy(t)

    # ---- TransformerDecoderLayer with self-attention and tensor position embeddings (:94-173)
    lay = T.TransformerDecoderLayer(128, 4, 64, 0.1, "relu", self_posembed=None, cross_posembed=None, cross_only=False).eval()
    synth.fill_state_dict(lay, SEED)
    q, k, qp, kp = rnd(2, 10, 128), rnd(2, 37, 128), 0.3 * rnd(2, 10, 128), 0.3 * rnd(2, 37, 128)
    out["lay_q"], out["lay_k"], out["lay_qp"], out["lay_kp"] = npy(q), npy(k), npy(qp), npy(kp)
    out["lay_out"] = npy(lay(q, k, qp, kp))
    meta["lay_keys"] = {k_: list(v.shape) for k_, v in lay.state_dict().items()}

    # ---- detrDecoder (:560-632) and spatial_aggregate_TR (:711-783) on a non-square 10 x 12 map
    anchors, img = rnd(2, 21, 128), rnd(2, 128, 10, 12)
    out["dec_anchor"], out["dec_img"] = npy(anchors), npy(img)
    for name, cls, args in (("detr", T.detrDecoder, (anchors, img)), ("satr", T.spatial_aggregate_TR, (img, anchors))):
        d = cls(joint_num=21, hidden_channel=128, num_heads=4, ffn_channel=128, dropout=0.1, num_decoder_layers=2).eval()
        synth.fill_state_dict(d, SEED)
        out[f"{name}_out"] = npy(d(*args))
        meta[f"{name}_keys"] = {k_: list(v.shape) for k_, v in d.state_dict().items()}

    np.savez_compressed(os.path.join(HERE, "golden_heads.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "golden_heads_meta.json"), "w"), indent=0, sort_keys=True)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
