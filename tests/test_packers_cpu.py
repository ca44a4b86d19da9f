"""CPU: host-side packers and schedules the kernels rely on (no device needed).

* the token stack's weight-streaming table (three 32 KB shared-memory slots): simulate the kernel's issue rule and check that no
  slot is overwritten before the GEMM that last read it has completed, and that every GEMM's weights were requested before it runs;
* the point stage's operand column permutation: the packed W1 applied to the kernel's column order equals the reference layers
  applied to the reference order;
* the upload arena: typed views alias one buffer at 256-byte aligned offsets and round-trip their contents."""
import numpy as np
import pytest
import torch

from keypointfusion_b200 import ops
from keypointfusion_b200.runtime import GraphedFusionPath, InputArena
from keypointfusion_b200.utils import synth


def _uncanon(flat, N, K):
    """inverse of ops._canon: [K/8][N][8] -> [N,K]"""
    return flat.reshape(K // 8, N, 8).permute(1, 0, 2).reshape(N, K).float()


@pytest.mark.parametrize("mode", ["init", "final", "encoder"])
def test_token_program_streaming_schedule(path_params, mode):
    p = path_params
    if mode == "init":
        Wfu, bfu = torch.randn(128, 512), torch.randn(128)
        pk = ops.pack_token_program(21, enc=(p, "block1.init_TR."), fusion=(Wfu, bfu))
    elif mode == "final":
        pk = ops.pack_token_program(21, cross=({k[len("block1.crossTR.decoder.0."):]: v for k, v in p.items()
                                                if k.startswith("block1.crossTR.decoder.0.")}, ""), enc=(p, "block1.final_TR."))
    else:
        pk = ops.pack_token_program(21, enc=(p, "block1.final_TR."))
    seq = pk.wseq.tolist()
    G = len(seq)
    assert pk.wseq.dtype == torch.int32 and pk.wseq.shape[1] == 4 and G <= 64
    SLOT = 2048
    total = pk.wmat.numel() // 8
    for g, (src, cnt, dst, after) in enumerate(seq):
        assert 0 <= src and src + cnt <= total and cnt > 0
        assert dst % SLOT == 0 and dst + cnt <= 3 * SLOT, "destination must stay inside the three slots"
        assert -1 <= after < g, "a transfer can only wait for an earlier GEMM"
    # the kernel's rule: after GEMM `done` completes (done = -1 at start), issue transfers in order while after[nxt] <= done
    owner = {}            # slot -> (transfer index that currently owns it)
    issued = set()
    nxt = 0

    def issue_upto(done):
        nonlocal nxt
        while nxt < G and seq[nxt][3] <= done:
            src, cnt, dst, after = seq[nxt]
            for s in range(dst // SLOT, (dst + cnt + SLOT - 1) // SLOT):
                prev = owner.get(s)
                assert prev is None or prev <= done, f"transfer {nxt} overwrites slot {s} still needed by GEMM {prev}"
                owner[s] = nxt
            issued.add(nxt)
            nxt += 1
    issue_upto(-1)
    for g in range(G):
        assert g in issued, f"GEMM {g} would wait for weights that were never requested"
        src, cnt, dst, after = seq[g]
        for s in range(dst // SLOT, (dst + cnt + SLOT - 1) // SLOT):
            assert owner[s] == g, f"GEMM {g} reads slot {s} but it holds transfer {owner[s]}"
        issue_upto(g)     # GEMM g has completed
    assert nxt == G


def test_token_program_kv_tile_is_k_then_v(path_params):
    p = path_params
    pk = ops.pack_token_program(21, enc=(p, "block1.final_TR."))
    seq = pk.wseq.tolist()
    # sequence: embedding, then per layer Q, K|V, O, F1, F2
    src, cnt, dst, _ = seq[2]
    kv = _uncanon(pk.wmat[src * 8:(src + cnt) * 8], 256, 128)
    Wk = p["block1.final_TR.bert.encoder.layer.0.attention.self.key.weight"].bfloat16().float()
    Wv = p["block1.final_TR.bert.encoder.layer.0.attention.self.value.weight"].bfloat16().float()
    assert torch.equal(kv[:128], Wk) and torch.equal(kv[128:], Wv)
    assert cnt == 2 * 2048 and dst == 2048      # one N = 256 tile across slots 1 and 2


def test_point_embed_column_permutation():
    J, C = 21, 128
    g = torch.Generator().manual_seed(3)
    Wf, Wx, Wp, Wr = torch.randn(C, 128, generator=g), torch.randn(C, 3, generator=g), torch.randn(C, 5 * J, generator=g), torch.randn(C, 128, generator=g)
    bf, bx, bp, br = (torch.randn(C, generator=g) for _ in range(4))
    wmat, wvec = ops.pack_point_embed(Wf, bf, Wx, bx, Wp, bp, Wr, br, J)
    W1 = torch.cat([_uncanon(wmat[:128 * 128], 128, 128), _uncanon(wmat[128 * 128:2 * 128 * 128], 128, 128)], 1)   # [128, 256]
    # reference input pieces
    feat, xyz = torch.randn(7, 128, generator=g), torch.randn(7, 3, generator=g)
    wmap, off3, heat = torch.randn(7, J, generator=g), torch.randn(7, 3 * J, generator=g), torch.randn(7, J, generator=g)
    ref = feat @ Wf.T + xyz @ Wx.T + torch.cat([wmap, off3, heat], 1) @ Wp.T + (bf + bx + bp)
    # kernel column order: [feat 128 | wmap J pad to 32 | (ox, oy, oz, heat) per joint | xyz | pad]
    x = torch.zeros(7, 256)
    x[:, :128] = feat
    x[:, 128:128 + J] = wmap
    for j in range(J):
        x[:, 160 + 4 * j:160 + 4 * j + 3] = off3[:, 3 * j:3 * j + 3]
        x[:, 160 + 4 * j + 3] = heat[:, j]
    x[:, 160 + 4 * J:160 + 4 * J + 3] = xyz
    got = x @ W1.T + wvec[:128]
    assert torch.allclose(got, ref, rtol=2e-2, atol=2e-1)     # W1 is stored in bf16
    W1_exact = torch.zeros(C, 256)
    W1_exact[:, :128], W1_exact[:, 128:128 + J] = Wf, Wp[:, :J]
    assert torch.equal(W1[:, :128], Wf.bfloat16().float()) and torch.equal(W1[:, 128:128 + J], Wp[:, :J].bfloat16().float())
    assert torch.equal(wvec[:128], bf + bx + bp) and torch.equal(wvec[128:256], br)


def test_input_arena_views():
    inp = synth.make_inputs(3, 128, 21, 128, seed=5)
    for k in ("img_feat", "img_feat_rgb", "img_offset"):
        inp[k] = inp[k].bfloat16()
    keys = GraphedFusionPath.KEYS
    a = InputArena(inp, keys)
    base = a.buf.data_ptr()
    end = 0
    for k in keys:
        v = a.views[k]
        off = v.data_ptr() - base
        assert off % 256 == 0 and off >= end, k
        assert v.dtype == inp[k].dtype and tuple(v.shape) == tuple(inp[k].shape) and v.is_contiguous()
        end = off + v.numel() * v.element_size()
        v.copy_(inp[k])
    assert end <= a.nbytes
    b = InputArena(inp, keys)
    b.buf.copy_(a.buf)                 # the single-copy upload, on the host
    for k in keys:
        assert torch.equal(b.views[k], inp[k]), k


def test_isclose_bands_as_float_thresholds():
    """K1 tests its pixels against float thresholds instead of the reference's float64 np.isclose bands (backproject.cu:make_bands).
    Re-derive the thresholds the way the kernel does and check, for EVERY float around the bands, that the decisions are the same."""
    f32 = np.float32
    band = 1e-8 + 1e-5 * 1.0

    def bg(v):
        return abs(float(v) - 1.0) <= band

    def na(x, to):
        return np.nextafter(f32(x), f32(to))
    hi = f32(1.0 + band)
    while not bg(hi):
        hi = na(hi, 0)
    while bg(na(hi, 2)):
        hi = na(hi, 2)
    lo = f32(1.0 - band)
    while not bg(lo):
        lo = na(lo, 2)
    while bg(na(lo, 0)):
        lo = na(lo, 0)
    z = f32(1e-8)
    while float(z) > 1e-8:
        z = na(z, 0)
    while float(na(z, 1)) <= 1e-8:
        z = na(z, 1)
    v = f32(1 - 3e-5)
    n = 0
    while v < f32(1 + 3e-5):
        assert bool(np.isclose(float(v), 1.0)) == bg(v) == bool(lo <= v <= hi), v
        v = na(v, 2)
        n += 1
    assert n > 500
    vs = np.arange(0, 4000, dtype=np.int64)
    base = np.array([z], dtype=f32).view(np.int32)[0]
    near = (base + vs - 2000).astype(np.int32).view(f32)          # 4000 consecutive floats around 1e-8
    for sign in (1.0, -1.0):
        d = (near * f32(sign)).astype(f32)
        assert np.array_equal(np.isclose(d.astype(np.float64), 0.0), np.abs(d) <= z)


def test_desa_and_spatial_packers_layout():
    """The offsets the kernels hard-code (desa_fused.cu: Wj at 0, then per scale W1 main 2048 | W1 tail 256 | W2 2048 uint4;
    spatial_agg_tc.cu: Wa main 512 | heat-map part 128 uint4) match what the packers emit, and the contents survive the round trip."""
    g = torch.Generator().manual_seed(11)
    Wj, bj, Wjx, bjx = torch.randn(128, 128, generator=g), torch.randn(128, generator=g), torch.randn(128, 3, generator=g), torch.randn(128, generator=g)
    scales = [(torch.randn(128, 128, generator=g), torch.randn(128, generator=g), torch.randn(128, 3, generator=g), torch.randn(128, generator=g),
               torch.randn(128, 128, generator=g), torch.randn(128, generator=g)) for _ in range(3)]
    wmat, wvec = ops.pack_desa(Wj, bj, Wjx, bjx, scales)
    per_scale = 2048 + 256 + 2048
    assert wmat.dtype == torch.bfloat16 and wmat.numel() == (2048 + 3 * per_scale) * 8
    assert torch.equal(_uncanon(wmat[:2048 * 8], 128, 128), Wj.bfloat16().float())
    for s_, (Wf0, bf0, Wl0, bl0, W2, b2) in enumerate(scales):
        o = (2048 + s_ * per_scale) * 8
        assert torch.equal(_uncanon(wmat[o:o + 2048 * 8], 128, 128), Wf0.bfloat16().float())
        tail = _uncanon(wmat[o + 2048 * 8:o + (2048 + 256) * 8], 128, 16)
        assert torch.equal(tail[:, :3], Wl0.bfloat16().float()) and not tail[:, 3:].any()
        assert torch.equal(_uncanon(wmat[o + (2048 + 256) * 8:o + per_scale * 8], 128, 128), W2.bfloat16().float())
        assert torch.equal(wvec[128 + 512 + s_ * 256:128 + 512 + s_ * 256 + 128], bf0 + bl0)
        assert torch.equal(wvec[128 + 512 + s_ * 256 + 128:128 + 512 + (s_ + 1) * 256], b2)
    assert torch.equal(wvec[:128], bj + bjx) and torch.equal(wvec[128:640].reshape(128, 4)[:, :3], Wjx)
    J = 21
    Wa = torch.randn(J, 128 + J, 1, 1, generator=g)
    wa = ops.pack_spatial_wa(Wa, J)
    assert wa.numel() == (512 + 128) * 8
    main, hm = _uncanon(wa[:512 * 8], 32, 128), _uncanon(wa[512 * 8:], 32, 32)
    assert torch.equal(main[:J], Wa[:, :128, 0, 0].bfloat16().float()) and not main[J:].any()
    assert torch.equal(hm[:J, :J], Wa[:, 128:, 0, 0].bfloat16().float()) and not hm[J:].any() and not hm[:, J:].any()
