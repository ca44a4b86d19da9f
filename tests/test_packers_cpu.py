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
    """inverse of ops._canon16: [K/8][N][8] 16-bit -> [N,K] f32"""
    return flat.reshape(K // 8, N, 8).permute(1, 0, 2).reshape(N, K).float()


def _unplanes(flat_i16, N, K, fmt=None):
    """inverse of ops._canon: canonical hi plane followed by canonical lo plane (raw int16 bits) -> [N,K] f32 = hi + lo"""
    dt = ops._fmt_dtype(ops.SPLIT_FMT if fmt is None else fmt)
    n = N * K
    assert flat_i16.dtype == torch.int16 and flat_i16.numel() == 2 * n
    return _uncanon(flat_i16[:n].view(dt), N, K) + _uncanon(flat_i16[n:].view(dt), N, K)


def _split_tol(fmt=None):
    return 2.0 ** -20 if (ops.SPLIT_FMT if fmt is None else fmt) == ops.FMT_F16 else 2.0 ** -15


@pytest.mark.parametrize("fmt", [ops.FMT_F16, ops.FMT_BF16])
def test_split_planes_reconstruct(fmt):
    """hi + lo reproduces an fp32 tensor to ~2^-22 (fp16 planes; plus the 2^-25 absolute floor of fp16 subnormals, which the lo
    plane of anything below ~0.1 falls into) / ~2^-16 (bf16 planes, full fp32 range) relative; fp16 refuses out-of-range weights."""
    g = torch.Generator().manual_seed(1)
    W = torch.randn(128, 128, generator=g) * torch.logspace(-2, 2, 128)[:, None]
    hi, lo = ops.split_planes(W, fmt)
    err = ((hi.float() + lo.float()) - W).abs()
    floor = 2.0 ** -24 if fmt == ops.FMT_F16 else 0.0
    assert bool((err <= _split_tol(fmt) * W.abs() + floor).all())
    if fmt == ops.FMT_F16:
        with pytest.raises(ValueError, match="KPF_SPLIT_FMT"):
            ops.split_planes(W * 1e4, fmt)


@pytest.mark.parametrize("mode", ["init", "final", "encoder", "prologue"])
def test_token_program_ring_schedule(path_params, mode):
    """The weight ring of csrc/token_stack.cu: entries in consumption order, each <= one 32 KB slot; per-layer vectors are requested
    in front of the right entries (layer 0 at entry 0, layer it+1 at layer it's V weights, i.e. after the four-deep ring has
    forced layer it-1 -- the previous user of that vector buffer -- to finish)."""
    p = path_params
    Wfu, bfu = torch.randn(128, 512), torch.randn(128)
    if mode == "init":
        pk = ops.pack_token_program(21, enc=(p, "block1.init_TR."), fusion=(Wfu, bfu))
        expect = 8 + 2 + 4 * 9
    elif mode == "final":
        pk = ops.pack_token_program(21, cross=({k[len("block1.crossTR.decoder.0."):]: v for k, v in p.items()
                                                if k.startswith("block1.crossTR.decoder.0.")}, ""), enc=(p, "block1.final_TR."))
        expect = 12 + 3 + 4 * 9
    elif mode == "encoder":
        pk = ops.pack_token_program(21, enc=(p, "block1.final_TR."))
        expect = 3 + 4 * 9
    else:
        pk = ops.pack_token_program(21, fusion=(Wfu, bfu))
        expect = 8
    seq = pk.wseq.tolist()
    G = len(seq)
    assert pk.wseq.dtype == torch.int32 and pk.wseq.shape[1] == 4 and G == expect <= 64 and pk.n_weights == G
    assert pk.wmat.dtype == torch.int16
    total = pk.wmat.numel() // 8
    off = 0
    for e, (src, cnt, vec, _) in enumerate(seq):
        assert src == off and 0 < cnt <= 2048, "entries are contiguous, in order, at most one ring slot"
        off += cnt
    assert off == total
    n_layers = pk.cross + pk.L
    vec_at = {vec: e for e, (_, _, vec, _) in enumerate(seq) if vec >= 0}
    assert sorted(vec_at) == list(range(n_layers))
    if n_layers:
        assert vec_at[0] == 0
        for it in range(1, n_layers):
            # requested at entry base(it-1) + 4: the producer gets there only once base(it-1) has been consumed (ring of four),
            # i.e. layer it-2 (the previous user of buffer it & 1) has finished; and strictly before layer it's own first entry
            assert vec_at[it] > vec_at[it - 1] and vec_at[it] - 4 >= (vec_at[it - 1] - 4 if it > 1 else 0)


def test_token_program_tiles_reconstruct_weights(path_params):
    """A [128,128] matrix = two half-K ring entries, each canonical hi plane + lo plane: reassembling them gives the fp32 weight."""
    p = path_params
    pk = ops.pack_token_program(21, enc=(p, "block1.init_TR."))
    seq = pk.wseq.tolist()
    # entries: embedding (2), then per layer K, Q, V, O (2 each), FFN (1)
    for name, e0 in (("key", 2), ("query", 4), ("value", 6)):
        W = p[f"block1.init_TR.bert.encoder.layer.0.attention.self.{name}.weight"]
        halves = []
        for h in range(2):
            src, cnt, _, _ = seq[e0 + h]
            assert cnt == 2048
            halves.append(_unplanes(pk.wmat[src * 8:(src + cnt) * 8], 128, 64))
        got = torch.cat(halves, 1)
        assert float((got - W).abs().max()) <= _split_tol() * float(W.abs().max())
    src, cnt, _, _ = seq[10]      # layer 0's FFN entry: W1 [16,128] hi | lo, W2 [128,16] hi | lo
    assert cnt == 1024
    blob = pk.wmat[src * 8:(src + cnt) * 8]
    W1 = p["block1.init_TR.bert.encoder.layer.0.intermediate.dense.weight"]
    W2 = p["block1.init_TR.bert.encoder.layer.0.output.dense.weight"]
    assert float((_unplanes(blob[:2 * 16 * 128], 16, 128) - W1).abs().max()) <= _split_tol() * float(W1.abs().max())
    assert float((_unplanes(blob[2 * 16 * 128:], 128, 16) - W2).abs().max()) <= _split_tol() * float(W2.abs().max())


def test_point_embed_column_permutation():
    J, C = 21, 128
    g = torch.Generator().manual_seed(3)
    Wf, Wx, Wp, Wr = torch.randn(C, 128, generator=g), torch.randn(C, 3, generator=g), torch.randn(C, 5 * J, generator=g), torch.randn(C, 128, generator=g)
    bf, bx, bp, br = (torch.randn(C, generator=g) for _ in range(4))
    wmat, wvec = ops.pack_point_embed(Wf, bf, Wx, bx, Wp, bp, Wr, br, J)
    assert wmat.dtype == torch.int16 and wmat.numel() == 2 * 128 * 256 + 2 * 128 * 128   # W1 hi | lo, W2 hi | lo (point_embed.cu offsets)
    W1 = _unplanes(wmat[:2 * 128 * 256], 128, 256)                                      # [128, 256]
    W2 = _unplanes(wmat[2 * 128 * 256:], 128, 128)
    assert float((W2 - Wr).abs().max()) <= _split_tol() * float(Wr.abs().max())
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
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-3)
    tol = _split_tol() * float(max(Wf.abs().max(), Wp.abs().max()))
    assert float((W1[:, :128] - Wf).abs().max()) <= tol and float((W1[:, 128:128 + J] - Wp[:, :J]).abs().max()) <= tol
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
    """The offsets the kernels hard-code (desa_fused.cu: Wj hi | lo at 0, then per scale W1 main hi | lo (2 x 2048), W1 tail hi | lo
    (2 x 256), W2 hi | lo (2 x 2048) uint4; spatial_agg_tc.cu: Wa main 3 bf16 planes (3 x 512), heat-map part hi | lo (2 x 128) uint4) match what
    the packers emit, and the contents survive the round trip."""
    g = torch.Generator().manual_seed(11)
    Wj, bj, Wjx, bjx = torch.randn(128, 128, generator=g), torch.randn(128, generator=g), torch.randn(128, 3, generator=g), torch.randn(128, generator=g)
    scales = [(torch.randn(128, 128, generator=g), torch.randn(128, generator=g), torch.randn(128, 3, generator=g), torch.randn(128, generator=g),
               torch.randn(128, 128, generator=g), torch.randn(128, generator=g)) for _ in range(3)]
    wmat, wvec = ops.pack_desa(Wj, bj, Wjx, bjx, scales)
    per_scale = 2 * (2048 + 256 + 2048)
    tol = _split_tol() * 6.0
    assert wmat.dtype == torch.int16 and wmat.numel() == (4096 + 3 * per_scale) * 8
    assert float((_unplanes(wmat[:4096 * 8], 128, 128) - Wj).abs().max()) <= tol
    for s_, (Wf0, bf0, Wl0, bl0, W2, b2) in enumerate(scales):
        o = (4096 + s_ * per_scale) * 8
        assert float((_unplanes(wmat[o:o + 4096 * 8], 128, 128) - Wf0).abs().max()) <= tol
        tail = _unplanes(wmat[o + 4096 * 8:o + (4096 + 512) * 8], 128, 16)
        assert float((tail[:, :3] - Wl0).abs().max()) <= tol and not tail[:, 3:].any()
        assert float((_unplanes(wmat[o + (4096 + 512) * 8:o + per_scale * 8], 128, 128) - W2).abs().max()) <= tol
        assert torch.equal(wvec[128 + 512 + s_ * 256:128 + 512 + s_ * 256 + 128], bf0 + bl0)
        assert torch.equal(wvec[128 + 512 + s_ * 256 + 128:128 + 512 + (s_ + 1) * 256], b2)
    assert torch.equal(wvec[:128], bj + bjx) and torch.equal(wvec[128:640].reshape(128, 4)[:, :3], Wjx)
    J = 21
    Wa = torch.randn(J, 128 + J, 1, 1, generator=g)
    wa = ops.pack_spatial_wa(Wa, J)
    assert wa.dtype == torch.int16 and wa.numel() == (3 * 512 + 2 * 128) * 8     # feature part: 3 bf16 planes; heat-map part: hi | lo
    main = sum(_uncanon(wa[i * 512 * 8:(i + 1) * 512 * 8].view(torch.bfloat16), 32, 128) for i in range(3))
    hm = _unplanes(wa[1536 * 8:], 32, 32)
    assert float((main[:J] - Wa[:, :128, 0, 0]).abs().max()) <= 2.0 ** -22 * 6.0 and not main[J:].any()
    assert float((hm[:J, :J] - Wa[:, 128:, 0, 0]).abs().max()) <= tol and not hm[J:].any() and not hm[:, J:].any()


def test_kernel_cache_follows_parent_level_loads(path_params):
    """a top-level net.load_state_dict(ckpt) recurses through _load_from_state_dict and never calls the children's
    load_state_dict; the packed-weight caches must still be rebuilt.  (CPU-only: packing needs no device.)"""
    from keypointfusion_b200.model.model import KPFusion
    net = KPFusion(joint_num=21).eval()
    net.load_state_dict(path_params)
    k1 = net.block1.kc()
    assert net.block1.kc() is k1                       # unchanged weights: cached
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd["block1.pcl_feat_emb_RGB.0.weight"] += 1.0
    net.load_state_dict(sd)                            # parent-level load
    k2 = net.block1.kc()
    assert k2 is not k1 and not torch.equal(k1["pe_wmat"], k2["pe_wmat"])
    with torch.no_grad():
        net.block1.atten_spatial.weight.mul_(2.0)      # in-place edit
    assert net.block1.kc() is not k2
    layer = net.block1.crossTR.decoder[-1]
    t1 = layer.packed_tc(21)
    assert layer.packed_tc(21) is t1
    with torch.no_grad():
        layer.linear1.weight.add_(1.0)
    assert layer.packed_tc(21) is not t1


def test_state_dict_accepts_transformers_4_25_position_ids(path_params):
    """Released checkpoints (transformers 4.25.1) carry `...bert.embeddings.position_ids`; newer key lists do not.  Both load strictly."""
    from keypointfusion_b200.model.model import KPFusion
    net = KPFusion(joint_num=21)
    sd = {k: v for k, v in path_params.items() if not k.endswith("position_ids")}
    net.load_state_dict(sd, strict=True)
    sd_old = dict(sd)
    for blk in ("block1", "block2"):
        for tr in ("init_TR", "final_TR"):
            sd_old[f"{blk}.{tr}.bert.embeddings.position_ids"] = torch.arange(512).expand(1, -1)
    net.load_state_dict(sd_old, strict=True)
    assert any(k.endswith("bert.embeddings.position_ids") for k in net.state_dict())


def test_modules_refuse_training_mode(path_params):
    """Inference only: a drop-in left in .train() (or fed tensors that require grad) raises instead of silently dropping gradients."""
    from keypointfusion_b200.model.fusion_layer import RGBDFusion
    from keypointfusion_b200.model.model import KPFusion
    m = RGBDFusion(8, 8)
    x = [torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4)]
    with pytest.raises(RuntimeError, match="inference only"):
        m(x)
    net = KPFusion(joint_num=21)
    with pytest.raises(RuntimeError, match="inference only"):
        net.block1.init_TR(torch.zeros(1, 21, 128))
    net.eval()
    with pytest.raises(RuntimeError, match="requires grad"):
        net.block1.init_TR(torch.zeros(1, 21, 128, requires_grad=True))
