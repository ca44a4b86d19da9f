"""Seeded synthetic inputs and deterministic parameter filling.

numpy `RandomState` only (bit-stable across numpy versions), so the same tensors can be fed to the
reference (when the golden fixtures are generated), to the oracle and to the CUDA path.
"""
import zlib

import numpy as np
import torch


def make_depth_crops(B, S=128, seed=0, noise=0.35):
    """`img [B,1,S,S]` f32: background exactly 1.0, hand = disc of radius ~0.3125*S holding a smooth surface plus U(-noise, noise)
    per-pixel noise (normalised depth: 1.0 = half the crop cube, 125 mm), plus planted edge cases (isclose band, 0.99 threshold).
    noise = 0.35 (+-44 mm, the default the parity tests use) scatters a point's 3-D nearest cells over many feature-map cells --
    a worst case for every locality heuristic; a depth sensor is closer to noise = 0.02 (+-2.5 mm)."""
    rs = np.random.RandomState(seed)
    img = np.ones((B, 1, S, S), np.float32)
    yy, xx = np.mgrid[0:S, 0:S]
    for b in range(B):
        cx = S / 2 + rs.uniform(-0.06, 0.06) * S
        cy = S / 2 + rs.uniform(-0.06, 0.06) * S
        rad = (0.3125 + rs.uniform(-0.08, 0.04)) * S
        disc = (xx + 0.5 - cx) ** 2 + (yy + 0.5 - cy) ** 2 <= rad ** 2
        # smooth-ish surface + noise so neighbouring cells are close in 3-D, like a real hand
        base = 0.25 * np.sin((xx + 3 * b) / S * 5.0) * np.cos(yy / S * 4.0)
        vals = (base + rs.uniform(-noise, noise, size=(S, S))).astype(np.float32)
        img[b, 0][disc] = vals[disc]
        # planted edge cases inside the disc
        c = S // 2
        img[b, 0, c, c] = np.float32(1.0 - 5e-6)      # inside isclose(.,1) band -> background
        img[b, 0, c, c + 1] = np.float32(1.0 - 2e-5)  # just outside the band -> valid point
        img[b, 0, c + 1, c] = np.float32(0.99)        # exactly the feature threshold
        img[b, 0, c + 4, c + 4] = np.float32(0.995)   # foreground for getpcl, background for features
    return img


def make_camera(B, S=128, seed=0, cube_mm=250.0):
    """center [B,3] mm, M [B,3,3] (full-image px -> crop px, with a small rotation), cube [B,3], cam [B,4]."""
    rs = np.random.RandomState(seed + 1000)
    cam = np.tile(np.array([617.0, 617.0, 312.0, 241.0], np.float32), (B, 1))
    center = np.stack([20 + rs.uniform(-40, 40, B), -30 + rs.uniform(-40, 40, B),
                       600 + rs.uniform(-150, 200, B)], 1).astype(np.float32)
    cube = np.full((B, 3), cube_mm, np.float32)
    M = np.zeros((B, 3, 3), np.float32)
    for b in range(B):
        fx, fy, fu, fv = cam[b]
        uc = center[b, 0] * fx / center[b, 2] + fu
        vc = center[b, 1] * fy / center[b, 2] + fv
        s = S / (cube_mm * fx / center[b, 2]) * (1 + rs.uniform(-0.05, 0.05))
        th = np.deg2rad(rs.uniform(-10, 10))
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) * s
        t = np.array([S / 2, S / 2]) - R @ np.array([uc, vc])
        M[b] = np.array([[R[0, 0], R[0, 1], t[0]], [R[1, 0], R[1, 1], t[1]], [0, 0, 1]], np.float32)
    return center, M, cube, cam


def make_feature_maps(B, J=21, C=128, H=32, seed=0):
    rs = np.random.RandomState(seed + 2000)
    img_feat = rs.standard_normal((B, C, H, H)).astype(np.float32)
    img_feat_rgb = rs.standard_normal((B, C, H, H)).astype(np.float32)
    img_offset = rs.standard_normal((B, 5 * J, H, H)).astype(np.float32)
    return img_feat, img_feat_rgb, img_offset


def make_inputs(B, S=128, J=21, C=128, seed=0, as_torch=True, bf16_round=False, depth_noise=0.35):
    """Everything the fusion path consumes (fusion-path-only configs feed feature maps directly)."""
    H = S // 4
    img = make_depth_crops(B, S, seed, depth_noise)
    center, M, cube, cam = make_camera(B, S, seed)
    img_feat, img_feat_rgb, img_offset = make_feature_maps(B, J, C, H, seed)
    rs = np.random.RandomState(seed + 3000)
    out = dict(img=img, img_rgb=rs.uniform(0, 1, (B, 3, S, S)).astype(np.float32), center=center, M=M, cube=cube,
               cam=cam, img_feat=img_feat, img_feat_rgb=img_feat_rgb, img_offset=img_offset)
    if as_torch:
        out = {k: torch.from_numpy(v) for k, v in out.items()}
        if bf16_round:
            for k in ("img_feat", "img_feat_rgb", "img_offset"):
                out[k] = out[k].bfloat16().float()
    return out


def _key_rng(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def fill_state_dict(module_or_sd, seed=0):
    """Deterministically overwrite every tensor of a state_dict from its KEY NAME (order independent), so a
    reference module and its drop-in get identical parameters without shipping checkpoints.  Scales keep
    activations O(1): fan-in scaled normals for matrices, U(0.8,1.2) norm scales, N(0,0.05) biases, randomised
    BatchNorm running statistics (eval mode)."""
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    with torch.no_grad():
        for key in sorted(sd.keys()):
            t = sd[key]
            if key.endswith("num_batches_tracked") or not t.is_floating_point():
                continue
            rs = _key_rng(key, seed)
            shape = tuple(t.shape)
            if key.endswith("running_var"):
                v = rs.uniform(0.5, 1.5, shape)
            elif key.endswith("running_mean"):
                v = rs.standard_normal(shape) * 0.1
            elif key.endswith("weight_dis"):
                v = np.full(shape, 0.3)
            elif key.endswith("bias"):
                v = rs.standard_normal(shape) * 0.05
            elif t.dim() == 1:  # norm scale
                v = rs.uniform(0.8, 1.2, shape)
            elif "embeddings" in key or "posembed" in key:  # embedding tables
                v = rs.standard_normal(shape) * 0.1
            else:
                fan_in = int(np.prod(shape[1:]))
                v = rs.standard_normal(shape) / np.sqrt(fan_in)
            t.copy_(torch.from_numpy(np.asarray(v, np.float32)).to(t.dtype))
    return sd


def explicit_ranks(P, n=1024, seed=0):
    """Seeded explicit selection for a3 (ranks into the ordered valid-pixel list), multiset semantics of
    loader.py:1179-1185, so reference, oracle and kernel can be fed the SAME selection."""
    rs = np.random.RandomState(seed + 4000)
    if P == 0:
        return np.zeros(n, np.int32)
    if P >= n:
        return rs.permutation(P)[:n].astype(np.int32)
    tmp = n // P
    pool = np.concatenate([np.repeat(np.arange(P), tmp), rs.permutation(P)[: n - tmp * P]])
    return pool[rs.permutation(n)].astype(np.int32)
