"""CPU: the property behind the DESA tile kernel's narrow groups (csrc/desa_fused.cu): pointnet2's ball query pads a ball of fewer than
nsample hits with copies of its first hit (model.py:158, :174; oracle.ball_query), and DESA max-pools over the group (model.py:197-198),
so grouping only the first W >= (widest ball) indices per joint leaves the output unchanged, bit for bit.  Checked on the oracle."""
import numpy as np
import torch

from keypointfusion_b200.utils import synth
from oracle import kpf_oracle as O


def _cloud(B, seed):
    inp = synth.make_inputs(B, 128, 21, 128, seed=seed)
    g = [inp[k].numpy() for k in ("center", "M", "cube", "cam")]
    pcl = np.stack([O.getpcl_sample(inp["img"][b, 0].numpy(), g[0][b], g[2][b], g[1][b], g[3][b], seed=2, b=b)[0] for b in range(B)])
    return torch.from_numpy(pcl.astype(np.float32))


def test_ball_query_prefix_and_padding():
    """The first W slots of an nsample-wide query ARE the W-wide query whenever the ball holds <= W points."""
    pcl = _cloud(2, 5)
    joint = pcl[:, ::48][:, :21].contiguous() + 0.01
    xyz = torch.cat([pcl, joint], 1).numpy()
    for r in (0.06, 0.1, 0.2):
        wide, cnt = O.ball_query(xyz, joint.numpy(), r, 64, return_counts=True)
        for W in (16, 32):
            ok = cnt <= W
            assert ok.any()
            narrow = O.ball_query(xyz, joint.numpy(), r, W)
            assert np.array_equal(wide[ok][:, :W], narrow[ok])
            assert (wide[ok][:, W:] == wide[ok][:, :1]).all()      # everything behind the ball's population is the first hit


def test_desa_invariant_to_group_width(path_params):
    B = 2
    pcl = _cloud(B, 6)
    joint = pcl[:, ::48][:, :21].contiguous() + 0.01
    gen = torch.Generator().manual_seed(0)
    feat, jf = torch.randn(B, pcl.shape[1], 128, generator=gen), torch.randn(B, 21, 128, generator=gen)
    radius = (0.05, 0.1, 0.12)
    xyz = torch.cat([pcl, joint], 1).numpy()
    widest = [int(O.ball_query(xyz, joint.numpy(), r, 64, return_counts=True)[1].max()) for r in radius]
    widths = tuple(16 if w <= 16 else 32 if w <= 32 else 64 for w in widest)
    assert min(widths) < 64, widest                               # at least one scale is actually narrowed
    ref = O.desa(path_params, "block1.FA.", feat, jf, pcl, joint, radius=radius, nsample=(64, 64, 64))
    out = O.desa(path_params, "block1.FA.", feat, jf, pcl, joint, radius=radius, nsample=widths)
    assert torch.equal(out, ref), (widest, widths)
