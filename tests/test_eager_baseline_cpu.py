"""CPU: the eager-PyTorch baseline that bench.py times on the GPU (baseline/eager_torch_path.py) computes the same function as the
oracle (so its time is the time of the reference's algorithm, not of something cheaper)."""
import numpy as np
import torch

from baseline import eager_torch_path as E
from keypointfusion_b200.utils import synth
from oracle import kpf_oracle as O


def test_eager_path_matches_oracle(path_params):
    B = 2
    inp = synth.make_inputs(B, 128, 21, 128, seed=7)
    g = [inp[k].numpy() for k in ("center", "M", "cube", "cam")]
    pcl = np.stack([O.getpcl_sample(inp["img"][b, 0].numpy(), g[0][b], g[2][b], g[1][b], g[3][b], seed=1, b=b)[0] for b in range(B)])
    pcl = torch.from_numpy(pcl)
    ores, osw, _ = O.fusion_path(path_params, inp["img"], pcl, inp["img_offset"], inp["img_feat"], inp["img_feat_rgb"], *g)
    res, sw = E.fusion_path(path_params, inp["img"], pcl, inp["img_offset"], inp["img_feat"], inp["img_feat_rgb"], inp["center"], inp["M"],
                            inp["cube"], inp["cam"])
    for a, b in zip(res, ores):
        assert float(np.linalg.norm((a - b).numpy() * 125.0, axis=-1).mean()) < 0.05     # mm
    for a, b in zip(sw, osw):
        assert float((a - b).norm() / b.norm()) < 1e-3
