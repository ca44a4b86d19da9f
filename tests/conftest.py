import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "golden_path.npz")))


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_inputs(golden_meta):
    from keypointfusion_b200.utils import synth
    m = golden_meta
    inp = synth.make_inputs(m["B"], m["S"], m["J"], 128, seed=m["seed"])
    for k, v in m["input_sums"].items():  # generator drift guard
        assert abs(float(inp[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    return inp


@pytest.fixture(scope="session")
def path_params(golden_meta):
    """state_dict {block1.*, block2.*} with the SAME deterministic fill the golden run used."""
    from keypointfusion_b200.utils import synth
    sd = {}
    for blk in ("block1.", "block2."):
        for k, shp in golden_meta["Block_KPFusion_keys"].items():
            dt = torch.int64 if k.endswith("num_batches_tracked") or k.endswith("position_ids") else torch.float32
            sd[blk + k] = torch.zeros(shp, dtype=dt)
    synth.fill_state_dict(sd, golden_meta["seed"])
    return sd
