"""CPU: the oracle's crop front end vs the reference's own demo_RGBD.py methods run on the repo's real RGB-D frame
and on synthetic 640x480 frames (fixtures: tests/golden/make_golden_crop.py).  Integer work (bounds, nearest source indices,
paste offsets, z-threshold casts) must be bit-exact; the normalised crop is compared exactly as well (same numpy ops)."""
import os

import numpy as np
import pytest

from oracle import kpf_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_crop.npz")


@pytest.fixture(scope="module")
def gc():
    return dict(np.load(G))


def test_demo_frame(gc):
    c = O.center_from_bbox(gc["box_depth"], gc["box_bbox"])
    assert np.allclose(c, gc["box_center"], rtol=0, atol=1e-9)
    cube = [250, 250, 250]
    d, M, c3 = O.crop_depth(gc["box_depth"], gc["box_center"], cube, gc["box_cam"])
    assert np.array_equal(d, gc["box_crop_d"])
    assert np.allclose(M, gc["box_M"], rtol=0, atol=1e-12) and np.allclose(c3, gc["box_com3d"], rtol=1e-6)
    assert np.array_equal(O.crop_rgb(gc["box_rgb"], gc["box_center"], cube, gc["box_cam"]), gc["box_crop_rgb"])
    assert (d < 1).sum() > 500     # a real hand is in the crop


def test_synthetic_frames_incl_borders(gc):
    cube = [250, 250, 250]
    for i in range(gc["syn_depth"].shape[0]):
        c = O.center_from_bbox(gc["syn_depth"][i], gc["syn_bbox"][i])
        assert np.allclose(c, gc["syn_center"][i], rtol=0, atol=1e-9), i
        d, M, c3 = O.crop_depth(gc["syn_depth"][i], gc["syn_center"][i], cube, gc["syn_cam"])
        assert np.array_equal(d, gc["syn_crop_d"][i]), i
        assert np.allclose(M, gc["syn_M"][i], rtol=0, atol=1e-12) and np.allclose(c3, gc["syn_com3d"][i], rtol=1e-6), i
        assert np.array_equal(O.crop_rgb(gc["syn_rgb"][i], gc["syn_center"][i], cube, gc["syn_cam"]), gc["syn_crop_rgb"][i]), i
