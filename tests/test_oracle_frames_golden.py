"""CPU: the committed end-to-end oracle images (tests/golden/oracle_frames_golden.npz, SURVEY.md §8c golden vector 4)
are reproduced by the oracle -- a regression pin for the whole CPU pipeline (SMPL, field, occupancy grid, march,
Broyden, hash grid + MLPs, compositing)."""
import numpy as np
import pytest

from oracle import render as orender
from oracle import scene as oscene
from oracle import testing as scene_util


@pytest.mark.parametrize("frame", [0, 57])
def test_oracle_reproduces_committed_frames(frame):
    gold = scene_util.load_golden_frame(frame)
    sc = scene_util.oracle_scene(frame)
    np.testing.assert_array_equal(sc["occ"], gold["occ"])      # same occupancy grid (jitter seed 42)
    fr = sc["frame"]
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    idx = scene_util.GOLDEN_PIXELS
    ref = orender.render_test(o[idx], d[idx], near[idx], far[idx], sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1],
                              scene_util.oracle_model(sc, True))
    assert (gold["alpha"] > 0.5).sum() > 600
    for k in ("rgb", "alpha", "depth"):
        np.testing.assert_allclose(ref[k], gold[k], atol=1e-5, err_msg=k)
    np.testing.assert_array_equal(ref["counter"], gold["counter"])


def test_golden_frames_are_distinct_poses():
    a, b, c = (scene_util.load_golden_frame(f) for f in (0, 20, 57))
    assert a["rgb"].shape == (128 * 128, 3) and a["alpha"].shape == (128 * 128,)
    assert np.abs(a["alpha"] - b["alpha"]).max() > 0.5 and np.abs(b["alpha"] - c["alpha"]).max() > 0.5
    for g in (a, b, c):
        assert np.all((g["alpha"] >= -1e-6) & (g["alpha"] <= 1 + 1e-6))
        miss = g["counter"] == 0
        assert np.all(g["rgb"][miss] == 1.0) and np.all(g["alpha"][miss] == 0.0)   # untouched rays: white, transparent
