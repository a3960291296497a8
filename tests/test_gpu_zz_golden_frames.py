"""GPU: the fused renderer against the COMMITTED end-to-end oracle images of three poses
(tests/golden/oracle_frames_golden.npz; the other parity tests compare with the oracle executed on the spot)."""
import numpy as np
import pytest

from oracle import scene as oscene
from oracle import testing as scene_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("frame", [0, 20, 57])
def test_fused_render_matches_committed_oracle_frames(frame):
    import torch
    from instantavatar_b200 import ops
    gold = scene_util.load_golden_frame(frame)
    sc = scene_util.oracle_scene(frame)
    assert np.array_equal(sc["occ"], gold["occ"])
    scene, _ = scene_util.upload(sc)
    o, d, near, far = oscene.camera_rays(sc["frame"], 512, 512)
    idx = scene_util.GOLDEN_PIXELS
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = ops.render_fwd(scene, t(o[idx]), t(d[idx]), t(near[idx]), t(far[idx]), None, 0, None)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    # the 1e-3 contract, no ray exempted (oracle/testing.py::assert_render_contract)
    scene_util.assert_render_contract(gold, got, allowed_threshold_flips=0, min_hit=600, label=f"frame {frame}")
    miss = gold["counter"] == 0
    assert np.array_equal(got["rgb"][miss], gold["rgb"][miss]) and np.array_equal(got["alpha"][miss], gold["alpha"][miss])
