"""Writes tests/golden/oracle_frames_golden.npz: end-to-end images of the CPU oracle (golden vector 4) for
three poses of the synthetic male-3-casual-shaped scene -- every 4th pixel of the 512x512 demo camera (128x128 rays),
complete pipeline: SMPL -> bone transforms -> skinning field -> 5-pass occupancy grid (seed 42) -> windowed march.
Run from the repository root:  python tests/golden/make_oracle_frames_golden.py   (about two minutes on 8 cores)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import render as orender  # noqa: E402
from oracle import scene as oscene  # noqa: E402
from oracle import testing as scene_util  # noqa: E402

FRAMES = (0, 20, 57)
out = {"frames": np.array(FRAMES), "pixel_index": scene_util.GOLDEN_PIXELS}
for f in FRAMES:
    sc = scene_util.oracle_scene(f)
    fr = sc["frame"]
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    idx = scene_util.GOLDEN_PIXELS
    ref = orender.render_test(o[idx], d[idx], near[idx], far[idx], sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1],
                              scene_util.oracle_model(sc, True))
    for k in ("rgb", "alpha", "depth", "counter"):
        out[f"{f}/{k}"] = np.asarray(ref[k], np.float32)
    out[f"{f}/occ_bits"] = np.packbits(sc["occ"].astype(np.uint8).ravel())
    print(f, "rays hit:", int((ref["alpha"] > 0.5).sum()), "occupied cells:", int(sc["occ"].sum()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_frames_golden.npz"), **out)
