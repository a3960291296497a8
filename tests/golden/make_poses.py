"""Copies a handful of real SMPL pose frames (inputs only) from the reference's pose tracks into tests/golden/poses.npz:
PeopleSnapshot (BASELINE.json configs 2/4), the NeuMan "seattle" track (config 5, data/custom/seattle/poses/train.npz)
and two frames of the AIST animation (config 3, data/animation/aist_demo.npz, prepared as animate.py:45-54 does: first
frame's translation removed, + (0, 0.15, 5), the training subject's betas).  Run in the build container (needs
/root/reference)."""
import os

import numpy as np

REF = "/root/reference/data"
out = {}
for track, frames in {"male-3-casual": [0, 20, 57, 100], "female-4-casual": [0, 40]}.items():
    z = np.load(f"{REF}/PeopleSnapshot/{track}/poses/anim_nerf_train.npz")
    out[f"{track}/frames"] = np.array(frames)
    out[f"{track}/betas"] = z["betas"]
    for k in ["global_orient", "body_pose", "transl"]:
        out[f"{track}/{k}"] = z[k][frames]
z = np.load(f"{REF}/custom/seattle/poses/train.npz")
frames = [0, 20]
out["seattle/frames"] = np.array(frames)
out["seattle/betas"] = z["betas"]
for k in ["global_orient", "body_pose", "transl"]:
    out[f"seattle/{k}"] = z[k][frames]
z = np.load(f"{REF}/animation/aist_demo.npz")
frames = [40, 200]
thetas = z["poses"][..., :72].astype(np.float32)
transl = (z["trans"] - z["trans"][0:1] + np.array([0, 0.15, 5])).astype(np.float32)  # animate.py:49-50
out["aist_demo/frames"] = np.array(frames)
out["aist_demo/betas"] = out["male-3-casual/betas"]  # animate.py:100: the training subject's shape
out["aist_demo/global_orient"] = thetas[frames, :3]
out["aist_demo/body_pose"] = thetas[frames, 3:]
out["aist_demo/transl"] = transl[frames]
np.savez(os.path.join(os.path.dirname(__file__), "poses.npz"), **out)
print({k: v.shape for k, v in out.items()})
