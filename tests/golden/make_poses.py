"""Copies a handful of real SMPL pose frames (inputs only) from the reference's pose tracks
into tests/golden/poses.npz.  Run in the build container (needs /root/reference)."""
import numpy as np, os
REF = "/root/reference/data/PeopleSnapshot"
out = {}
for track, frames in {"male-3-casual": [0, 20, 57, 100], "female-4-casual": [0, 40]}.items():
    z = np.load(f"{REF}/{track}/poses/anim_nerf_train.npz")
    out[f"{track}/frames"] = np.array(frames)
    out[f"{track}/betas"] = z["betas"]
    for k in ["global_orient", "body_pose", "transl"]:
        out[f"{track}/{k}"] = z[k][frames]
np.savez(os.path.join(os.path.dirname(__file__), "poses.npz"), **out)
print({k: v.shape for k, v in out.items()})
