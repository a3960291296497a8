"""Scratch: single configuration render for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instantavatar_b200 import ops
from oracle import scene as oscene
from oracle import testing as scene_util
rpw = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = scene_util.oracle_scene(0)
scene, _ = scene_util.upload(sc)
o, d, near, far = oscene.camera_rays(sc["frame"], 512, 512)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
o, d, near, far = t(o), t(d), t(near), t(far)
ops.set_option("render_rays_per_warp", rpw)
for i in range(6):
    out = ops.render_fwd(scene, o, d, near, far, None, 512)
torch.cuda.synchronize()
print("done", float(out["alpha"].sum()))
