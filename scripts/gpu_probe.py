"""Scratch GPU probe: times the fused renderer on the oracle-built 512x512 synthetic frame."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instantavatar_b200 import ops
from oracle import scene as oscene
from oracle import testing as scene_util

sc = scene_util.oracle_scene(0)
scene, _ = scene_util.upload(sc)
fr = sc["frame"]
o, d, near, far = oscene.camera_rays(fr, 512, 512)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
o, d, near, far = t(o), t(d), t(near), t(far)
res = {}
for rpw, plan in ((8, 0), (8, 1), (4, 0), (4, 1), (2, 1)):
  ops.set_option("render_rays_per_warp", rpw); ops.set_option("render_plan", plan)
  for width in (512, 0):
    stats = ops.new_stats("cuda")
    out = ops.render_fwd(scene, o, d, near, far, None, width, stats)
    torch.cuda.synchronize()
    st = ops.stats_dict(stats)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(20):
        ev0.record(); ops.render_fwd(scene, o, d, near, far, None, width, None, out); ev1.record(); torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    key = f"rpw{rpw}_plan{plan}_width{width}"
    res[key] = {"ms_median": float(np.median(ts)), "ms_min": float(min(ts)), "stats": st,
                "rays_per_s": 262144 / (np.median(ts) * 1e-3), "alpha_sum": float(out["alpha"].sum())}
    print(key, res[key])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
