"""Full-frame parity report (GPU box): CUDA render of the complete 512x512 synthetic frame vs the CPU oracle, for
several poses.  Writes gpurun_out/parity_r1.json (copied to profiles/)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instantavatar_b200 import ops
from oracle import render as orender, scene as oscene, testing

rep = {}
for frame in (0, 57):
    sc = testing.oracle_scene(frame)
    scene, _ = testing.upload(sc)
    fr = sc["frame"]
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    t0 = time.time()
    ref = orender.render_test(o, d, near, far, sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1], testing.oracle_model(sc, True))
    t_cpu = time.time() - t0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = ops.render_fwd(scene, t(o), t(d), t(near), t(far), None, 512)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    e_rgb = np.abs(got["rgb"] - ref["rgb"]).max(-1); e_a = np.abs(got["alpha"] - ref["alpha"]); e_d = np.abs(got["depth"] - ref["depth"])
    rep[f"frame{frame}"] = {
        "rays": int(len(o)), "rays_hit": int((ref["counter"] > 0).sum()), "rays_opaque": int((ref["alpha"] > 0.5).sum()),
        "max_abs_err_rgb": float(e_rgb.max()), "max_abs_err_alpha": float(e_a.max()), "max_abs_err_depth": float(e_d.max()),
        "rays_rgb_err_gt_1e-3": int((e_rgb > 1e-3).sum()), "rays_alpha_err_gt_1e-3": int((e_a > 1e-3).sum()),
        "rays_rgb_err_gt_1e-4": int((e_rgb > 1e-4).sum()), "mean_abs_err_rgb_on_hit": float(e_rgb[ref["counter"] > 0].mean()),
        "bit_identical_rays": int(((got["rgb"] == ref["rgb"]).all(-1) & (got["alpha"] == ref["alpha"])).sum()),
        "oracle_seconds_march_only": t_cpu,
    }
    print(frame, rep[f"frame{frame}"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open("gpurun_out/parity_r1.json", "w"), indent=1)
