"""How far could real tiny-cuda-nn be from the rounding model the product implements?  

Renders the committed golden ray set (every 4th pixel of the 512x512 demo camera, three poses) with the CPU oracle in
rounding mode 1 (fp16 values, fp32 accumulation -- the product) and mode 2 (tiny-cuda-nn-like fp16 accumulation,
oracle/ia_oracle.c), each with its own end-to-end pipeline (own occupancy grid), and writes the image-level deviations to
profiles/parity_r2.json next to the network-level deltas of tests/golden/ngp_kat_golden.npz.  CPU only, ~2 minutes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import render as orender  # noqa: E402
from oracle import scene as oscene  # noqa: E402
from oracle import testing as scene_util  # noqa: E402

report = {"what": "oracle mode 2 (tcnn-like fp16 accumulation) vs mode 1 (product model), golden ray set 128x128 of 512x512", "frames": {}}
subj = oscene.build_subject()
for f in (0, 20, 57):
    from instantavatar_b200 import synthetic
    fr = subj.prepare_frame(synthetic.load_pose(f))
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    idx = scene_util.GOLDEN_PIXELS
    imgs, occs = {}, {}
    for mode in (1, 2):
        net = oscene.build_net(subj, emulate=mode)
        occ, _, _ = oscene.build_occupancy(subj, fr, net)
        occs[mode] = occ
        imgs[mode] = orender.render_test(o[idx], d[idx], near[idx], far[idx], occ, fr["bbox_deformed"][0], fr["bbox_deformed"][1],
                                         lambda p: orender.deform_query(p, fr, subj, net, True))
    drgb = np.abs(imgs[2]["rgb"] - imgs[1]["rgb"]).max(-1)
    dal = np.abs(imgs[2]["alpha"] - imgs[1]["alpha"])
    report["frames"][str(f)] = {
        "rays": int(len(idx)), "rays_hit": int((imgs[1]["alpha"] > 0.5).sum()),
        "occupancy_cells_differing": int((occs[1] != occs[2]).sum()),
        "rgb_linf": float(drgb.max()), "alpha_linf": float(dal.max()),
        "rays_rgb_gt_1e-3": int((drgb > 1e-3).sum()), "rays_rgb_gt_1e-2": int((drgb > 1e-2).sum()),
        "rgb_p99": float(np.quantile(drgb, 0.99)), "rgb_median_on_hit": float(np.median(drgb[imgs[1]["alpha"] > 0.5])),
    }
    print(f, report["frames"][str(f)])
z = np.load(os.path.join(ROOT, "tests", "golden", "ngp_kat_golden.npz"))
report["network_level_65536_points"] = {k: dict(zip(["sigma_linf", "sigma_rel_linf", "sigma_median", "rgb_linf", "rgb_median"], map(float, z[k])))
                                        for k in ("delta_1_vs_0", "delta_2_vs_1", "delta_2_vs_0")}
path = os.path.join(ROOT, "profiles", "parity_r2.json")
prev = {}
if os.path.exists(path):
    prev = json.load(open(path))
prev["tcnn_rounding_gap"] = report
json.dump(prev, open(path, "w"), indent=1)
print("wrote", path)
