"""Sweep of the residency knobs of the two frame kernels on the bench scene (one GPU): warps per persistent CTA of the
point-query kernel (12 / 16 / 20) and of the fused renderer (12 / 16) x rays per warp (4 / 8).  Prints one JSON line per
setting with the CUDA-event time of the kernel (median of 7, L2 flushed) and checks that results are bit-identical."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import ops  # noqa: E402
from instantavatar_b200.models.dnerf import Rays  # noqa: E402
from instantavatar_b200.renderers.raymarcher_acc import BoundModel  # noqa: E402

dev = torch.device("cuda", 0)
model, hb, batch = bench.build_model(dev, 0)
model.eval()
model.render_image_fast(dict(batch), (512, 512))
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n=7):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


grid = model.renderer.density_grid_test
scene_q = model.deformer.scene(model.net_coarse)
torch.manual_seed(5)
jit = torch.rand((5, 64, 64, 64, 3), device=dev)
ref_d = None
for qw in (12, 16, 20):
    ops.set_option("query_warps", qw)
    d = ops.occupancy_query(scene_q, jit, grid.aabb6()).clone()
    if ref_d is None:
        ref_d = d
    ms = timed(lambda: ops.occupancy_query(scene_q, jit, grid.aabb6()))
    print(json.dumps({"kernel": "deform_query_kernel", "query_warps": qw, "ms": ms, "bit_equal": bool(torch.equal(d, ref_d))}), flush=True)
ops.set_option("query_warps", 12)

r = Rays(o=batch["rays_o"].clone(), d=batch["rays_d"].clone(), near=batch["near"].clone(), far=batch["far"].clone())
model.deformer.transform_rays_w2s(r)
bm = BoundModel(model.deformer, model.net_coarse, True)
model.renderer.image_width = 512
ref_img = None
for rw in (12,):  # (16 was measured in round 2: profiles/sweep_warps_r2.jsonl; that build is gone)
    for rpw in (4, 8):
        ops.set_option("render_warps", rw); ops.set_option("render_rays_per_warp", rpw)
        out = model.renderer.render_test(r, bm, None)
        img = torch.cat([out["rgb_coarse"].reshape(-1, 3), out["alpha_coarse"].reshape(-1, 1)], 1).clone()
        if ref_img is None:
            ref_img = img
        ms = timed(lambda: model.renderer.render_test(r, bm, None))
        print(json.dumps({"kernel": "render_fwd_kernel", "render_warps": rw, "rays_per_warp": rpw, "ms": ms,
                          "bit_equal": bool(torch.equal(img, ref_img))}), flush=True)
