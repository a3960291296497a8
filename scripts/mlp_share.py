"""Two measured answers (one GPU, bench scene):
 (1) What do the MLPs cost?  CUDA-event times of the stand-alone operators on the frame's ~0.55 M network evaluations:
     hash encode + density MLP (ia_tcnn_encoder_forward), the colour MLP alone on tensor cores (ia_tcnn_mlp_forward),
     and both fused (ia_ngp_forward) -- against the 1.1 ms fused renderer that contains them.  This bounds what a tcgen05
     rewrite of the 64-wide layers could return.
 (2) What does the explicit-FMA discipline (-fmad=false, bit-identical geometry) cost?  Run this script again with
     IA_B200_LIB=instantavatar_b200/libia_b200_fmad.so (scripts/build_variant.sh fmad): same kernels, free contraction."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import _lib, ops  # noqa: E402
from instantavatar_b200.models.dnerf import Rays  # noqa: E402
from instantavatar_b200.renderers.raymarcher_acc import BoundModel  # noqa: E402

dev = torch.device("cuda", 0)
model, hb, batch = bench.build_model(dev, 0)
model.eval()
model.render_image_fast(dict(batch), (512, 512))
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n=9):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


out = {"lib": os.path.basename(_lib.LIB_PATH)}
grid = model.renderer.density_grid_test
scene_q = model.deformer.scene(model.net_coarse)
torch.manual_seed(5)
jit = torch.rand((5, 64, 64, 64, 3), device=dev)
out["occupancy_query_ms"] = timed(lambda: ops.occupancy_query(scene_q, jit, grid.aabb6()))
r = Rays(o=batch["rays_o"].clone(), d=batch["rays_d"].clone(), near=batch["near"].clone(), far=batch["far"].clone())
model.deformer.transform_rays_w2s(r)
bm = BoundModel(model.deformer, model.net_coarse, True)
model.renderer.image_width = 512
out["render_ms"] = timed(lambda: model.renderer.render_test(r, bm, None))
# stand-alone network operators on as many canonical points as the frame evaluates
n = 553463
bb = model.deformer.bbox
x = (torch.rand((n, 3), device=dev) * 0.3 + 0.35) * (bb[1] - bb[0]) + bb[0]
x01 = ((x - model.net_coarse.center) / model.net_coarse.scale + 0.5).clamp(0, 1).contiguous()
sc = ops.Scene(table_h=scene_q.table_h, mlp_h=scene_q.mlp_h, net_center=scene_q.net_center, net_scale=scene_q.net_scale)
in15 = torch.rand((n, 15), device=dev)
out["n_points"] = n
out["ngp_forward_ms (hash + 5 layers)"] = timed(lambda: ops.ngp_forward(sc, x))
out["encoder_forward_ms (hash + 2 layers)"] = timed(lambda: ops.tcnn_encoder_forward(sc, x01))
out["colour_mlp_forward_ms (3 layers, no hash)"] = timed(lambda: ops.tcnn_mlp_forward(scene_q.mlp_h, in15))
flops = n * 2 * (32 * 64 + 64 * 16 + 16 * 64 + 64 * 64 + 64 * 16)
out["mlp_flops_per_frame"] = flops
out["colour_mlp_TFLOPs_achieved"] = n * 2 * (16 * 64 + 64 * 64 + 64 * 16) / (out["colour_mlp_forward_ms (3 layers, no hash)"] * 1e-3) / 1e12
# optimiser step alone (flat 13 M parameters: finite check + Adam + fp16 image + MLP block refresh)
opt, scaler = model.optimizer, model.scaler
for name, fn in (("optimizer_step_ms", lambda: opt.step(scaler, 1)),
                 ("grad_check_finite_ms", lambda: ops.grad_check_finite(opt.flat_g[:opt.n], scaler.found_inf)),
                 ("adam_kernel_ms", lambda: ops.adam_step_dev(opt.flat_p[:opt.n], opt.flat_g[:opt.n], opt.flat_m[:opt.n], opt.flat_v[:opt.n], opt.state_t,
                                                              scaler.found_inf, opt.flat_h[:opt.n], 0))):
    out[name] = timed(fn)
scaler.found_inf.zero_()
out["adam_bytes"] = opt.n * (4 * 4 + 4 * 4 + 2)
out["adam_GBps"] = out["adam_bytes"] / (out["adam_kernel_ms"] * 1e-3) / 1e9
print(json.dumps(out))
