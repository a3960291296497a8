"""What would a reduced-precision skinning field cost in parity?  (fp16 x-pair records would halve the
sectors per trilinear footprint, 12 -> 6, on kernels bound by the L1 data pipe.)  Measured WITHOUT building the kernels:
the fp32 field that ia_precompute writes is rounded to the candidate storage precision and back (same layout, same
kernels, same arithmetic), the complete 512x512 frame is rendered through the public API, and compared with the frame
of the unrounded field -- the number of rays that leave the 1e-3 contract because of the storage format alone.
    python scripts/field_precision_parity.py        (one GPU)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
H = W = 512
_precompute = ops.precompute
MODE = {"name": "fp32"}


def rounded_precompute(*a, **k):
    fld, vd, aabb = _precompute(*a, **k)
    m = MODE["name"]
    if m == "fp16":
        fld = fld.half().float()
    elif m == "bf16":
        fld = fld.bfloat16().float()
    elif m == "fp16_rot_fp32_trans":
        # record = 2 x-neighbours x 12 coefficients (3x4 row-major: column 3 is the translation)
        v = fld.view(*fld.shape[:-1], 2, 3, 4)
        q = v.half().float()
        q[..., 3] = v[..., 3]
        fld = q.view(fld.shape).contiguous()
    return fld.contiguous(), vd, aabb


ops.precompute = rounded_precompute
out = {}
for frame in (0, 57):
    model, hb, batch = bench.build_model(dev, frame)
    model.eval()
    torch.manual_seed(3)
    jit = torch.rand((5, 64, 64, 64, 3), device=dev)
    res = {}
    for mode in ("fp32", "fp16", "fp16_rot_fp32_trans", "bf16"):
        MODE["name"] = mode
        rgb, depth, alpha, counter = model.render_image_fast(batch, (H, W), jitters=jit)
        res[mode] = (rgb.reshape(-1, 3).clone(), alpha.reshape(-1).clone(), model.renderer.density_grid_test.density_field.clone())
    MODE["name"] = "fp32"
    r0, a0, g0 = res["fp32"]
    for mode in ("fp16", "fp16_rot_fp32_trans", "bf16"):
        r, a, g = res[mode]
        e = torch.maximum((r - r0).abs().max(-1).values, (a - a0).abs())
        out[f"frame{frame}/{mode}"] = {"rays_above_1e-3": int((e > 1e-3).sum()), "rays_above_1e-2": int((e > 1e-2).sum()),
                                       "max_err": float(e.max()), "rays_hit": int((a0 > 0).sum()),
                                       "occupancy_cells_flipped": int((g != g0).sum())}
print(json.dumps(out), flush=True)
