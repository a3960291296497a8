"""Training step with the one-kernel forward (ia_train_fwd, 2 or 1 rays per warp) against the split forward
(ia_train_fwd_split: march -> sample list -> point query over the list -> compositing) at 4096 / 2048 / 1024 / 512 rays per
step on ONE GPU (the per-rank share of a 4096-ray step on 1 / 2 / 4 / 8 GPUs, without the exchange).  One JSON line.
    python scripts/train_fwd_split.py"""
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
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = {}
for rays in (4096, 2048, 1024, 512):
    for name, split, trpw, k in (("fused_2_rays_per_warp", 0, 2, 0), ("fused_1_ray_per_warp", 0, 1, 0), ("split_1_lane_per_sample", 1, 2, 1),
                                 ("split_2_lanes_per_sample", 1, 2, 2), ("split_4_lanes_per_sample", 1, 2, 4), ("split_auto", 1, 2, 0)):
        model, hb, batch = bench.build_model(dev, 0)
        ops.set_option("train_split", split)
        ops.set_option("query_lanes_per_sample", k)
        r = bench.bench_train(model, batch, dev, 0, 1, flush, steps=40, warmup=25, train_rays_per_warp=trpw, rays_cap=rays)
        out[f"train_ms/{rays}_rays/{name}"] = round(r["ms_per_step"], 4)
        del model
ops.set_option("train_split", 1); ops.set_option("query_lanes_per_sample", 0)
print(json.dumps(out), flush=True)
