"""Load-balance tail of the occupancy query when one frame is split over G GPUs, measured on ONE GPU by running each rank's
shard in turn: per-batch cost distribution (SM cycles written by the kernel), the kernel time of every shard with the
natural (strided) order, with the shard's batches started longest-first using THIS frame's costs (the bound), and using
the costs of the PREVIOUS frame of the track (what a renderer can actually know).  One JSON line per G.
    python scripts/query_schedule.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model, hb, batch = bench.build_model(dev, 0)
model.eval()
grid = model.renderer.density_grid_test
P, G3 = 5, 64
nb = ops.occupancy_batches(G3, P)
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)


def frame_state(frame_batch, seed):
    model.deformer.prepare_deformer(frame_batch)
    model.net_coarse.initialize(model.deformer.bbox)
    grid.aabb = model.deformer.get_bbox_deformed()
    torch.manual_seed(seed)
    return torch.rand((P, G3, G3, G3, 3), device=dev)


def timed(fn, iters=8):
    ms = []
    for i in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    return float(np.median(ms))


# previous frame of the track (another pose, other jitter) -> the costs a renderer would have at hand
b_prev = {k: v.clone() for k, v in batch.items()}
for k in ("body_pose", "global_orient"):
    b_prev[k] = b_prev[k] + 0.03 * torch.randn_like(b_prev[k])   # ~2 degrees per joint: a fast motion between two frames
jit_prev = frame_state(b_prev, 7)
cost_prev = torch.zeros(nb, dtype=torch.int32, device=dev)
ops.occupancy_query(model.deformer.scene(model.net_coarse), jit_prev, grid.aabb6(), cost=cost_prev)

jit = frame_state(batch, 99)
scene = model.deformer.scene(model.net_coarse)
aabb6 = grid.aabb6()
cost = torch.zeros(nb, dtype=torch.int32, device=dev)
ref = ops.occupancy_query(scene, jit, aabb6, cost=cost).clone()
ref2 = ops.occupancy_query(scene, jit, aabb6).clone()
assert torch.equal(ref, ref2)
c = cost.cpu().numpy().astype(np.float64)
cp = cost_prev.cpu().numpy().astype(np.float64)
clk_mhz = 1.0
try:
    import pynvml
    pynvml.nvmlInit()
    clk_mhz = float(pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(0), pynvml.NVML_CLOCK_SM))
except Exception:
    clk_mhz = 1900.0
us = c / clk_mhz
full_ms = timed(lambda: ops.occupancy_query(scene, jit, aabb6))
print(json.dumps({"batches": nb, "sm_mhz_assumed": clk_mhz, "full_query_ms": full_ms,
                  "batch_us": {"mean": us.mean(), "p50": float(np.percentile(us, 50)), "p90": float(np.percentile(us, 90)),
                               "p99": float(np.percentile(us, 99)), "max": us.max()},
                  "corr_prev_frame_cost": float(np.corrcoef(c, cp)[0, 1]),
                  "share_of_cycles_in_top_10pct_batches": float(np.sort(c)[::-1][: nb // 10].sum() / c.sum())}), flush=True)

for world in (1, 2, 4, 8):
    rows = {"natural": [], "lpt_same_frame": [], "lpt_prev_frame": [], "two_lanes_per_point": [], "four_lanes_per_point": []}
    dens = torch.zeros_like(ref)
    dens_k = {2: torch.zeros_like(ref), 4: torch.zeros_like(ref)}
    for r in range(world):
        mine = torch.arange(r, nb, world, device=dev, dtype=torch.int64)
        ops.set_option("occupancy_lanes_per_point", 1)
        rows["natural"].append(timed(lambda: ops.occupancy_query(scene, jit, aabb6, shard=(r, world))))
        # narrow batches: 2 / 4 lanes share a point's 13 root finds (batch latency / 2, / 3.25)
        for k, name in ((2, "two_lanes_per_point"), (4, "four_lanes_per_point")):
            ops.set_option("occupancy_lanes_per_point", k)
            dens_k[k] = torch.maximum(dens_k[k], ops.occupancy_query(scene, jit, aabb6, shard=(r, world)))
            rows[name].append(timed(lambda: ops.occupancy_query(scene, jit, aabb6, shard=(r, world))))
        ops.set_option("occupancy_lanes_per_point", 1)
        for name, cc in (("lpt_same_frame", cost), ("lpt_prev_frame", cost_prev)):
            order = mine[torch.argsort(cc[mine], descending=True, stable=True)].to(torch.int32).contiguous()
            d = ops.occupancy_query(scene, jit, aabb6, order=order)
            if name == "lpt_prev_frame":
                dens = torch.maximum(dens, d)
            rows[name].append(timed(lambda: ops.occupancy_query(scene, jit, aabb6, order=order)))
    assert torch.equal(dens, ref), "ordered shards do not reproduce the full grid"
    assert torch.equal(dens_k[2], ref) and torch.equal(dens_k[4], ref), "narrow batches do not reproduce the full grid"
    print(json.dumps({"n_shards": world, "ideal_ms": full_ms / world,
                      **{k + "_ms_max_over_shards": max(v) for k, v in rows.items()},
                      **{k + "_ms_mean": float(np.mean(v)) for k, v in rows.items()}}), flush=True)
