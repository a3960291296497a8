"""Per-phase device timeline of ONE cooperatively rendered 512x512 frame (DNeRFModel.render_image_sharded) on N GPUs:
CUDA events between the phases (eager launches, max over ranks per phase), printed by rank 0 as one JSON line.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_timeline.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import ops, parallel  # noqa: E402
from instantavatar_b200.models.dnerf import Rays  # noqa: E402
from instantavatar_b200.renderers.raymarcher_acc import BoundModel  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, hb, batch = bench.build_model(dev, 0)
model.eval()
H = W = 512
tile = 2048
torch.manual_seed(99)
jit = torch.rand((5, 64, 64, 64, 3), device=dev)
grid = model.renderer.density_grid_test
names = ["prep", "occ_query", "occ_allreduce", "occ_build", "ray_transform", "render", "cat", "all_gather"]
acc = np.zeros(len(names))
iters = 12
for it in range(iters + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    model.deformer.prepare_deformer(batch); model.net_coarse.initialize(model.deformer.bbox)
    ev[1].record()
    grid.aabb = model.deformer.get_bbox_deformed()
    grid._density = ops.occupancy_query(model.deformer.scene(model.net_coarse), jit, grid.aabb6(), getattr(grid, "_density", None), shard=(rank, world))
    ev[2].record()
    if world > 1:
        dist.all_reduce(grid._density, op=dist.ReduceOp.MAX)
    ev[3].record()
    grid.build_from_density(grid._density)
    ev[4].record()
    idx32 = parallel.shard_tiles_cached(H * W, rank, world, tile, dev, torch.int32)
    o, d, near, far = ops.transform_rays(model.deformer.w2s, batch["rays_o"], batch["rays_d"], idx32)
    ev[5].record()
    rays = Rays(o=o[None], d=d[None], near=near[None], far=far[None])
    model.renderer.image_width = W
    out = model.renderer.render_test(rays, BoundModel(model.deformer, model.net_coarse, True), None)
    ev[6].record()
    loc = torch.cat([out["rgb_coarse"].reshape(-1, 3), out["alpha_coarse"].reshape(-1, 1)], dim=1)
    ev[7].record()
    img = parallel.all_gather_image(loc, H * W, tile) if world > 1 else loc
    if img is not None:
        img = img.contiguous()
    ev[8].record()
    torch.cuda.synchronize()
    if it >= 3:
        ms = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        acc += ms.cpu().numpy()
if rank == 0:
    per = dict(zip(names, (acc / iters).round(4).tolist()))
    print(json.dumps({"n_gpus": world, "eager_phase_ms_max_over_ranks": per, "sum_ms": float(acc.sum() / iters)}), flush=True)

# ---- whole cooperative frame (CUDA graph), NCCL and peer-memory paths, rays per warp of the renderer swept: with 1/N of the
#      rays a rank has fewer active ray tiles than resident warps, so smaller tiles (more look-ahead per ray) may pay ----
if world > 1:
    from instantavatar_b200.graphs import GraphedShardedFrame
    b = {k: v.clone() for k, v in batch.items()}
    for k in ("betas", "body_pose", "global_orient", "transl"):
        dist.broadcast(b[k], 0)
    peer = None
    try:
        peer = parallel.PeerFrame(H * W, dev)
    except Exception as exc:
        if rank == 0:
            print(json.dumps({"peer_unavailable": str(exc)[:200]}), flush=True)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for path, pf in (("nccl", None), ("peer", peer)):
        if path == "peer" and pf is None:
            continue
        for rpw in (4, 2, 1, 8):
            model.sharded_render_rays_per_warp = rpw   # overrides the tile size render_image_sharded picks for this world size
            try:
                g = GraphedShardedFrame(model, b, (H, W), rank, world, jit, peer=pf)
                run = lambda: g()
            except Exception:
                torch.cuda.synchronize()
                run = lambda: model.render_image_sharded(b, (H, W), rank, world, jit, peer=pf)
            for _ in range(3):
                run()
            dist.barrier(); torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
            for a_, e_ in ev:
                flush.zero_(); a_.record(); run(); e_.record()
            dist.barrier(); torch.cuda.synchronize()
            ms = torch.tensor([sum(a_.elapsed_time(e_) for a_, e_ in ev) / len(ev)], device=dev, dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"n_gpus": world, "path": path, "render_rays_per_warp": rpw, "frame_ms": float(ms.item())}), flush=True)
    model.sharded_render_rays_per_warp = None
    dist.destroy_process_group()
