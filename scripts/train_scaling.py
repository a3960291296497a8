"""Strong-scaled training step on N GPUs (4096 rays split over the ranks), split into its parts with CUDA-graph replays so
that host launch latency does not pollute the numbers: the whole step for each `train_rays_per_warp`, the optimiser part
alone (finite check, poison, reduce-scatter, shard Adam, all-gather, MLP refresh) and the two collectives alone.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_scaling.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instantavatar_b200 import ops  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, hb, batch = bench.build_model(dev, bench.FRAMES[rank % len(bench.FRAMES)])
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = {"n_gpus": world}
for trpw in (2, 1):
    r = bench.bench_train(model, batch, dev, rank, world, flush, steps=40, warmup=25, train_rays_per_warp=trpw)
    out[f"train_ms_rays_per_warp_{trpw}"] = r["ms_per_step"]


def graph_ms(fn, iters=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# the compute part alone: the same captured step with the optimiser step and the scaler update disabled
_step, _upd = model.optimizer.step, model.scaler.update
model.optimizer.step = lambda *a, **k: None
model.scaler.update = lambda: None
for trpw in (2, 1):
    r = bench.bench_train(model, batch, dev, rank, world, flush, steps=40, warmup=25, train_rays_per_warp=trpw)
    out[f"train_compute_only_ms_rays_per_warp_{trpw}"] = r["ms_per_step"]
model.optimizer.step, model.scaler.update = _step, _upd
model.optimizer.zero_grad()
model.world_size = world
out["optimizer_step_ms"] = graph_ms(lambda: model.optimizer.step(model.scaler, world))
if world > 1:
    from instantavatar_b200 import parallel
    from instantavatar_b200.optim import shard_layout
    S, L = shard_layout(model.optimizer.n, world)
    g_ = torch.zeros(L, device=dev); sh = torch.zeros(S, device=dev); h = torch.zeros(L, device=dev, dtype=torch.float16)
    out["reduce_scatter_ms"] = graph_ms(lambda: parallel.reduce_scatter_sum(sh, g_))
    out["all_gather_fp16_ms"] = graph_ms(lambda: parallel.all_gather_inplace(h))
    a_ = torch.zeros(L, device=dev)
    out["all_reduce_fp32_ms (round-1 collective)"] = graph_ms(lambda: dist.all_reduce(a_))
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
