"""Profiling driver (run under ncu with --profile-from-start off): two rendered frames and two training steps
(one with, one without the occupancy-grid refresh) of the bench workload."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

device = torch.device("cuda", 0)
model, hb, batch = bench.build_model(device, 0)
flush = torch.empty(1, device=device)
for _ in range(3):
    model.render_image_fast(dict(batch), (bench.H, bench.W))
bench.bench_train(model, batch, device, 0, 1, flush, steps=2, warmup=21)   # leaves global_step = 2023
torch.cuda.synchronize()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.profiler.start()
if what in ("all", "render"):
    model.eval()
    for _ in range(2):
        model.render_image_fast(dict(batch), (bench.H, bench.W))
if what in ("all", "train"):
    model.global_step = 2039
    bench.bench_train.__globals__  # noqa
    # two more training steps: step 2039 (no refresh) and 2040 (refresh)
    import torch as _t
    from instantavatar_b200 import parallel
    b = dict(batch)
    idx = _t.cat([((_t.arange(y0, y0 + 32))[:, None] * bench.W + _t.arange(x0, x0 + 32)[None]).reshape(-1)
                  for (y0, x0) in ((150, 240), (200, 232), (250, 236), (300, 240))]).to(device)
    for k in ("rays_o", "rays_d", "near", "far"):
        b[k] = batch[k][:, idx].contiguous()
    n = len(idx)
    b["bg_color"] = _t.rand((1, n, 3), device=device); b["alpha"] = _t.ones((1, n), device=device); b["rgb"] = _t.rand((1, n, 3), device=device)
    model.training_step(b)
    model.training_step(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", what)
