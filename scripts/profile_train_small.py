"""Launch list of ONE training step on 1 GPU at a reduced ray count (what a rank of an N-GPU step computes): run under
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python scripts/profile_train_small.py 512
to see which kernels of the step do not shrink with the number of rays."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from instantavatar_b200 import ops  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 512
trpw = int(sys.argv[2]) if len(sys.argv) > 2 else 2
split = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
model, hb, batch = bench.build_model(dev, 0)
ops.set_option("train_rays_per_warp", trpw)
ops.set_option("train_split", split)
rgb, _, alpha, _ = model.render_image_fast(dict(batch), (bench.H, bench.W))
rgb, alpha = rgb.reshape(-1, 3), alpha.reshape(-1)
idx = torch.cat([((torch.arange(y0, y0 + 32))[:, None] * bench.W + torch.arange(x0, x0 + 32)[None]).reshape(-1)
                 for (y0, x0) in ((150, 240), (200, 232), (250, 236), (300, 240))])[:n_rays].to(dev)
b = dict(batch)
for k in ("rays_o", "rays_d", "near", "far"):
    b[k] = batch[k][:, idx].contiguous()
n = len(idx)
b["bg_color"] = torch.rand((1, n, 3), device=dev); b["alpha"] = alpha[idx][None]; b["rgb"] = rgb[idx][None].clone()
model.global_step = 2000   # the first warm-up step refreshes the train occupancy grid (every 20th step does)
for _ in range(5):
    model.training_step(b)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
ev0.record()
model.training_step(b)   # step 2005: no grid refresh
ev1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("rays", n, "train_rays_per_warp", trpw, "split", split, "eager step ms", ev0.elapsed_time(ev1))
