import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_pose_grad as T
from instantavatar_b200 import ops
orig = ops.pose_grad
def dbg(scene, lbs, xd, best, denc, count, g):
    orig(scene, lbs, xd, best, denc, count, g)
    n = int(count.item())
    print("pose_grad: count", n, "cap", xd.shape[0], "best>=0", int((best[:n] >= 0).sum()), "denc norm", float(denc[:n].norm()),
          "nan", bool(torch.isnan(denc[:n]).any()), "g norm", float(g.norm()), "tfs.requires_grad", scene.tfs.requires_grad)
ops.pose_grad = dbg
model, batch, rgb_gt, alpha_gt, (H, W) = T._gt_and_model()
ys, xs = np.arange(36, 96), np.arange(44, 86)
sel = torch.from_numpy((ys[:, None] * W + xs[None]).ravel()).cuda()
pose0 = {k: batch[k].clone() for k in ("betas", "global_orient", "body_pose", "transl")}
for fused in (True, False):
    model.global_step = 1
    model.fused_loss = fused
    model.enable_pose_optimisation(pose0, is_refine=True)
    b = T._ray_batch(batch, sel[:1024], rgb_gt, alpha_gt, 0)
    out = model.training_step(b, noise_tensor=torch.zeros((1024, 256), device="cuda"))
    print("fused", fused, "loss", float(out["loss"]), {k: float(getattr(model.SMPL_param, k).weight.grad.norm()) for k in ("global_orient", "body_pose", "transl")},
          "found_inf", float(model.scaler.found_inf), "scale", float(model.scaler.scale_t))
