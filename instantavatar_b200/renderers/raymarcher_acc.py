"""Host-side mirror of instant_avatar/renderers/raymarcher_acc.py::Raymarcher.

Same constructor, `initialize(N)`, `__call__(rays, model, eval_mode, noise, bg_color)` and result dictionary as the
reference.  When `model` is bound to a SNARFDeformer + NeRFNGPNet pair (a `BoundModel`, or the reference's
`lambda x, _: self.deformer(x, self.net_coarse, eval_mode)` closure) the whole per-ray path -- occupancy-grid march,
Broyden root finding, hash grid + MLPs, compositing -- runs as one fused kernel (`ia_render_fwd`, and the
`ia_train_fwd/bwd` pair for training) instead of the reference's host-synchronous window loop.
"""
from __future__ import annotations

import torch

from .. import ops
from ..models.structures.density_grid import DensityGrid


class BoundModel:
    """`model(x, d)` callable carrying the deformer/network it is bound to (what DNeRFModel.forward builds)."""

    def __init__(self, deformer, net, eval_mode=True):
        self.deformer, self.net, self.eval_mode = deformer, net, eval_mode

    def __call__(self, x, _=None):
        return self.deformer(x, self.net, self.eval_mode)


def _unwrap(model):
    """find (deformer, net) behind a model callable; None if it is not a SNARFDeformer + NeRFNGPNet pair (the fused
    kernels implement exactly that pair; anything else goes through the kernel-for-kernel legacy path)"""
    pair = None
    if hasattr(model, "deformer") and hasattr(model, "net"):
        pair = (model.deformer, model.net)
    else:
        for cell in getattr(model, "__closure__", None) or ():
            obj = cell.cell_contents
            if hasattr(obj, "deformer") and hasattr(obj, "net_coarse"):
                pair = (obj.deformer, obj.net_coarse)
                break
    if pair is None or not hasattr(pair[0], "scene") or not hasattr(pair[1], "half_params"):
        return None
    return pair


class Raymarcher(torch.nn.Module):
    def __init__(self, MAX_SAMPLES: int = 256, MAX_BATCH_SIZE: int = 291600, smpl_init: bool = False, device="cuda") -> None:
        super().__init__()
        if MAX_SAMPLES != 256:
            raise ValueError("the fused kernels are built for MAX_SAMPLES = 256 (confs/renderer/raymarcher_acc.yaml)")
        self.MAX_SAMPLES = MAX_SAMPLES
        self.MAX_BATCH_SIZE = MAX_BATCH_SIZE
        self.aabb = torch.tensor([[-1.25, -1.55, -1.25], [1.25, 0.95, 1.25]]).float().to(device)
        self.density_grid_test = DensityGrid(64, device=device)
        self.smpl_init = smpl_init
        self.idx = 0
        self.image_width = 0  # optional hint: rays form a row-major image of this width

    def initialize(self, N):
        dev = self.aabb.device
        self.density_grid_train_all = [DensityGrid(64, self.aabb, device=dev)]

    @property
    def density_grid_train(self):
        return self.density_grid_train_all[min(self.idx, len(self.density_grid_train_all) - 1)]

    def __call__(self, rays, model, eval_mode=True, noise=0, bg_color=None):
        if eval_mode:
            return self.render_test(rays, model, bg_color)
        return self.render_train(rays, model, noise, bg_color)

    @torch.no_grad()
    def render_test_legacy(self, rays, model, bg_color):
        """raymarcher_acc.py:82-138 verbatim control flow on the kernel-for-kernel operators, for any `model(pts, None)`
        callable (host-synchronous window loop, as in the reference)."""
        device = rays.o.device
        rays_o = rays.o.reshape(-1, 3).float().contiguous()
        rays_d = rays.d.reshape(-1, 3).float().contiguous()
        near = rays.near.reshape(-1).float().clone()
        far = rays.far.reshape(-1).float().contiguous()
        N = rays_o.shape[0]
        color = torch.zeros(N, 3, device=device); depth = torch.zeros(N, device=device)
        no_hit = torch.ones(N, device=device); counter = torch.zeros_like(depth)
        alive = torch.arange(N, device=device)
        step_size = ((far - near) / self.MAX_SAMPLES).contiguous()
        grid = self.density_grid_test
        offset = grid.min_corner.float().contiguous(); scale = (grid.max_corner - grid.min_corner).float().contiguous()
        k = 0
        while k < self.MAX_SAMPLES:
            N_alive = len(alive)
            if N_alive == 0:
                break
            N_step = max(min(self.MAX_BATCH_SIZE // N_alive, self.MAX_SAMPLES), 1)
            pts, d_new, z_new = ops.raymarch_test(rays_o, rays_d, near, far, alive, grid.density_field, scale, offset, step_size, N_step)
            counter[alive] += (d_new > 0).sum(dim=-1)
            mask = d_new > 0
            rgb_vals = torch.zeros_like(pts); sigma_vals = torch.zeros_like(rgb_vals[..., 0])
            if mask.any():
                r, s = model(pts[mask], None)
                rgb_vals[mask], sigma_vals[mask] = r.float(), s.float()
            ops.composite_test(rgb_vals, sigma_vals, d_new, z_new, alive, color, depth, no_hit, 0.01)
            alive = alive[(no_hit[alive] > 1e-4) & (z_new[:, -1] > 0)]
            k += N_step
        bg = bg_color.reshape(-1, 3) if bg_color is not None else 1.0
        color = color + no_hit[..., None] * bg
        return {"rgb_coarse": color.reshape(rays.o.shape), "depth_coarse": depth.reshape(rays.near.shape),
                "alpha_coarse": (1 - no_hit).reshape(rays.near.shape), "counter_coarse": counter.reshape(rays.near.shape)}

    @torch.no_grad()
    def render_test(self, rays, model, bg_color, stats=None):
        bound = _unwrap(model)
        if bound is None:
            return self.render_test_legacy(rays, model, bg_color)
        deformer, net = bound
        net.initialize(deformer.bbox)
        grid = self.density_grid_test
        scene = deformer.scene(net, grid.occupancy_bits(), grid.aabb6())
        rays_o = rays.o.reshape(-1, 3).float().contiguous()
        rays_d = rays.d.reshape(-1, 3).float().contiguous()
        near = rays.near.reshape(-1).float().contiguous()
        far = rays.far.reshape(-1).float().contiguous()
        bg = bg_color.reshape(-1, 3).float().contiguous() if bg_color is not None else None
        out = ops.render_fwd(scene, rays_o, rays_d, near, far, bg, self.image_width, stats)
        return {
            "rgb_coarse": out["rgb"].reshape(rays.o.shape),
            "depth_coarse": out["depth"].reshape(rays.near.shape),
            "alpha_coarse": out["alpha"].reshape(rays.near.shape),
            "counter_coarse": out["counter"].reshape(rays.near.shape),
        }

    def render_train_legacy(self, rays, model, noise, bg_color, jitter=None, noise_tensor=None):
        """raymarcher_acc.py:140-186 on the kernel-for-kernel march operator and torch compositing (autograd), for any
        differentiable `model(pts, None)` callable (e.g. SMPLDeformer + NeRFNGPNet)."""
        rays_o = rays.o.reshape(-1, 3).float().contiguous()
        rays_d = rays.d.reshape(-1, 3).float().contiguous()
        near = rays.near.reshape(-1).float().contiguous()
        far = rays.far.reshape(-1).float().contiguous()
        N_step = self.MAX_SAMPLES
        step_size = ((far - near) / N_step).contiguous()
        grid = self.density_grid_train
        offset = grid.min_corner.float().contiguous(); scale = (grid.max_corner - grid.min_corner).float().contiguous()
        with torch.no_grad():
            z_vals = ops.raymarch_train(rays_o.detach(), rays_d.detach(), near.detach(), far.detach(), grid.density_field, scale, offset,
                                        step_size.detach(), N_step)
        mask = z_vals > 0
        z_vals = z_vals + (torch.rand_like(z_vals) if jitter is None else jitter) * step_size[:, None]
        pts = z_vals[..., None] * rays_d[:, None] + rays_o[:, None]
        rgb_vals = torch.zeros_like(pts, dtype=torch.float32)
        sigma_vals = -torch.ones_like(rgb_vals[..., 0]) * 1e3
        if mask.any():
            r, s = model(pts[mask], None)
            mi = mask.nonzero(as_tuple=True)
            rgb_vals = rgb_vals.index_put(mi, r.float())
            sigma_vals = sigma_vals.index_put(mi, s.float())
        if noise_tensor is not None:
            sigma_vals = sigma_vals + noise_tensor
        elif noise > 0:
            sigma_vals = sigma_vals + noise * torch.randn_like(sigma_vals)
        dists = torch.ones_like(sigma_vals) * step_size[:, None]
        # composite (raymarcher_acc.py:25-36)
        alpha = 1.0 - torch.exp(-torch.relu(sigma_vals) * dists)
        trans = torch.cat([torch.ones_like(alpha[..., 0:1]), torch.cumprod(1 - alpha + 1e-10, dim=-1)], dim=-1)
        weights = alpha * trans[..., :-1]
        no_hit = trans[..., -1]
        color = (weights[..., None] * rgb_vals).sum(dim=-2)
        color = color + no_hit[..., None] * (bg_color.reshape(-1, 3) if bg_color is not None else 1.0)
        depth = (weights * z_vals).sum(dim=-1)
        return {"rgb_coarse": color.reshape(rays.o.shape), "depth_coarse": depth.reshape(rays.near.shape),
                "alpha_coarse": weights.sum(-1).reshape(rays.near.shape), "weight_coarse": weights.reshape(*rays.near.shape, -1)}

    def render_train(self, rays, model, noise, bg_color, jitter=None, noise_tensor=None):
        bound = _unwrap(model)
        if bound is None:
            return self.render_train_legacy(rays, model, noise, bg_color, jitter, noise_tensor)
        from ..autograd import render_train_fused
        deformer, net = bound
        return render_train_fused(self, deformer, net, rays, noise, bg_color, jitter, noise_tensor)
