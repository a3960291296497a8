"""ORACLE (test infrastructure): host-side control flow of the reference renderer in numpy,
calling the C restatements for the per-element kernels.

Restates /root/reference/instant_avatar/
  renderers/raymarcher_acc.py:25-36 (composite), :82-138 (render_test), :140-186 (render_train)
  deformers/snarf_deformer.py:109-159 (deform / deform_test / deform_train)
  deformers/fast_snarf/deformer_torch.py:100-116 (broyden_cuda + filter)
  models/structures/density_grid.py:46-125 (DensityGrid.update / initialize / max_connected_component)
  utils/loss.py:53-79 (NeRFLoss)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy import ndimage

from . import capi
from .frame import INIT_BONES

f32 = np.float32


@dataclass
class Net:
    """NeRFNGPNet state: flat fp32 params (tcnn ordering) + bbox normalisation (ngp.py:64-71)."""
    enc: np.ndarray
    col: np.ndarray
    center: np.ndarray
    scale: np.ndarray
    emulate: bool = True

    @staticmethod
    def from_bbox(enc, col, bbox, emulate=True):
        c = ((bbox[0] + bbox[1]) / f32(2)).astype(f32)
        s = (bbox[1] - bbox[0]).astype(f32)
        return Net(np.asarray(enc, f32), np.asarray(col, f32), c, s, emulate)

    def __call__(self, x):
        return capi.ngp_forward(x, self.center, self.scale, self.enc, self.col, self.emulate)


def search(pts, frame, subj):
    """deformer_torch.py:85-116: Broyden from 13 bone initialisations + duplicate filter."""
    xc, jinv, valid, iters = capi.broyden(pts, frame["voxel_J"], frame["tfs"], INIT_BONES, subj.offset_kernel,
                                          subj.scale_kernel)
    mask = capi.filter_roots(xc, valid)
    return xc, mask, iters, jinv


def deform_query(pts, frame, subj, net: Net, eval_mode=True, return_aux=False):
    """snarf_deformer.py:126-159: per point, max density over valid canonical correspondences."""
    pts = np.asarray(pts, f32).reshape(-1, 3)
    M = len(pts)
    xc, valid, iters, _ = search(pts, frame, subj)
    rgb_c = np.zeros((M, len(INIT_BONES), 3), f32)
    sig_c = np.zeros((M, len(INIT_BONES)), f32) if eval_mode else np.full((M, len(INIT_BONES)), f32(-1e5), f32)
    if valid.any():
        s, c = net(xc[valid])
        if eval_mode:  # :137-138 nan_to_num(x, 0, 0, 0)
            s = np.nan_to_num(s, nan=0.0, posinf=0.0, neginf=0.0)
            c = np.nan_to_num(c, nan=0.0, posinf=0.0, neginf=0.0)
        sig_c[valid] = s
        rgb_c[valid] = c
    idx = np.argmax(sig_c, axis=-1)  # torch.max returns the first maximal index on CPU; ties are measure-zero
    sigma = np.take_along_axis(sig_c, idx[:, None], 1)[:, 0]
    rgb = np.take_along_axis(rgb_c, idx[:, None, None].repeat(3, 2), 1)[:, 0]
    if return_aux:
        return rgb, sigma, {"xc": xc, "valid": valid, "iters": iters, "idx": idx, "sig_c": sig_c}
    return rgb, sigma


def render_test(rays_o, rays_d, near, far, grid_field, grid_min, grid_max, model, bg_color=None, MAX_SAMPLES=256,
                MAX_BATCH_SIZE=291600, stats=None):
    """raymarcher_acc.py:82-138 with the same windowed host loop."""
    rays_o = np.ascontiguousarray(rays_o, f32).reshape(-1, 3)
    rays_d = np.ascontiguousarray(rays_d, f32).reshape(-1, 3)
    near = np.array(near, f32).reshape(-1).copy()
    far = np.ascontiguousarray(far, f32).reshape(-1)
    N = len(rays_o)
    color = np.zeros((N, 3), f32); depth = np.zeros(N, f32); no_hit = np.ones(N, f32); counter = np.zeros(N, f32)
    alive = np.arange(N, dtype=np.int64)
    step_size = ((far - near) / f32(MAX_SAMPLES)).astype(f32)
    offset = np.asarray(grid_min, f32); scale = (np.asarray(grid_max, f32) - offset).astype(f32)
    k = 0
    n_eval = 0
    while k < MAX_SAMPLES:
        N_alive = len(alive)
        if N_alive == 0:
            break
        N_step = max(min(MAX_BATCH_SIZE // N_alive, MAX_SAMPLES), 1)
        pts, d_new, z_new = capi.raymarch_test(rays_o, rays_d, near, far, alive, grid_field, scale, offset, step_size, N_step)
        mask = d_new > 0
        counter[alive] += mask.sum(-1).astype(f32)
        rgb_vals = np.zeros_like(pts); sigma_vals = np.zeros(pts.shape[:2], f32)
        if mask.any():
            r, s = model(pts[mask])
            rgb_vals[mask] = r; sigma_vals[mask] = s
            n_eval += int(mask.sum())
        capi.composite_test(rgb_vals, sigma_vals, d_new, z_new, alive, color, depth, no_hit, 0.01)
        alive = alive[(no_hit[alive] > 1e-4) & (z_new[:, -1] > 0)]
        k += N_step
    if stats is not None:
        stats["n_eval"] = n_eval
    bg = np.ones((N, 3), f32) if bg_color is None else np.asarray(bg_color, f32).reshape(-1, 3)
    color = color + no_hit[:, None] * bg
    return {"rgb": color.astype(f32), "depth": depth, "alpha": (f32(1) - no_hit).astype(f32), "counter": counter}


def composite_train(sigma_vals, dists):
    """raymarcher_acc.py:25-36 (thresh = 0)."""
    tau = np.maximum(sigma_vals, f32(0)) * dists
    alpha = (f32(1.0) - np.exp(-tau)).astype(f32)
    trans = np.concatenate([np.ones_like(alpha[..., :1]), np.cumprod((f32(1) - alpha + f32(1e-10)).astype(f32), axis=-1, dtype=f32)], -1)
    w = (alpha * trans[..., :-1]).astype(f32)
    return w, trans.astype(f32)


def render_train(rays_o, rays_d, near, far, grid_field, grid_min, grid_max, model, jitter, noise=None, bg_color=None,
                 MAX_SAMPLES=256, return_aux=False):
    """raymarcher_acc.py:140-186.  `jitter` [N,S] replaces torch.rand_like (:158), `noise` [N,S] (or None)
    replaces noise * torch.randn_like (:167) -- random tensors are injected for reproducibility."""
    rays_o = np.ascontiguousarray(rays_o, f32).reshape(-1, 3)
    rays_d = np.ascontiguousarray(rays_d, f32).reshape(-1, 3)
    near = np.ascontiguousarray(near, f32).reshape(-1); far = np.ascontiguousarray(far, f32).reshape(-1)
    S = MAX_SAMPLES
    step_size = ((far - near) / f32(S)).astype(f32)
    offset = np.asarray(grid_min, f32); scale = (np.asarray(grid_max, f32) - offset).astype(f32)
    z_vals = capi.raymarch_train(rays_o, rays_d, near, far, grid_field, scale, offset, step_size, S)
    mask = z_vals > 0
    z_vals = (z_vals + np.asarray(jitter, f32) * step_size[:, None]).astype(f32)
    pts = (z_vals[..., None] * rays_d[:, None] + rays_o[:, None]).astype(f32)
    rgb_vals = np.zeros_like(pts); sigma_vals = np.full(z_vals.shape, f32(-1e3), f32)
    aux = None
    if mask.sum() > 0:
        if return_aux:
            r, s, aux = model(pts[mask], True)
        else:
            r, s = model(pts[mask])
        rgb_vals[mask] = r; sigma_vals[mask] = s
    if noise is not None:
        sigma_vals = (sigma_vals + np.asarray(noise, f32)).astype(f32)
    dists = np.ones_like(sigma_vals) * step_size[:, None]
    w, trans = composite_train(sigma_vals, dists)
    no_hit = trans[..., -1]
    color = (w[..., None] * rgb_vals).sum(-2)
    bg = np.ones((len(rays_o), 3), f32) if bg_color is None else np.asarray(bg_color, f32).reshape(-1, 3)
    color = (color + no_hit[:, None] * bg).astype(f32)
    depth = (w * z_vals).sum(-1).astype(f32)
    out = {"rgb": color, "depth": depth, "alpha": w.sum(-1).astype(f32), "weights": w, "mask": mask, "z": z_vals,
           "sigma": sigma_vals, "rgb_vals": rgb_vals, "pts": pts}
    if return_aux:
        out["aux"] = aux
    return out


def nerf_loss(pred, target_rgb, target_alpha, w_rgb=1.0, w_alpha=0.1, w_reg=0.1):
    """utils/loss.py:58-79 (float64 accumulation for a stable reference value)."""
    OFFSET = 0.313262
    l_rgb = np.mean((pred["rgb"].astype(np.float64) - target_rgb) ** 2)
    l_a = np.mean((pred["alpha"].astype(np.float64) - target_alpha) ** 2)
    reg = lambda x: np.mean(-np.log(np.exp(-x.astype(np.float64)) + np.exp(x.astype(np.float64) - 1))) + OFFSET
    ra, rd = reg(pred["alpha"]), reg(pred["weights"])
    return {"mse_loss": l_rgb, "loss_alpha_coarse": l_a, "reg_alpha": ra, "reg_density": rd,
            "loss": w_rgb * l_rgb + w_alpha * l_a + w_reg * ra + w_reg * rd}


# ------------------------------------------------------------------------------------------------
# occupancy grid: models/structures/density_grid.py
# ------------------------------------------------------------------------------------------------
def max_pool3(x):
    """F.max_pool3d(kernel 3, stride 1, padding 1) -- implicit -inf padding."""
    return ndimage.maximum_filter(x, size=3, mode="constant", cval=-np.inf)


def max_connected_component(grid: np.ndarray):
    """density_grid.py:118-125: 3*G rounds of 3x3x3 max-pool label flooding (float32 labels)."""
    G = grid.shape[-1]
    comp = np.arange(1, grid.size + 1, dtype=f32).reshape(grid.shape)
    comp[~grid] = 0
    g = grid.astype(f32)
    for _ in range(G * 3):
        new = max_pool3(comp) * g
        if np.array_equal(new, comp):  # fixed point: further rounds are no-ops
            break
        comp = new
    return comp


def _field_from_density(density: np.ndarray):
    """density_grid.py:78-85 / :104-110"""
    field = (f32(1) - np.exp(f32(0.01) * -density)).astype(f32)
    field = max_pool3(field)
    thr = min(f32(field.mean(dtype=f32)), f32(0.01))
    field = field > thr
    mcc = max_connected_component(field)
    labels = mcc[field]
    if labels.size == 0:
        return field & False
    vals, counts = np.unique(labels, return_counts=True)
    label = vals[np.argmax(counts)]  # torch.mode: most frequent, smallest value on ties (np.unique is sorted)
    return mcc == label


def grid_coords(G=64):
    idx = np.arange(G)
    c = np.stack(np.meshgrid(idx, idx, idx, indexing="ij"), -1).astype(f32) / f32(G)
    return c


def density_grid_initialize(deform_fn, aabb, jitters, G=64):
    """DensityGrid.initialize (density_grid.py:94-110). jitters [iters,G,G,G,3] replace torch.rand_like.
    deform_fn(pts) -> (rgb, sigma) in eval mode."""
    coords = grid_coords(G)
    density = np.zeros((G, G, G), f32)
    for j in jitters:
        c = (coords + np.asarray(j, f32) / f32(G)) * (aabb[1] - aabb[0]) + aabb[0]
        _, d = deform_fn(c.reshape(-1, 3).astype(f32))
        density = np.maximum(density, d.reshape(G, G, G))
    return _field_from_density(density), density


def density_grid_update(deform_fn, aabb, jitter, density_cached, old_field, step, G=64):
    """DensityGrid.update (density_grid.py:46-92), non-smpl_init branch.
    deform_fn(pts) -> (rgb, sigma) in TRAIN mode. Returns (density_reg [G,G,G], valid, new_cached, new_field)."""
    coords = grid_coords(G)
    c = (coords + np.asarray(jitter, f32) / f32(G)) * (aabb[1] - aabb[0]) + aabb[0]
    _, sig = deform_fn(c.reshape(-1, 3).astype(f32))
    sig = sig.reshape(G, G, G)
    density = np.clip(sig, 0, None).astype(f32)
    cached = np.maximum(density_cached * f32(0.8), density)
    field = _field_from_density(cached)
    dens_reg = (f32(1) - np.exp(f32(0.01) * -np.maximum(density, 0))).astype(f32)
    valid = field if step < 500 else old_field
    return dens_reg, valid, cached, field
