"""CPU: the host-side control flow of Raymarcher's kernel-for-kernel paths (windowed inference loop, training march +
torch compositing) with the three CUDA operators replaced by the oracle's C restatements of the same reference kernels
(raymarcher.cu) -- the loop must reproduce the oracle's own window loop (oracle/render.py, raymarcher_acc.py:82-186)."""
import numpy as np
import pytest
import torch

from instantavatar_b200 import ops
from instantavatar_b200.models.dnerf import Rays
from instantavatar_b200.renderers.raymarcher_acc import Raymarcher
from oracle import capi
from oracle import render as orender
from oracle import scene as oscene
from oracle import testing as scene_util

f32 = np.float32


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


@pytest.fixture()
def cpu_ops(monkeypatch):
    """oracle-backed stand-ins with the signatures and in-place semantics of ops.raymarch_* / ops.composite_test"""
    def raymarch_test(rays_o, rays_d, nears, fars, alives, grid, scale, offset, step_size, n_steps):
        near_np = _np(nears).astype(f32)
        pts, deltas, depths = capi.raymarch_test(_np(rays_o), _np(rays_d), near_np, _np(fars), _np(alives), _np(grid), _np(scale),
                                                 _np(offset), _np(step_size), n_steps)
        nears.copy_(torch.from_numpy(near_np))   # the operator advances `nears` in place
        return [torch.from_numpy(pts), torch.from_numpy(deltas), torch.from_numpy(depths)]

    def composite_test(rgb, sigma, delta, z, alive, color, depth, no_hit, thresh):
        c, dd, nh = _np(color).astype(f32), _np(depth).astype(f32), _np(no_hit).astype(f32)
        capi.composite_test(_np(rgb), _np(sigma), _np(delta), _np(z), _np(alive), c, dd, nh, thresh)
        color.copy_(torch.from_numpy(c)); depth.copy_(torch.from_numpy(dd)); no_hit.copy_(torch.from_numpy(nh))

    def raymarch_train(rays_o, rays_d, nears, fars, grid, scale, offset, step_size, n_steps):
        return torch.from_numpy(capi.raymarch_train(_np(rays_o), _np(rays_d), _np(nears), _np(fars), _np(grid), _np(scale), _np(offset),
                                                    _np(step_size), n_steps))

    monkeypatch.setattr(ops, "raymarch_test", raymarch_test)
    monkeypatch.setattr(ops, "composite_test", composite_test)
    monkeypatch.setattr(ops, "raymarch_train", raymarch_train)


def _setup(n_side=24):
    sc = scene_util.oracle_scene(0)
    fr = sc["frame"]
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    ys, xs = np.linspace(150, 360, n_side).astype(int), np.linspace(200, 310, n_side).astype(int)
    idx = (ys[:, None] * 512 + xs[None]).ravel()
    rays = Rays(o=torch.from_numpy(o[idx][None]), d=torch.from_numpy(d[idx][None]), near=torch.from_numpy(near[idx][None]),
                far=torch.from_numpy(far[idx][None]))
    rm = Raymarcher(256, 4000, device="cpu")   # small sample budget: several windows per frame
    rm.initialize(1)
    aabb = [torch.from_numpy(fr["bbox_deformed"][0]), torch.from_numpy(fr["bbox_deformed"][1])]
    for grid in (rm.density_grid_test, rm.density_grid_train):
        grid.aabb = aabb
        grid.density_field = torch.from_numpy(sc["occ"])
    return sc, fr, (o[idx], d[idx], near[idx], far[idx]), rays, rm


def test_windowed_inference_loop_matches_the_oracle_loop(cpu_ops):
    sc, fr, (o, d, near, far), rays, rm = _setup()
    oracle_model = scene_util.oracle_model(sc, True)

    def model(pts, _):   # foreign model: no .deformer/.net attributes -> kernel-for-kernel path
        rgb, sigma = oracle_model(_np(pts))
        return torch.from_numpy(rgb), torch.from_numpy(sigma)

    bg = np.random.default_rng(0).random((len(o), 3)).astype(f32)
    out = rm(rays, model, eval_mode=True, bg_color=torch.from_numpy(bg))
    ref = orender.render_test(o, d, near, far, sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1], oracle_model, bg_color=bg,
                              MAX_BATCH_SIZE=4000)
    assert (ref["alpha"] > 0.5).sum() > 50
    np.testing.assert_array_equal(_np(out["rgb_coarse"]).reshape(-1, 3), ref["rgb"])
    np.testing.assert_array_equal(_np(out["alpha_coarse"]).reshape(-1), ref["alpha"])
    np.testing.assert_array_equal(_np(out["depth_coarse"]).reshape(-1), ref["depth"])
    np.testing.assert_array_equal(_np(out["counter_coarse"]).reshape(-1), ref["counter"])


def test_training_march_and_compositing_match_the_oracle(cpu_ops):
    sc, fr, (o, d, near, far), rays, rm = _setup(16)
    aux_model = scene_util.oracle_model_aux(sc, False)
    n = len(o)
    rng = np.random.default_rng(1)
    jitter = rng.random((n, 256), dtype=f32); noise = rng.normal(0, 1, (n, 256)).astype(f32); bg = rng.random((n, 3), dtype=f32)

    def model(pts, _):
        rgb, sigma = aux_model(_np(pts))
        return torch.from_numpy(rgb), torch.from_numpy(sigma)

    out = rm.render_train(rays, model, 1, torch.from_numpy(bg), jitter=torch.from_numpy(jitter), noise_tensor=torch.from_numpy(noise))
    ref = orender.render_train(o, d, near, far, sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1], aux_model, jitter, noise, bg)
    assert (ref["alpha"] > 0.5).sum() > 20
    np.testing.assert_allclose(_np(out["weight_coarse"]).reshape(n, 256), ref["weights"], atol=2e-6)
    np.testing.assert_allclose(_np(out["rgb_coarse"]).reshape(-1, 3), ref["rgb"], atol=5e-6)
    np.testing.assert_allclose(_np(out["alpha_coarse"]).reshape(-1), ref["alpha"], atol=5e-6)
    np.testing.assert_allclose(_np(out["depth_coarse"]).reshape(-1), ref["depth"], atol=5e-5)


def test_density_grid_update_and_initialize_host_logic(monkeypatch):
    """DensityGrid.update (EMA of the cached density, 1 - exp(-0.01 d) regulariser input, `valid` selection) and the generic
    DensityGrid.initialize loop against oracle.render.density_grid_update / density_grid_initialize, with the CUDA grid
    post-processing replaced by the oracle's (density_grid.py:46-125)"""
    from instantavatar_b200.models.structures.density_grid import DensityGrid

    def occupancy_build(density, bits=None, want_field=True, workspace=None, field=None):
        f = torch.from_numpy(orender._field_from_density(_np(density).astype(f32)))
        if field is not None:
            field.copy_(f)
        else:
            field = f
        return field, torch.zeros(64 ** 3 // 32 + 8, dtype=torch.int32)

    monkeypatch.setattr(ops, "occupancy_build", occupancy_build)
    rng = np.random.default_rng(5)
    centres = rng.uniform(-0.6, 0.6, (3, 3)).astype(f32)

    def blob_np(p, scale=40.0):    # a smooth density with three blobs; train mode: may be negative
        d2 = ((p[:, None, :] - centres[None]) ** 2).sum(-1)
        sig = (scale * np.exp(-d2 / 0.02).sum(-1) - 2.0).astype(f32)
        return np.zeros((len(p), 3), f32), sig

    def deformer(pts, net, eval_mode=True):
        rgb, sig = blob_np(_np(pts))
        return torch.from_numpy(rgb), torch.from_numpy(sig)

    aabb_np = [np.array([-1.25, -1.55, -1.25], f32), np.array([1.25, 0.95, 1.25], f32)]
    grid = DensityGrid(64, [torch.from_numpy(a) for a in aabb_np], device="cpu")
    cached = np.zeros((64, 64, 64), f32); old = np.zeros((64, 64, 64), bool)
    for step in (0, 600):   # `valid` is the new field before step 500 and the previous one afterwards
        jit = rng.random((64, 64, 64, 3), dtype=f32)
        dens, valid = grid.update(deformer, None, step, torch.from_numpy(jit))
        ref_reg, ref_valid, cached, field = orender.density_grid_update(lambda p: blob_np(p), aabb_np, jit, cached, old, step)
        np.testing.assert_allclose(_np(dens), ref_reg, atol=1e-6)
        np.testing.assert_array_equal(_np(valid), ref_valid)
        np.testing.assert_array_equal(_np(grid.density_field), field)
        np.testing.assert_allclose(_np(grid.density_cached), cached, atol=1e-6)
        old = field
        assert field.sum() > 50
    # generic initialize (foreign deformer): max over the jitter passes, then the same post-processing
    class Foreign:
        def get_bbox_deformed(self):
            return [torch.from_numpy(a) for a in aabb_np]
        def __call__(self, pts, net, eval_mode=True):
            rgb, sig = blob_np(_np(pts))
            return torch.from_numpy(rgb), torch.from_numpy(np.maximum(sig, 0))   # eval mode: invalid -> 0
    g2 = DensityGrid(64, device="cpu")
    jits = rng.random((5, 64, 64, 64, 3), dtype=f32)
    g2.initialize(Foreign(), None, jitters=torch.from_numpy(jits))
    ref_field, _ = orender.density_grid_initialize(lambda p: (blob_np(p)[0], np.maximum(blob_np(p)[1], 0)), aabb_np, jits, 64)
    np.testing.assert_array_equal(_np(g2.density_field), ref_field)
