"""CPU: the oracle's pose-gradient restatement (oracle/torch_ref.pose_grad_reference, deformer_torch.py:50-67) against
finite differences of the oracle's own forward -- the implicit-function gradient d x_c / d tfs = -J^-1 dLBS/dtfs must
reproduce what re-solving the roots for perturbed bone transforms gives.  fp32 network (no fp16 emulation) so that the
differences are not dominated by rounding steps."""
import numpy as np
import torch

from oracle import capi
from oracle import frame as oframe
from oracle import testing as scene_util
from oracle import torch_ref


def test_implicit_pose_gradient_matches_finite_differences_of_the_root_solver():
    sc = scene_util.oracle_scene(0)
    subj, fr, net = sc["subj"], sc["frame"], sc["net"]
    d, h, w = subj.dhw
    rng = np.random.default_rng(3)
    n = 400
    # posed points inside the body: canonical surface points pulled inwards and skinned with the voxelised weights
    xc0 = (subj.verts_cano[rng.integers(0, len(subj.verts_cano), n)] * 0.96).astype(np.float32)
    lbs = torch.from_numpy(subj.lbs_voxel).reshape(1, 24, d, h, w)
    off, scl = torch.from_numpy(subj.offset_kernel).reshape(3), torch.from_numpy(subj.scale_kernel).reshape(3)
    q = (scl * (torch.from_numpy(xc0) + off)).reshape(1, 1, 1, -1, 3)
    wts = torch.nn.functional.grid_sample(lbs, q, align_corners=True, padding_mode="border").reshape(24, -1).T
    tfs0 = torch.from_numpy(fr["tfs"]).reshape(24, 4, 4)
    xd = torch.einsum("pn,nij,pj->pi", wts, tfs0, torch.cat([torch.from_numpy(xc0), torch.ones(n, 1)], 1))[:, :3].numpy()
    g_sigma = rng.normal(0, 1, n).astype(np.float32)
    g_rgb = rng.normal(0, 1, (n, 3)).astype(np.float32)
    # a smooth network: only the three coarsest hash levels (7-15 cm cells) carry features, so that a finite difference
    # over millimetres measures the same slope as the analytic derivative (the scene's avatar has sub-millimetre detail)
    lay = capi.hashgrid_layout()
    lim = lambda fi, fo: np.sqrt(6.0 / (fi + fo))
    mk = lambda o, i: rng.uniform(-lim(i, o), lim(i, o), (o, i)).astype(np.float32).ravel()
    grid = np.zeros((int(lay["total"]), 2), np.float32)
    coarse_end = int(lay["offset"][3])
    grid[:coarse_end] = rng.uniform(-1, 1, (coarse_end, 2)).astype(np.float32)
    enc = torch.from_numpy(np.concatenate([mk(64, 32), mk(16, 64), grid.ravel()]))
    col = torch.from_numpy(np.concatenate([mk(64, 16), mk(64, 64), mk(16, 64)]))

    def solve(tfs):
        _, vJ = capi.precompute(subj.lbs_voxel, tfs, subj.offset_kernel, subj.scale_kernel, d, h, w)
        return capi.broyden(xd, vJ, tfs, oframe.INIT_BONES, subj.offset_kernel, subj.scale_kernel)

    xc, jinv, valid, _ = solve(fr["tfs"])
    best = np.where(valid.any(1), valid.argmax(1), -1)           # first valid initialisation of every point
    assert (best >= 0).mean() > 0.9
    sel = best >= 0

    def loss(tfs):
        xc_p, _, valid_p, _ = solve(tfs)
        x = torch.from_numpy(xc_p[np.arange(n), np.maximum(best, 0)][sel])
        s, c = torch_ref.ngp_forward(x, net.center, net.scale, enc, col, emulate=False)
        return float((s * torch.from_numpy(g_sigma[sel])).sum() + (c * torch.from_numpy(g_rgb[sel])).sum())

    grad = torch_ref.pose_grad_reference(torch.from_numpy(xd), torch.from_numpy(best), torch.from_numpy(xc), torch.from_numpy(jinv), lbs, off,
                                         scl, tfs0, net.center, net.scale, enc, col, torch.from_numpy(g_sigma), torch.from_numpy(g_rgb),
                                         emulate=False).numpy()
    assert np.all(grad[:, 3] == 0)
    # finite differences on the entries with the largest analytic gradient (translations and rotation entries alike)
    flat = np.argsort(-np.abs(grad[:, :3]).ravel())[:8]
    eps = 2e-3
    fd, an = [], []
    for f in flat:
        j, r, c = np.unravel_index(f, (24, 3, 4))
        tp, tm = fr["tfs"].copy(), fr["tfs"].copy()
        tp[j, r, c] += eps; tm[j, r, c] -= eps
        fd.append((loss(tp) - loss(tm)) / (2 * eps)); an.append(grad[j, r, c])
    fd, an = np.array(fd), np.array(an)
    cos = float(fd @ an / (np.linalg.norm(fd) * np.linalg.norm(an)))
    ratio = float(np.linalg.norm(an) / np.linalg.norm(fd))
    assert cos > 0.95, (cos, an, fd)
    assert 0.7 < ratio < 1.4, (ratio, an, fd)
