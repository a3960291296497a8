"""ORACLE (test infrastructure): differentiable plain-PyTorch fp32 restatement of the network and of the training
compositing, used as the autograd reference for the CUDA backward kernels (CPU).

 * hash grid + MLPs: same specification as oracle/ia_oracle.c::orc_ngp_forward (tiny-cuda-nn v1.6 semantics,
   PARITY UNPINNED -- see that file); fp16 roundings are applied with a straight-through estimator, which is the
   gradient definition the CUDA backward implements.
 * compositing: instant_avatar/renderers/raymarcher_acc.py:25-36,161-186, loss: utils/loss.py:53-79.
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi


def ste_half(x: torch.Tensor, emulate=True) -> torch.Tensor:
    if not emulate:
        return x
    return x + (x.half().float() - x).detach()


def hash_encode(x01: torch.Tensor, grid: torch.Tensor, emulate=True, pos_grad=False) -> torch.Tensor:
    """x01 [P,3] in [0,1]; grid [total,2] fp32 master table -> [P,32].  pos_grad: keep the interpolation weights in the
    autograd graph (tiny-cuda-nn's input gradient, needed by the pose-gradient test)."""
    lay = capi.hashgrid_layout()
    feats = []
    gtab = ste_half(grid, emulate)
    for l in range(16):
        s = float(lay["scale"][l]); res = int(lay["res"][l]); size = int(lay["size"][l]); off = int(lay["offset"][l])
        pos = (x01 * np.float32(s) + np.float32(0.5)).float()
        fl = torch.floor(pos)
        w = pos - fl
        c0 = fl.long()
        acc = torch.zeros((x01.shape[0], 2), dtype=torch.float32, device=x01.device)
        for k in range(8):
            dx, dy, dz = k & 1, (k >> 1) & 1, (k >> 2) & 1
            cx, cy, cz = c0[:, 0] + dx, c0[:, 1] + dy, c0[:, 2] + dz
            wx = w[:, 0] if dx else 1 - w[:, 0]
            wy = w[:, 1] if dy else 1 - w[:, 1]
            wz = w[:, 2] if dz else 1 - w[:, 2]
            wt = (wx * wy * wz) if pos_grad else (wx * wy * wz).detach()
            stride, idx, hashed = 1, torch.zeros_like(cx), False
            for comp in (cx, cy, cz):
                if stride <= size:
                    idx = idx + comp * stride
                    stride *= res
            if size < stride:
                idx = ((cx * 1) ^ ((cy * 2654435761) & 0xFFFFFFFF) ^ ((cz * 805459861) & 0xFFFFFFFF)) & 0xFFFFFFFF
            idx = idx % size
            acc = acc + wt[:, None] * gtab[off + idx]
        feats.append(ste_half(acc, emulate))
    return torch.cat(feats, dim=1)


def ngp_forward(x: torch.Tensor, center, scale, enc_params: torch.Tensor, col_params: torch.Tensor, emulate=True, pos_grad=False):
    """ngp.py:73-83 -> (sigma [P], rgb [P,3])"""
    W1 = ste_half(enc_params[:2048].view(64, 32), emulate); W2 = ste_half(enc_params[2048:3072].view(16, 64), emulate)
    grid = enc_params[3072:].view(-1, 2)
    W3 = ste_half(col_params[:1024].view(64, 16), emulate); W4 = ste_half(col_params[1024:5120].view(64, 64), emulate)
    W5 = ste_half(col_params[5120:].view(16, 64), emulate)
    xn = ((x - torch.as_tensor(center)) / torch.as_tensor(scale) + 0.5).clamp(0, 1)
    enc = hash_encode(xn, grid, emulate, pos_grad)
    h1 = ste_half(torch.relu(enc @ W1.T), emulate)
    o16 = ste_half(h1 @ W2.T, emulate)
    sigma = o16[:, 0]
    cin = torch.cat([o16[:, 1:], torch.ones_like(o16[:, :1])], dim=1)
    h2 = ste_half(torch.relu(cin @ W3.T), emulate)
    h3 = ste_half(torch.relu(h2 @ W4.T), emulate)
    o3 = h3 @ W5.T
    rgb = ste_half(torch.sigmoid(o3[:, :3]), emulate)
    return sigma, rgb


def composite_train(sigma_vals, dists):
    """raymarcher_acc.py:25-36"""
    tau = torch.relu(sigma_vals) * dists
    alpha = 1.0 - torch.exp(-tau)
    trans = torch.cat([torch.ones_like(alpha[..., 0:1]), torch.cumprod(1 - alpha + 1e-10, dim=-1)], dim=-1)
    return alpha * trans[..., :-1], trans


def nerf_loss(rgb, alpha, weights, target_rgb, target_alpha, w_rgb=1.0, w_alpha=0.1, w_reg=0.1):
    """utils/loss.py:58-79"""
    OFFSET = 0.313262
    loss = w_rgb * torch.mean((rgb - target_rgb) ** 2) + w_alpha * torch.mean((alpha - target_alpha) ** 2)
    reg = lambda x: (-torch.log(torch.exp(-x) + torch.exp(x - 1))).mean() + OFFSET
    return loss + w_reg * reg(alpha) + w_reg * reg(weights)


def render_train_torch(out_np: dict, net, enc_params: torch.Tensor, col_params: torch.Tensor, step_size, bg, noise=None):
    """Differentiable tail of render_train given the oracle's numpy forward (`oracle.render.render_train(...,
    return_aux=True)`): network at the arg-max canonical points -> compositing.  Returns dict of torch outputs."""
    mask = out_np["mask"]
    aux = out_np["aux"]
    N, S = mask.shape
    sigma = torch.full((N, S), -1e3, dtype=torch.float32)
    rgbv = torch.zeros((N, S, 3), dtype=torch.float32)
    valid_any = np.take_along_axis(aux["valid"], aux["idx"][:, None], 1)[:, 0]  # arg-max candidate is a real root
    xc_best = aux["xc"][np.arange(len(aux["idx"])), aux["idx"]]
    s_flat = torch.full((int(mask.sum()),), -1e5, dtype=torch.float32)
    c_flat = torch.zeros((int(mask.sum()), 3), dtype=torch.float32)
    if valid_any.any():
        s_v, c_v = ngp_forward(torch.from_numpy(xc_best[valid_any]), net.center, net.scale, enc_params, col_params, net.emulate)
        vi = torch.from_numpy(np.nonzero(valid_any)[0])
        s_flat = s_flat.index_put((vi,), s_v)
        c_flat = c_flat.index_put((vi,), c_v)
    mi = torch.from_numpy(np.stack(np.nonzero(mask), 1))
    sigma = sigma.index_put((mi[:, 0], mi[:, 1]), s_flat)
    rgbv = rgbv.index_put((mi[:, 0], mi[:, 1]), c_flat)
    if noise is not None:
        sigma = sigma + torch.as_tensor(noise)
    dists = torch.ones_like(sigma) * torch.as_tensor(step_size)[:, None]
    w, trans = composite_train(sigma, dists)
    no_hit = trans[..., -1]
    color = (w[..., None] * rgbv).sum(-2) + no_hit[:, None] * torch.as_tensor(bg)
    depth = (w * torch.from_numpy(out_np["z"])).sum(-1)
    return {"rgb": color, "depth": depth, "alpha": w.sum(-1), "weights": w, "n_net": int(valid_any.sum())}


def pose_grad_reference(xd, best, xc_opt, j_inv, lbs_voxel, offset_k, scale_k, tfs, center, scale, enc_params, col_params,
                        g_sigma, g_rgb, emulate=True, g_xc=None):
    """d loss / d tfs of the selected roots, restating deformers/fast_snarf/deformer_torch.py:50-67 (version 1) literally:
    x_c = x_c*.detach() + bmv(-J_inv, x_d* - x_d*.detach()),  x_d* = forward_skinning(x_c*, tfs) (:171-188) with weights
    from F.grid_sample(lbs_voxel, scale*(x_c+offset), align_corners=True, padding_mode='border') (:190-201), followed by
    the network (ngp.py:73-83) whose positional gradient is tiny-cuda-nn's.  All inputs CPU torch tensors; `best` [P]
    selects the initialisation per point (-1: none).  Returns grad [24,4,4] for loss = sum(g_sigma*sigma + g_rgb*rgb).
    g_xc [P,3] (optional): d loss / d x_c handed over instead of differentiating the network here -- isolates the
    implicit-differentiation algebra (J_inv, skinning weights, outer products) from the network's fp16 gradient chain."""
    import torch.nn.functional as F
    sel = best >= 0
    idx = best.clamp(min=0).long()
    P = xd.shape[0]
    xc0 = xc_opt[torch.arange(P), idx][sel].detach()
    Ji = j_inv[torch.arange(P), idx][sel].detach()
    tfs = tfs.detach().clone().requires_grad_(True)
    q = (scale_k.reshape(1, 3) * (xc0 + offset_k.reshape(1, 3))).reshape(1, 1, 1, -1, 3)
    w = F.grid_sample(lbs_voxel.reshape(1, 24, *lbs_voxel.shape[-3:]), q, align_corners=True, mode="bilinear", padding_mode="border")
    w = w.reshape(24, -1).T                                           # [Q,24]
    xh = torch.cat([xc0, torch.ones_like(xc0[:, :1])], dim=1)           # [Q,4]
    xd_opt = torch.einsum("pn,nij,pj->pi", w, tfs.reshape(24, 4, 4), xh)[:, :3]
    corr = torch.einsum("pij,pj->pi", -Ji, xd_opt - xd_opt.detach())
    xc = xc0 + corr
    if g_xc is not None:
        loss = (xc * g_xc[sel].detach()).sum()
    else:
        sigma, rgb = ngp_forward(xc, center, scale, enc_params, col_params, emulate, pos_grad=True)
        loss = (sigma * g_sigma[sel]).sum() + (rgb * g_rgb[sel]).sum()
    loss.backward()
    return tfs.grad.reshape(24, 4, 4)
