"""ORACLE (test infrastructure): per-subject initialisation and per-frame preparation.

Restates /root/reference/instant_avatar/deformers/snarf_deformer.py:6-31,41-107 and
deformers/fast_snarf/deformer_torch.py:130-202,225-244 in numpy float32.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

from . import capi
from .smpl_np import SMPLNumpy

f32 = np.float32
INIT_BONES = [0, 1, 2, 4, 5, 10, 11, 12, 15, 16, 17, 18, 19]  # deformer_torch.py:28


def rest_pose_a():  # snarf_deformer.py:11-15
    bp = np.zeros((1, 69), f32)
    bp[0, 2] = 0.2; bp[0, 5] = -0.2; bp[0, 47] = -0.8; bp[0, 50] = 0.8
    return bp


def bbox_from_smpl(vs: np.ndarray, factor=1.2):  # snarf_deformer.py:20-31
    mn, mx = vs.min(0), vs.max(0)
    c = (mx + mn) / f32(2)
    s = ((mx - mn) / f32(2)).max() * f32(factor)
    return np.stack([c - s, c + s]).astype(f32)


def query_weights_smpl(x: np.ndarray, verts: np.ndarray, weights: np.ndarray, resolution=128):
    """deformer_torch.py:225-244: KNN-30 inverse-distance blend + 30 Laplacian smoothing passes.
    x [Vox,3] (d,h,w raster order), returns [24,d,h,w]."""
    tree = cKDTree(verts.astype(np.float64))
    dist, idx = tree.query(x.astype(np.float64), k=30, workers=-1)
    # the reference gets SQUARED float32 distances from knn_points, then sqrt + clamp
    diff = x[:, None, :].astype(f32) - verts[idx].astype(f32)
    d2 = (diff * diff).sum(-1).astype(f32)
    dist = np.clip(np.sqrt(d2), f32(1e-4), f32(1.0)).astype(f32)
    w = weights[idx].astype(f32)  # [Vox,30,24]
    ws = f32(1.0) / dist
    ws = ws / ws.sum(-1, keepdims=True)
    w = (ws[..., None] * w).sum(-2).astype(f32)  # [Vox,24]
    d, h, ww = resolution // 4, resolution, resolution
    w = np.ascontiguousarray(w.T).reshape(1, 24, d, h, ww)
    for _ in range(30):
        mean = (w[:, :, 2:, 1:-1, 1:-1] + w[:, :, :-2, 1:-1, 1:-1] + w[:, :, 1:-1, 2:, 1:-1]
                + w[:, :, 1:-1, :-2, 1:-1] + w[:, :, 1:-1, 1:-1, 2:] + w[:, :, 1:-1, 1:-1, :-2]) / f32(6.0)
        w[:, :, 1:-1, 1:-1, 1:-1] = (w[:, :, 1:-1, 1:-1, 1:-1] - mean) * f32(0.7) + mean
        w = (w / w.sum(1, keepdims=True)).astype(f32)
    return w[0]


class SubjectOracle:
    """State built once per subject (SNARFDeformer.initialize, snarf_deformer.py:41-69)."""

    def __init__(self, smpl_data: dict, betas: np.ndarray, resolution: int = 128, lbs_voxel: np.ndarray | None = None):
        self.smpl = SMPLNumpy(smpl_data)
        self.betas = np.asarray(betas, f32).reshape(1, 10)
        out = self.smpl.forward(self.betas, rest_pose_a())
        self.A_cano = out["A"]
        self.tfs_inv_t = np.linalg.inv(out["A"].astype(f32)).astype(f32)  # :52
        self.verts_cano = out["vertices"]
        self.joints_cano = out["joints"]
        self.bbox = bbox_from_smpl(out["vertices"])  # :61
        # switch_to_explicit, deformer_torch.py:130-158
        self.res = resolution
        d, h, w = resolution // 4, resolution, resolution
        self.dhw = (d, h, w)
        ratio = h / d
        gt_min, gt_max = out["vertices"].min(0), out["vertices"].max(0)
        offset = ((gt_min + gt_max) * f32(0.5)).astype(f32)
        scale = f32((gt_max - gt_min).max() / f32(2) * f32(1.2))
        self.offset, self.scale, self.ratio = offset, scale, ratio
        self.offset_kernel = (-offset).astype(f32)
        sk = np.full(3, f32(1.0) / scale, f32)
        sk[2] = sk[2] * f32(ratio)
        self.scale_kernel = sk
        if lbs_voxel is None:
            xr = np.linspace(-1, 1, w, dtype=f32); yr = np.linspace(-1, 1, h, dtype=f32); zr = np.linspace(-1, 1, d, dtype=f32)
            gz, gy, gx = np.meshgrid(zr, yr, xr, indexing="ij")
            grid = np.stack([gx, gy, gz], -1).reshape(-1, 3).astype(f32)
            # denormalize (:166-171)
            g = grid.copy(); g[:, 2] /= f32(ratio); g *= scale; g += offset
            lbs_voxel = query_weights_smpl(g, out["vertices"], self.smpl.lbs_weights, resolution)
        self.lbs_voxel = np.ascontiguousarray(lbs_voxel, f32)  # [24,d,h,w]

    def prepare_frame(self, pose: dict, tfs: np.ndarray | None = None, w2s: np.ndarray | None = None):
        """snarf_deformer.py:71-93 -> dict(tfs, w2s, voxel_d, voxel_J, vertices, bbox_deformed).
        `tfs` / `w2s` given: continue from those bone transforms instead of this file's SMPL algebra -- the public-API
        parity tests hand over the product's transforms (equal to the ones computed here to ~1e-6: another summation
        order of the same kinematic chain) so that everything downstream is compared on identical inputs."""
        out = self.smpl.forward(self.betas, pose["body_pose"], pose["global_orient"], pose["transl"])
        A = out["A"].astype(f32)
        s2w = A[0]
        if w2s is None:
            w2s = np.linalg.inv(s2w).astype(f32)
        w2s = np.asarray(w2s, f32).reshape(4, 4)
        if tfs is None:
            tfs = (w2s[None] @ A @ self.tfs_inv_t).astype(f32)  # :86
        tfs = np.ascontiguousarray(np.asarray(tfs, f32).reshape(24, 4, 4))
        d, h, w = self.dhw
        voxel_d, voxel_J = capi.precompute(self.lbs_voxel, tfs, self.offset_kernel, self.scale_kernel, d, h, w)
        verts = out["vertices"] @ w2s[:3, :3].T + w2s[:3, 3]
        vd = voxel_d.reshape(3, -1)
        return {
            "tfs": tfs, "w2s": w2s, "voxel_d": voxel_d, "voxel_J": voxel_J, "vertices": verts.astype(f32),
            "bbox_deformed": np.stack([vd.min(1), vd.max(1)]).astype(f32),  # :105-107
        }


def transform_rays_w2s(o: np.ndarray, d: np.ndarray, w2s: np.ndarray):
    """snarf_deformer.py:95-103"""
    o2 = (o @ w2s[:3, :3].T + w2s[:3, 3]).astype(f32)
    d2 = (d @ w2s[:3, :3].T).astype(f32)
    dist = np.linalg.norm(o2, axis=-1).astype(f32)
    return o2, d2, dist - f32(1), dist + f32(1)
