"""Host-side mirror of instant_avatar/deformers/smpl_deformer.py::SMPLDeformer (`fit.py deformer=smpl`,
bash/run-neuman-demo.sh): every sample point takes the inverse skinning transform of its nearest posed SMPL vertex.

The nearest-vertex search (pytorch3d `knn_points`, K = 1, per sample per frame) is `ia_knn1`; the per-vertex inverse
transforms are assembled with torch exactly as the reference does (`:60-76`), once per frame.  The network is called
through `model(pts_cano, None)`; `NeRFNGPNet.forward` is differentiable w.r.t. parameters and points, so the pose
gradients of `optimize_SMPL` flow through `T_inv`.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .smpl import SMPL
from .snarf_deformer import get_bbox_from_smpl


class SMPLDeformer:
    def __init__(self, model_path=None, gender="male", threshold=0.05, k=1, smpl_data=None) -> None:
        if k != 1:
            raise ValueError("SMPLDeformer: the nearest-neighbour strategy uses k = 1 (smpl_deformer.py:26,102-103)")
        self.body_model = SMPL(model_path, gender=gender, data_struct=smpl_data)
        self.k = k
        self.threshold = threshold
        self.strategy = "nearest_neighbor"
        self.initialized = False

    def initialize(self, betas, device):
        """smpl_deformer.py:33-45: the canonical template is a 30-degree A-pose"""
        batch_size = betas.shape[0]
        body_pose_t = torch.zeros((batch_size, 69), device=device)
        body_pose_t[:, 2] = math.pi / 6
        body_pose_t[:, 5] = -math.pi / 6
        out = self.body_model(betas=betas, body_pose=body_pose_t)
        self.bbox = get_bbox_from_smpl(out.vertices[0:1].detach())
        self.T_template = out.T
        self.vs_template = out.vertices
        self.pose_offset_t = out.pose_offsets
        self.shape_offset_t = out.shape_offsets

    def get_bbox_deformed(self):
        return get_bbox_from_smpl(self.vertices[0:1].detach())

    def prepare_deformer(self, smpl_params):
        """smpl_deformer.py:50-76"""
        device = smpl_params["betas"].device
        if self.body_model.v_template.device != device:
            self.body_model = self.body_model.to(device)
        if not self.initialized:
            self.initialize(smpl_params["betas"], device)  # every frame, as in the reference (betas may change)
        out = self.body_model(betas=smpl_params["betas"], body_pose=smpl_params["body_pose"],
                              global_orient=smpl_params["global_orient"], transl=smpl_params["transl"])
        s2w = out.A[:, 0]
        w2s = torch.inverse(s2w)
        # remove and re-apply the blend shapes: posed -> T pose -> template pose
        T_inv = torch.inverse(out.T.float()).clone() @ s2w[:, None]
        T_inv[..., :3, 3] += self.pose_offset_t - out.pose_offsets
        T_inv[..., :3, 3] += self.shape_offset_t - out.shape_offsets
        T_inv = self.T_template @ T_inv
        self.T_inv = T_inv
        self.vertices = (out.vertices @ w2s[:, :3, :3].permute(0, 2, 1)) + w2s[:, None, :3, 3]
        self.w2s = w2s

    def transform_rays_w2s(self, rays):
        """smpl_deformer.py:78-85"""
        w2s = self.w2s
        rays.o = (rays.o @ w2s[:, :3, :3].permute(0, 2, 1)) + w2s[:, None, :3, 3]
        rays.d = (rays.d @ w2s[:, :3, :3].permute(0, 2, 1)).to(rays.d)
        d = torch.norm(rays.o, dim=-1)
        rays.near = d - 1
        rays.far = d + 1

    def deform(self, pts):
        """smpl_deformer.py:87-110: canonical point = T_inv[nearest vertex] . pts, valid if the vertex is within `threshold`"""
        batch_size = self.vertices.shape[0]
        pts = pts.reshape(batch_size, -1, 3)
        pts_cano = torch.zeros_like(pts, dtype=torch.float32)
        valid = torch.zeros(pts.shape[:2], device=pts.device, dtype=torch.bool)
        for i in range(batch_size):
            with torch.no_grad():
                dist_sq, idx = ops.knn1(pts[i].detach().float(), self.vertices[i].detach().float())
            valid[i] = dist_sq < self.threshold ** 2
            Tv_inv = self.T_inv[i][idx]
            pts_cano[i] = (Tv_inv[..., :3, :3] @ pts[i][..., None]).squeeze(-1) + Tv_inv[..., :3, 3]
        return pts_cano.reshape(-1, 3), valid.reshape(-1)

    def deform_train(self, pts, model):
        """smpl_deformer.py:112-123"""
        pts_cano, valid = self.deform(pts)
        rgb_cano = torch.zeros_like(pts, dtype=torch.float32)
        sigma_cano = torch.ones_like(pts[..., 0]) * -1e5
        if valid.any():
            r, s = model(pts_cano[valid], None)
            rgb_cano = rgb_cano.index_put((valid.nonzero(as_tuple=True)[0],), r.float())
            sigma_cano = sigma_cano.index_put((valid.nonzero(as_tuple=True)[0],), s.float())
            finite = torch.isfinite(rgb_cano).all(-1) & torch.isfinite(sigma_cano)
            rgb_cano = torch.where(finite[:, None], rgb_cano, torch.zeros_like(rgb_cano))
            sigma_cano = torch.where(finite, sigma_cano, torch.full_like(sigma_cano, -1e5))
        return rgb_cano, sigma_cano

    def deform_test(self, pts, model):
        """smpl_deformer.py:125-132"""
        pts_cano, valid = self.deform(pts)
        rgb_cano = torch.zeros_like(pts, dtype=torch.float32)
        sigma_cano = torch.zeros_like(pts[..., 0])
        if valid.any():
            r, s = model(pts_cano[valid], None)
            rgb_cano[valid], sigma_cano[valid] = r.float(), s.float()
        return rgb_cano, sigma_cano

    def __call__(self, pts, model, eval_mode=True):
        pts = pts.reshape(-1, 3)
        if eval_mode:
            return self.deform_test(pts, model)
        return self.deform_train(pts, model)
