"""SMPL body model forward (shape + pose blend shapes, 24-joint kinematic chain, linear blend skinning).

Host-side mirror of the reference's vendored smplx subset
(instant_avatar/deformers/smplx/body_models.py:289-372, lbs.py:152-248,295-329,345-401), including the fork's
behaviour of adding `transl` into the bone transforms A.  Written batch-1-first with the kinematic chain
evaluated level-by-level (7 batched matmuls instead of 23 sequential ones); PyTorch ops only -- this is per-frame
plumbing that produces the 24 bone transforms, not the per-ray hot path.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn


@dataclass
class SMPLOutput:
    vertices: torch.Tensor
    joints: torch.Tensor
    A: torch.Tensor
    T: torch.Tensor
    betas: torch.Tensor
    body_pose: torch.Tensor
    global_orient: torch.Tensor
    pose_offsets: torch.Tensor = None    # [B,V,3] pose blend-shape displacements (the vendored smplx exposes them)
    shape_offsets: torch.Tensor = None   # [B,V,3] shape blend-shape displacements


def _chain_levels(parents: np.ndarray):
    depth = np.zeros(len(parents), np.int64)
    for i in range(1, len(parents)):
        depth[i] = depth[parents[i]] + 1
    return [np.nonzero(depth == d)[0] for d in range(1, depth.max() + 1)]


class SMPL(nn.Module):
    NUM_BODY_JOINTS = 23

    def __init__(self, model_path=None, gender="neutral", data_struct: dict | None = None, dtype=torch.float32):
        super().__init__()
        if data_struct is None:
            import os
            path = model_path
            if os.path.isdir(model_path):
                path = os.path.join(model_path, f"SMPL_{gender.upper()}.pkl")
            if not os.path.exists(path):
                raise FileNotFoundError(f"Path {path} does not exist!")
            with open(path, "rb") as f:
                data_struct = pickle.load(f, encoding="latin1")
        g = lambda k: np.asarray(data_struct[k])
        self.gender = gender
        self.register_buffer("v_template", torch.tensor(g("v_template"), dtype=dtype))
        self.register_buffer("shapedirs", torch.tensor(g("shapedirs")[:, :, :10], dtype=dtype))
        pd = g("posedirs")
        self.register_buffer("posedirs", torch.tensor(pd.reshape(-1, pd.shape[-1]).T.copy(), dtype=dtype))
        jr = data_struct["J_regressor"]
        jr = jr.toarray() if hasattr(jr, "toarray") else np.asarray(jr)
        self.register_buffer("J_regressor", torch.tensor(jr, dtype=dtype))
        parents = g("kintree_table")[0].astype(np.int64).copy()
        parents[0] = -1
        self.register_buffer("parents", torch.tensor(parents))
        self.register_buffer("lbs_weights", torch.tensor(g("weights"), dtype=dtype))
        self.register_buffer("faces_tensor", torch.tensor(g("f").astype(np.int64)))
        # kinematic-chain levels as device index tensors (no per-frame host->device index uploads)
        self._n_levels = 0
        for lv in _chain_levels(parents):
            self.register_buffer(f"_lvl_idx{self._n_levels}", torch.tensor(lv), persistent=False)
            self.register_buffer(f"_lvl_par{self._n_levels}", torch.tensor(parents[lv]), persistent=False)
            self._n_levels += 1
        self.register_buffer("_par1", torch.tensor(parents[1:]), persistent=False)

    @staticmethod
    def rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
        angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
        d = rot_vecs / angle
        cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
        rx, ry, rz = d[:, 0], d[:, 1], d[:, 2]
        z = torch.zeros_like(rx)
        K = torch.stack([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
        ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
        return ident + sin * K + (1 - cos) * torch.bmm(K, K)

    def forward(self, betas, body_pose, global_orient=None, transl=None) -> SMPLOutput:
        B = max(betas.shape[0], body_pose.shape[0])
        dev, dt = betas.device, self.v_template.dtype
        if global_orient is None:
            global_orient = torch.zeros((B, 3), device=dev, dtype=dt)
        betas = betas.expand(B, -1)
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        shape_offsets = torch.einsum("bl,mkl->bmk", betas, self.shapedirs)
        v_shaped = self.v_template + shape_offsets
        J = torch.einsum("jv,bvk->bjk", self.J_regressor, v_shaped)
        rot = self.rodrigues(full_pose.reshape(-1, 3)).view(B, 24, 3, 3)
        ident = torch.eye(3, dtype=dt, device=dev)
        pose_feature = (rot[:, 1:] - ident).reshape(B, -1)
        pose_offsets = torch.matmul(pose_feature, self.posedirs).view(B, -1, 3)
        v_posed = v_shaped + pose_offsets
        # kinematic chain, level by level
        rel = J.clone()
        rel[:, 1:] = J[:, 1:] - J[:, self._par1]
        tm = torch.zeros((B, 24, 4, 4), device=dev, dtype=dt)
        tm[:, :, :3, :3] = rot
        tm[:, :, :3, 3] = rel
        tm[:, :, 3, 3] = 1
        chain = torch.empty_like(tm)
        chain[:, 0] = tm[:, 0]
        for l in range(self._n_levels):
            idx, par = getattr(self, f"_lvl_idx{l}"), getattr(self, f"_lvl_par{l}")
            chain[:, idx] = torch.matmul(chain[:, par], tm[:, idx])
        posed_joints = chain[:, :, :3, 3]
        jh = torch.cat([J, torch.zeros((B, 24, 1), device=dev, dtype=dt)], dim=2)[..., None]
        A = chain.clone()
        A[:, :, :, 3] = A[:, :, :, 3] - torch.matmul(chain, jh)[..., 0]
        T = torch.matmul(self.lbs_weights[None], A.view(B, 24, 16)).view(B, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones((B, v_posed.shape[1], 1), device=dev, dtype=dt)], dim=2)
        verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
        if transl is not None:
            verts = verts + transl[:, None]
            posed_joints = posed_joints + transl[:, None]
            A = A.clone(); A[..., :3, 3] += transl[:, None]
            T = T.clone(); T[..., :3, 3] += transl[:, None]
        return SMPLOutput(vertices=verts, joints=posed_joints, A=A, T=T, betas=betas, body_pose=body_pose,
                          global_orient=global_orient, pose_offsets=pose_offsets, shape_offsets=shape_offsets)
