"""ORACLE (test infrastructure): numpy float32 restatement of the reference's SMPL forward.

Follows /root/reference/instant_avatar/deformers/smplx/lbs.py:152-248 (lbs), :295-329
(batch_rodrigues), :345-401 (batch_rigid_transform) and body_models.py:289-372 (SMPL.forward,
including the fork's addition of `transl` into A and T at :353-360).  Pinned against the reference
module itself imported on CPU (tests/golden/make_smpl_golden.py -> tests/golden/smpl_golden.npz).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


class SMPLNumpy:
    def __init__(self, data: dict):
        self.v_template = np.asarray(data["v_template"], f32)  # [V,3]
        self.shapedirs = np.asarray(data["shapedirs"], f32)[:, :, :10]  # [V,3,10]
        pd = np.asarray(data["posedirs"], f32)
        self.posedirs = pd.reshape(-1, pd.shape[-1]).T.copy()  # [207, V*3]  (body_models.py:243-246)
        self.J_regressor = np.asarray(data["J_regressor"], f32)  # [24,V]
        parents = np.asarray(data["kintree_table"][0]).astype(np.int64)
        parents[0] = -1
        self.parents = parents
        self.lbs_weights = np.asarray(data["weights"], f32)  # [V,24]

    @staticmethod
    def rodrigues(rot_vecs: np.ndarray) -> np.ndarray:
        """lbs.py:295-329"""
        rv = rot_vecs.astype(f32)
        angle = np.linalg.norm(rv + f32(1e-8), axis=1, keepdims=True).astype(f32)
        d = (rv / angle).astype(f32)
        cos = np.cos(angle)[:, None].astype(f32)
        sin = np.sin(angle)[:, None].astype(f32)
        rx, ry, rz = d[:, 0], d[:, 1], d[:, 2]
        z = np.zeros_like(rx)
        K = np.stack([z, -rz, ry, rz, z, -rx, -ry, rx, z], axis=1).reshape(-1, 3, 3).astype(f32)
        ident = np.eye(3, dtype=f32)[None]
        return (ident + sin * K + (f32(1) - cos) * (K @ K)).astype(f32)

    def forward(self, betas, body_pose, global_orient=None, transl=None):
        """Returns dict(vertices [V,3], A [24,4,4], T [V,4,4], joints [24,3]) for batch size 1."""
        betas = np.asarray(betas, f32).reshape(1, -1)
        body_pose = np.asarray(body_pose, f32).reshape(1, -1)
        if global_orient is None:
            global_orient = np.zeros((1, 3), f32)
        global_orient = np.asarray(global_orient, f32).reshape(1, 3)
        full_pose = np.concatenate([global_orient, body_pose], axis=1)
        # lbs.py:204-206 shape blend shapes
        v_shaped = self.v_template + np.einsum("bl,mkl->bmk", betas, self.shapedirs)[0].astype(f32)
        # lbs.py:210 joints
        J = (self.J_regressor @ v_shaped).astype(f32)  # [24,3]
        rot = self.rodrigues(full_pose.reshape(-1, 3))  # [24,3,3]
        # lbs.py:218-222 pose blend shapes
        pose_feature = (rot[1:] - np.eye(3, dtype=f32)).reshape(1, -1)
        pose_offsets = (pose_feature @ self.posedirs).reshape(-1, 3).astype(f32)
        v_posed = pose_offsets + v_shaped
        # lbs.py:345-401 kinematic chain
        rel = J.copy()
        rel[1:] -= J[self.parents[1:]]
        tm = np.zeros((24, 4, 4), f32)
        tm[:, :3, :3] = rot
        tm[:, :3, 3] = rel
        tm[:, 3, 3] = 1
        chain = [tm[0]]
        for i in range(1, 24):
            chain.append((chain[self.parents[i]] @ tm[i]).astype(f32))
        transforms = np.stack(chain)
        posed_joints = transforms[:, :3, 3].copy()
        jh = np.concatenate([J, np.zeros((24, 1), f32)], axis=1)[..., None]  # [24,4,1]
        tj = (transforms @ jh)[..., 0]  # [24,4]
        A = transforms.copy()
        A[:, :, 3] -= tj
        # lbs.py:236-246 skinning
        T = (self.lbs_weights @ A.reshape(24, 16)).reshape(-1, 4, 4).astype(f32)
        vh = np.concatenate([v_posed, np.ones((len(v_posed), 1), f32)], axis=1)
        verts = np.einsum("vij,vj->vi", T, vh)[:, :3].astype(f32)
        if transl is not None:  # body_models.py:353-360
            t = np.asarray(transl, f32).reshape(3)
            verts = verts + t
            posed_joints = posed_joints + t
            A = A.copy(); A[:, :3, 3] += t
            T = T.copy(); T[:, :3, 3] += t
        return {"vertices": verts, "A": A, "T": T, "joints": posed_joints, "v_shaped": v_shaped}
