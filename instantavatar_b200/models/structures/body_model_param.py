"""Learnable per-frame SMPL parameters (the capability of instant_avatar/models/structures/body_model_param.py::
SMPLParamEmbedding, `optimize_SMPL.enable`): one lookup table per parameter, one row per training frame, a single shared
row for the shape.  Attribute names match the reference so that `SMPL_param.<name>.weight` state-dict keys carry over.
The pose tables are refined through `ia_pose_grad` -> `ia_smpl_tfs_backward` during training."""
from __future__ import annotations

import torch

POSE_KEYS = ("global_orient", "body_pose", "transl")


class SMPLParamEmbedding(torch.nn.Module):
    def __init__(self, **tables) -> None:
        super().__init__()
        self.keys = ["betas", "global_orient", "transl", "body_pose"]
        for name, init in tables.items():
            table = torch.nn.Embedding(*init.shape)
            table.weight.data.copy_(init.detach().float())
            self.add_module(name, table)

    def rows(self, name, idx):
        return getattr(self, name)(idx)

    def forward(self, idx):
        params = {name: self.rows(name, idx) for name in POSE_KEYS}
        params["betas"] = self.rows("betas", idx * 0)  # the shape is shared by all frames
        return params

    def tv_loss(self, idx):
        """temporal smoothness: mean squared difference of every pose table row to its two neighbouring frames
        (the first / last frame count themselves as the missing neighbour)"""
        n_frames = self.global_orient.num_embeddings
        prev_idx, next_idx = (idx - 1).clamp_min(0), (idx + 1).clamp_max(n_frames - 1)
        penalty = 0
        for name in POSE_KEYS:
            here, before, after = self.rows(name, idx), self.rows(name, prev_idx), self.rows(name, next_idx)
            penalty = penalty + (here - before).pow(2).mean() + (after - here).pow(2).mean()
        return penalty
