"""Host-side mirror of instant_avatar/models/structures/body_model_param.py::SMPLParamEmbedding: per-frame SMPL
parameters held as learnable lookup tables (one row per training frame; `betas` has a single shared row), refined by
the pose-gradient path (`ia_pose_grad`) during training."""
from __future__ import annotations

import torch

POSE_KEYS = ("global_orient", "body_pose", "transl")


class SMPLParamEmbedding(torch.nn.Module):
    def __init__(self, **tables) -> None:
        super().__init__()
        for name, init in tables.items():
            setattr(self, name, torch.nn.Embedding(init.shape[0], init.shape[1], _weight=init.detach().clone().float()))
        self.keys = ["betas", "global_orient", "transl", "body_pose"]

    def forward(self, idx):
        row0 = torch.zeros_like(idx)  # shape parameters are shared by all frames
        out = {"betas": self.betas(row0)}
        for k in POSE_KEYS:
            out[k] = getattr(self, k)(idx)
        return out

    def tv_loss(self, idx):
        """temporal smoothness of the per-frame pose tables (squared first differences to both neighbours)"""
        last = self.global_orient.weight.shape[0] - 1
        before, after = (idx - 1).clamp(min=0), (idx + 1).clamp(max=last)
        total = 0
        for k in POSE_KEYS:
            tab = getattr(self, k)
            total = total + (tab(idx) - tab(before)).square().mean() + (tab(after) - tab(idx)).square().mean()
        return total
