"""Mirror of instant_avatar/utils/loss.py::NeRFLoss (PyTorch ops: it only touches the per-ray outputs)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class NeRFLoss(nn.Module):
    def __init__(self, opt=None, w_rgb=1.0, w_alpha=0.1, w_reg=0.1) -> None:
        super().__init__()
        g = lambda k, d: (opt.get(k, d) if hasattr(opt, "get") else getattr(opt, k, d)) if opt is not None else d
        self.w_rgb, self.w_alpha, self.w_reg = g("w_rgb", w_rgb), g("w_alpha", w_alpha), g("w_reg", w_reg)

    def forward(self, predicts, targets):
        losses = {}
        loss_rgb = F.mse_loss(predicts["rgb_coarse"], targets["rgb"], reduction="mean")
        loss_alpha = F.mse_loss(predicts["alpha_coarse"], targets["alpha"])
        OFFSET = 0.313262
        reg = lambda x: (-torch.log(torch.exp(-x) + torch.exp(x - 1))).mean() + OFFSET
        reg_alpha, reg_density = reg(predicts["alpha_coarse"]), reg(predicts["weight_coarse"])
        losses["mse_loss"], losses["loss_alpha_coarse"] = loss_rgb, loss_alpha
        losses["reg_alpha"], losses["reg_density"] = reg_alpha, reg_density
        losses["loss"] = self.w_rgb * loss_rgb + self.w_alpha * loss_alpha + self.w_reg * reg_alpha + self.w_reg * reg_density
        return losses
