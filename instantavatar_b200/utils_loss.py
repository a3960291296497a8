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


class NGPLoss(nn.Module):
    """Mirror of instant_avatar/utils/loss.py::NGPLoss (:8-51): NeRFLoss's terms plus, on patch-shaped predictions
    ([B, P, h, w, 3], PatchSampler), an optional LPIPS term and the depth-variance regulariser.  The LPIPS-VGG weights
    cannot be downloaded here: pass a callable `lpips(pred_nchw, target_nchw)` to use the term; `w_lpips > 0` without
    one raises.  SNARF_NGP_refine.yaml uses w_rgb / w_alpha / w_reg only."""

    def __init__(self, opt=None, w_rgb=1.0, w_alpha=0.1, w_reg=0.1, w_lpips=0.0, w_depth_reg=0.0, lpips=None) -> None:
        super().__init__()
        g = lambda k, d: (opt.get(k, d) if hasattr(opt, "get") else getattr(opt, k, d)) if opt is not None else d
        self.w_rgb, self.w_alpha, self.w_reg = g("w_rgb", w_rgb), g("w_alpha", w_alpha), g("w_reg", w_reg)
        self.w_lpips, self.w_depth_reg = g("w_lpips", w_lpips), g("w_depth_reg", w_depth_reg)
        self.lpips = lpips
        if self.w_lpips > 0 and lpips is None:
            raise RuntimeError("NGPLoss: w_lpips > 0 needs an LPIPS callable (the VGG weights are not available offline)")

    def forward(self, predicts, targets):
        losses = {}
        rgb, alpha = predicts["rgb_coarse"], predicts["alpha_coarse"]
        losses["mse_loss"] = F.mse_loss(rgb, targets["rgb"], reduction="mean")
        losses["loss_alpha_coarse"] = F.mse_loss(alpha, targets["alpha"])
        loss = self.w_rgb * losses["mse_loss"] + self.w_alpha * losses["loss_alpha_coarse"]
        patches = rgb.dim() == 5
        if self.w_lpips > 0 and patches:
            nchw = lambda t: t[..., [2, 1, 0]].flatten(0, 1).permute(0, 3, 1, 2)
            losses["loss_lpips"] = self.lpips(nchw(rgb).clip(max=1), nchw(targets["rgb"])).sum()
            loss = loss + self.w_lpips * losses["loss_lpips"]
        if self.w_depth_reg > 0 and patches:
            depth = predicts["depth_coarse"]
            mean_depth = (depth * alpha).sum(dim=(-1, -2)) / (alpha.sum(dim=(-1, -2)) + 1e-3)
            losses["loss_depth_reg"] = (alpha * (depth - mean_depth[..., None, None]).abs()).mean()
            loss = loss + self.w_depth_reg * losses["loss_depth_reg"]
        OFFSET = 0.313262
        reg = lambda x: (-torch.log(torch.exp(-x) + torch.exp(x - 1))).mean() + OFFSET
        losses["reg_alpha"], losses["reg_density"] = reg(alpha), reg(predicts["weight_coarse"])
        losses["loss"] = loss + self.w_reg * losses["reg_alpha"] + self.w_reg * losses["reg_density"]
        return losses
