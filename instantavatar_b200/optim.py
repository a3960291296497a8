"""Fused dense Adam + device-side GradScaler for the two flat parameter tensors of NeRFNGPNet.

Mirrors DNeRFModel.configure_optimizers (models/DNeRF.py:32-59): one torch.optim.Adam with lr 1e-2, betas (0.9, 0.99),
eps 1e-15 (confs/SNARF_NGP.yaml:28-31), LambdaLR (1 - epoch/max_epochs)^1.5, and a manual GradScaler(init_scale=1024).
Dense semantics are kept (momentum of untouched hash-grid entries keeps decaying); the step, the unscale and the
inf/NaN skip run in one kernel per tensor with no host synchronisation.
"""
from __future__ import annotations

import torch

from . import ops


class GradScaler:
    """torch.cuda.amp.GradScaler semantics (growth 2x every 2000 clean steps, backoff 0.5) kept on the device"""

    def __init__(self, device, init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.scale_t = torch.full((1,), init_scale, device=device, dtype=torch.float32)
        self.growth_tracker = torch.zeros(1, device=device, dtype=torch.int32)
        self.found_inf = torch.zeros(1, device=device, dtype=torch.float32)
        self.gf, self.bf, self.gi = growth_factor, backoff_factor, growth_interval

    def scale(self, loss):
        return loss * self.scale_t

    def update(self):
        torch._amp_update_scale_(self.scale_t, self.growth_tracker, self.found_inf, self.gf, self.bf, self.gi)
        self.found_inf.zero_()


class FusedAdam:
    """All per-step scalars (step count, bias corrections, lr, 1/scale) live in a device tensor, so a training step is
    a fixed launch sequence that can be captured in a CUDA graph and replayed."""

    def __init__(self, net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, max_epochs=30):
        self.net = net
        self.base_lr, self.betas, self.eps = lr, betas, eps
        self.max_epochs, self.epoch = max_epochs, 0
        self.params = [net.encoder.params, net.color_net.params]
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.params]
        dev = net.encoder.params.device
        # {lr, beta1, beta2, eps, step, bc1, bc2_sqrt, inv_scale}
        self.state_t = torch.tensor([lr, betas[0], betas[1], eps, 0.0, 1.0, 1.0, 1.0], dtype=torch.float32).to(dev)

    @property
    def lr_factor(self):  # LambdaLR of DNeRF.py:52-55, stepped in on_validation_epoch_end only
        return (1 - self.epoch / self.max_epochs) ** 1.5

    @property
    def lr(self):
        return self.base_lr * self.lr_factor

    @property
    def step_count(self):
        return int(self.state_t[4].item())

    def scheduler_step(self):
        self.epoch += 1
        self.state_t[0:1].fill_(self.lr)

    def zero_grad(self):
        """explicit zeroing (the fused step already leaves the gradient buffers zeroed)"""
        for g in self.net.grad_buffers():
            g.zero_()

    def step(self, scaler: GradScaler | None = None, world_size: int = 1):
        g_enc, g_col = self.net.grad_buffers()
        found = scaler.found_inf if scaler is not None else None
        if scaler is not None:
            ops.grad_check_finite(g_enc, found)
            ops.grad_check_finite(g_col, found)
        ops.adam_prepare(self.state_t, 1.0 / world_size, scaler.scale_t if scaler is not None else None, found)
        table_h, mlp_h = self.net.half_buffers()
        (m0, v0), (m1, v1) = self.state
        ops.adam_step_dev(self.params[0].data, g_enc, m0, v0, self.state_t, found, table_h, 3072)
        ops.adam_step_dev(self.params[1].data, g_col, m1, v1, self.state_t, found, None, 0)
        ops.mlp_to_half(self.params[0].data, self.params[1].data, mlp_h)
        self.net.mark_clean()


class DeviceAdam:
    """torch.optim.Adam semantics for a short list of small dense tensors (the SMPL pose embeddings, DNeRF.py:40-51)
    on the same device-state kernels as FusedAdam: step count, bias corrections and the GradScaler's 1/scale live in an
    8-float device tensor and an overflow skips the step on the device -- no host read-back, capturable in a CUDA graph."""

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.99), eps=1e-15):
        self.params = [p for p in params]
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.params]
        dev = self.params[0].device
        self.state_t = torch.tensor([lr, betas[0], betas[1], eps, 0.0, 1.0, 1.0, 1.0], dtype=torch.float32).to(dev)
        self.base_lr = lr

    @property
    def lr(self):
        """current learning rate (host copy of the device-resident value the kernels read)"""
        return self._lr if hasattr(self, "_lr") else self.base_lr

    def set_lr(self, lr: float):
        """the learning rate lives in the device state the kernels read (state_t[0]); written without a host sync"""
        self._lr = float(lr)
        self.state_t[0:1].fill_(self._lr)

    def set_lr_factor(self, factor: float):
        """LambdaLR semantics: lr = base_lr * factor (DNeRF.py:52-55 applies one lambda to all groups)"""
        self.set_lr(self.base_lr * float(factor))

    def zero_grad(self, set_to_none=False):
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()

    def grads(self):
        return [p.grad for p in self.params if p.grad is not None]

    def check_finite(self, scaler: GradScaler):
        for g in self.grads():
            ops.grad_check_finite(g, scaler.found_inf)

    def step(self, scaler: GradScaler | None = None, world_size: int = 1):
        found = scaler.found_inf if scaler is not None else None
        ops.adam_prepare(self.state_t, 1.0 / world_size, scaler.scale_t if scaler is not None else None, found)
        for p, (m, v) in zip(self.params, self.state):
            if p.grad is None:
                continue
            ops.adam_step_dev(p.data, p.grad, m, v, self.state_t, found, None, 0)
