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


def shard_layout(n_total: int, world: int) -> tuple[int, int]:
    """(shard_elems, padded_total): equal shards of the flat parameter vector, each a multiple of 4 floats (the Adam kernel
    works on float4 / writes half2 pairs), covering n_total"""
    per = -(-n_total // world)
    per = -(-per // 4) * 4
    return per, per * world


class FusedAdam:
    """Dense Adam over ONE flat fp32 vector [encoder.params | color_net.params | pad] (DNeRF.py:32-59).

    `net.encoder.params` / `net.color_net.params` (and their `.grad`) become views into flat buffers, so that a training
    step needs ONE gradient collective.  All per-step scalars (step count, bias corrections, lr, 1/scale) live in a device
    tensor: the step is a fixed launch sequence that can be captured in a CUDA graph.

    world_size > 1 (sharded): reduce-scatter (sum) of the flat gradient -> Adam on this rank's 1/G of the
    parameters (fp32 master, m, v touched only there: 365 MB / G of HBM traffic instead of 365 MB replicated) ->
    all-gather of the flat fp16 image (26 MB) that the forward kernels read.  The fp32 masters of the other shards go
    stale on this rank; `gather_master_params()` all-gathers them (checkpointing / state_dict).  A local overflow is
    broadcast inside the reduce-scatter (ia_grad_poison_shards), so every rank skips the same steps."""

    PAD = 64  # room for shard rounding up to 16 ranks

    def __init__(self, net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, max_epochs=30):
        self.net = net
        self.base_lr, self.betas, self.eps = lr, betas, eps
        self.max_epochs, self.epoch = max_epochs, 0
        enc, col = net.encoder.params, net.color_net.params
        dev = enc.device
        self.n_enc, self.n_col = enc.numel(), col.numel()
        self.n = self.n_enc + self.n_col
        z = lambda dt=torch.float32: torch.zeros(self.n + self.PAD, device=dev, dtype=dt)
        self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_h = z(), z(), z(), z(), z(torch.float16)
        with torch.no_grad():
            self.flat_p[:self.n_enc].copy_(enc.detach()); self.flat_p[self.n_enc:self.n].copy_(col.detach())
        enc.data, col.data = self.flat_p[:self.n_enc], self.flat_p[self.n_enc:self.n]
        enc.grad, col.grad = self.flat_g[:self.n_enc], self.flat_g[self.n_enc:self.n]
        self.params = [enc, col]
        # the fp16 hash table the kernels read is a view into the flat fp16 image (refreshed by the Adam kernel)
        from . import _lib
        self.n_mlp = _lib.IA_ENC_MLP_PARAMS
        if hasattr(net, "adopt_half_table"):
            net.adopt_half_table(self.flat_h[self.n_mlp:self.n_enc].view((self.n_enc - self.n_mlp) // 2, 2))
        # {lr, beta1, beta2, eps, step, bc1, bc2_sqrt, inv_scale}
        self.state_t = torch.tensor([lr, betas[0], betas[1], eps, 0.0, 1.0, 1.0, 1.0], dtype=torch.float32).to(dev)
        self._shard_g = None
        self.masters_stale = False

    @property
    def lr_factor(self):  # LambdaLR of DNeRF.py:52-55, stepped in on_validation_epoch_end only
        return (1 - self.epoch / self.max_epochs) ** 1.5

    @property
    def lr(self):
        return self.base_lr * self.lr_factor

    @property
    def step_count(self):
        return int(self.state_t[4].item())

    @property
    def state(self):
        """[(exp_avg, exp_avg_sq)] per parameter tensor (views into the flat moment buffers)"""
        e, n = self.n_enc, self.n
        return [(self.flat_m[:e], self.flat_v[:e]), (self.flat_m[e:n], self.flat_v[e:n])]

    def scheduler_step(self):
        self.epoch += 1
        self.state_t[0:1].fill_(self.lr)

    def zero_grad(self):
        """explicit zeroing (the fused step already leaves the gradient buffer zeroed)"""
        self.flat_g.zero_()

    def _refresh_mlp(self):
        _, mlp_h = self.net.half_buffers()
        ops.mlp_to_half_from_half(self.flat_h[:self.n_mlp], self.flat_h[self.n_enc:self.n], mlp_h)
        self.net.mark_clean()

    def step(self, scaler: GradScaler | None = None, world_size: int = 1, group=None):
        found = scaler.found_inf if scaler is not None else None
        scale_t = scaler.scale_t if scaler is not None else None
        self.net.half_buffers()  # make sure the fp16 working copies exist (first step)
        if world_size == 1:
            n4 = -(-self.n // 4) * 4
            if found is not None:
                ops.grad_check_finite(self.flat_g[:self.n], found)
            ops.adam_prepare(self.state_t, 1.0, scale_t, found)
            ops.adam_step_dev(self.flat_p[:n4], self.flat_g[:n4], self.flat_m[:n4], self.flat_v[:n4], self.state_t, found, self.flat_h[:n4], 0)
            self._refresh_mlp()
            return
        import torch.distributed as dist
        from . import parallel
        rank = dist.get_rank(group)
        S, L = shard_layout(self.n, world_size)
        if self._shard_g is None or self._shard_g.numel() != S:
            self._shard_g = torch.zeros(S, device=self.flat_g.device, dtype=torch.float32)
        if found is not None:  # a local overflow must skip the step on EVERY rank: it rides inside the reduce-scatter
            ops.grad_check_finite(self.flat_g[:self.n], found)
            ops.grad_poison_shards(self.flat_g[:L], S, world_size, found)  # `found` may already carry the pose group's flag
        parallel.reduce_scatter_sum(self._shard_g, self.flat_g[:L], group)
        self.flat_g.zero_()
        if found is not None:
            ops.grad_check_finite(self._shard_g, found)
        ops.adam_prepare(self.state_t, 1.0 / world_size, scale_t, found)
        lo, hi = rank * S, (rank + 1) * S
        ops.adam_step_dev(self.flat_p[lo:hi], self._shard_g, self.flat_m[lo:hi], self.flat_v[lo:hi], self.state_t, found, self.flat_h[lo:hi], 0)
        parallel.all_gather_inplace(self.flat_h[:L], group)
        self._refresh_mlp()
        self.masters_stale = True

    def gather_master_params(self, world_size: int, group=None):
        """all-gather the fp32 master shards (each rank only keeps its own shard current while training sharded)"""
        if world_size == 1 or not self.masters_stale:
            return
        from . import parallel
        S, L = shard_layout(self.n, world_size)
        parallel.all_gather_inplace(self.flat_p[:L], group)
        self.masters_stale = False


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
