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
    parameters (fp32 master, m, v touched only there: 443 MB / G of HBM traffic instead of 443 MB replicated) ->
    all-gather of the flat fp16 image (26 MB) that the forward kernels read.  On an NVLink box both exchanges run INSIDE two
    kernels over peer memory (`_step_peer`: every rank's gradient vector and fp16 image are symmetric-memory mappings; the
    shard sum reads all peers' gradients, the Adam pass stores its fp16 result into all peers' images); otherwise NCCL.  The fp32 masters of the other shards go
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
        self.use_peer = True   # sharded step: exchange over NVLink peer memory when the platform can map it (else NCCL)
        self._peer = None      # None: not tried yet; False: unavailable; dict: symmetric-memory handles

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

    def _try_enable_peer(self, world_size: int, group=None):
        """Move the gradient vector and the fp16 image into symmetric (peer-mapped) memory; collective.  All ranks agree."""
        import os
        import torch.distributed as dist
        ok, handles = False, None
        dev = self.flat_g.device
        if self.use_peer and dev.type == "cuda" and dist.get_backend(group) == "nccl" and not os.environ.get("IA_B200_NO_PEER"):
            try:
                import torch.distributed._symmetric_memory as symm
                grp = group if group is not None else dist.group.WORLD
                g = symm.empty(self.flat_g.numel(), dtype=torch.float32, device=dev)
                h = symm.empty(self.flat_h.numel(), dtype=torch.float16, device=dev)
                fl = symm.empty(64, dtype=torch.float32, device=dev)
                handles = {"g": symm.rendezvous(g, grp), "h": symm.rendezvous(h, grp), "f": symm.rendezvous(fl, grp), "flags": fl,
                           "bufs": (g, h)}
                ok = True
            except Exception as exc:  # no peer mapping on this platform
                self._peer_error = f"{type(exc).__name__}: {exc}"[:200]
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() != 1.0:
            self._peer = False
            return
        g, h = handles["bufs"]
        g.copy_(self.flat_g); h.copy_(self.flat_h); handles["flags"].zero_()
        enc, col = self.params
        self.flat_g, self.flat_h = g, h
        enc.grad, col.grad = g[:self.n_enc], g[self.n_enc:self.n]
        if hasattr(self.net, "adopt_half_table"):
            dirty = getattr(self.net, "_dirty", True)
            self.net.adopt_half_table(h[self.n_mlp:self.n_enc].view((self.n_enc - self.n_mlp) // 2, 2))
            self.net._dirty = dirty  # the contents were copied: no refresh from the masters needed
        torch.cuda.synchronize(dev)
        handles["g"].barrier(channel=0)
        torch.cuda.synchronize(dev)
        self._peer = handles

    def prepare(self, world_size: int, group=None):
        """one-time collective set-up of the sharded step (shard buffer, peer mappings); call it before capturing a CUDA
        graph or snapshotting optimiser state -- `step` does it lazily otherwise"""
        if world_size <= 1:
            return
        S, _ = shard_layout(self.n, world_size)
        if self._shard_g is None or self._shard_g.numel() != S:
            self._shard_g = torch.zeros(S, device=self.flat_g.device, dtype=torch.float32)
        if self._peer is None:
            self._try_enable_peer(world_size, group)

    def _step_peer(self, found, scale_t, world_size, rank, S):
        """reduce-scatter, overflow agreement and all-gather INSIDE two kernels over NVLink peer memory (3 barriers)"""
        pr = self._peer
        lo, hi = rank * S, (rank + 1) * S
        pr["g"].barrier(channel=0)                       # every rank's backward has finished writing its gradient
        ops.peer_reduce_check(pr["g"].buffer_ptrs_dev, world_size, lo, self._shard_g, pr["f"].buffer_ptrs_dev, rank, found)
        pr["f"].barrier(channel=0)                       # every rank has read all gradients and raised its flag
        ops.peer_flags_to_found(pr["flags"], world_size, found)
        self.flat_g.zero_()
        ops.adam_prepare(self.state_t, 1.0 / world_size, scale_t, found)
        ops.adam_step_dev_peer(self.flat_p[lo:hi], self._shard_g, self.flat_m[lo:hi], self.flat_v[lo:hi], self.state_t, found,
                               pr["h"].buffer_ptrs_dev, world_size, lo)
        pr["h"].barrier(channel=0)                       # every rank's shard of the fp16 image has landed everywhere
        self._refresh_mlp()
        self.masters_stale = True

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
        self.prepare(world_size, group)
        if self._peer:
            if found is None:
                found = self._own_found = getattr(self, "_own_found", None) or torch.zeros(1, device=self.flat_g.device)
            return self._step_peer(found, scale_t, world_size, rank, S)
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
