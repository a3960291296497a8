"""CPU (gloo, world_size 2): the sharded optimiser step of optim.FusedAdam -- reduce-scatter of the flat gradient, Adam on
this rank's shard, all-gather of the fp16 image, overflow flag carried inside the gradient collective -- against one
replicated torch.optim.Adam on the summed gradient (DNeRF.py:46-59,152-159).  The CUDA operators are replaced by torch
stand-ins that restate the kernels (csrc/ia_train.cu: adam_prepare_kernel, adam_dev_kernel, grad_finite_kernel,
grad_poison_kernel); the partition arithmetic, the collectives and the skip protocol are the product's."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _P(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.params = torch.nn.Parameter(torch.zeros(n))


class FakeNet(torch.nn.Module):
    """two flat fp32 tensors like NeRFNGPNet (3072 MLP weights + a tiny table, 6144 colour weights)"""

    def __init__(self, n_table=2 * 37):
        super().__init__()
        self.encoder, self.color_net = _P(3072 + n_table), _P(6144)
        self.mlp_refreshes = 0
        self._mlp_h = torch.zeros(8, dtype=torch.float16)

    def adopt_half_table(self, t):
        self.table_h = t

    def half_buffers(self):
        return self.table_h, self._mlp_h

    def mark_clean(self):
        pass


def _install_standins(ops, net):
    def grad_check_finite(g, found):
        if not torch.isfinite(g).all():
            found.fill_(1.0)

    def grad_poison_shards(g, S, n, found):
        if found.item() != 0:
            for k in range(n):
                g[k * S] = float("nan")

    def adam_prepare(state, inv_world=1.0, scale=None, found=None):
        if not (found is not None and found.item() != 0):
            state[4] += 1
        t = max(state[4].item(), 1.0)
        state[5] = 1.0 - state[1].item() ** t
        state[6] = math.sqrt(1.0 - state[2].item() ** t)
        state[7] = inv_world / scale.item() if scale is not None else inv_world

    def adam_step_dev(p, g, m, v, state, found=None, half_out=None, half_skip=0):
        gi = g.clone() * state[7]
        g.zero_()
        if not (found is not None and found.item() != 0):
            lr, b1, b2, eps, bc1, bc2s = (state[i].item() for i in (0, 1, 2, 3, 5, 6))
            m.mul_(b1).add_(gi, alpha=1 - b1)
            v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
            p.sub_((lr / bc1) * m / (v.sqrt() / bc2s + eps))
        if half_out is not None:
            half_out[: p.numel() - half_skip].copy_(p[half_skip:].half())

    def mlp_to_half_from_half(a, b, out):
        net.mlp_refreshes += 1

    for f in (grad_check_finite, grad_poison_shards, adam_prepare, adam_step_dev, mlp_to_half_from_half):
        setattr(ops, f.__name__, f)


def _worker(rank, world, port, ret):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo")
    from instantavatar_b200 import ops, optim
    torch.manual_seed(0)
    net = FakeNet()
    with torch.no_grad():
        net.encoder.params.copy_(torch.randn_like(net.encoder.params) * 0.1)
        net.color_net.params.copy_(torch.randn_like(net.color_net.params) * 0.1)
    _install_standins(ops, net)
    # reference: ONE replicated torch Adam on the world-averaged gradient
    ref_p = torch.cat([net.encoder.params.detach().clone(), net.color_net.params.detach().clone()]).requires_grad_(True)
    ref = torch.optim.Adam([ref_p], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    opt = optim.FusedAdam(net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = optim.GradScaler("cpu", init_scale=1024.0)
    scaler.update = lambda: scaler.found_inf.zero_()  # torch._amp_update_scale_ is CUDA-only; growth is not under test
    n = opt.n
    S, L = optim.shard_layout(n, world)
    assert S % 4 == 0 and L >= n and L - n < 4 * world + 4
    g_enc, g_col = net.encoder.params.grad, net.color_net.params.grad
    assert g_enc.data_ptr() == opt.flat_g.data_ptr() and net.encoder.params.data_ptr() == opt.flat_p.data_ptr()
    for step in range(3):
        gens = [torch.Generator().manual_seed(100 * step + r) for r in range(world)]
        per_rank = [torch.randn(n, generator=g) for g in gens]          # every rank can form the reference sum
        mine = per_rank[rank] * 1024.0                                   # scaled by the GradScaler, as the backward leaves it
        g_enc.copy_(mine[: opt.n_enc]); g_col.copy_(mine[opt.n_enc:])
        opt.step(scaler, world)
        scaler.update()
        ref_p.grad = sum(per_rank) / world
        ref.step()
    lo, hi = rank * S, min((rank + 1) * S, n)
    assert torch.allclose(opt.flat_p[lo:hi], ref_p.detach()[lo:hi], rtol=1e-5, atol=1e-7), "own shard of the fp32 masters"
    assert torch.equal(opt.flat_h[:n], opt.flat_h[:n]) and torch.allclose(opt.flat_h[:n].float(), ref_p.detach(), atol=2e-3), "fp16 image"
    other = (rank + 1) % world
    olo, ohi = other * S, min((other + 1) * S, n)
    assert not torch.allclose(opt.flat_p[olo:ohi], ref_p.detach()[olo:ohi], atol=1e-6), "foreign shard masters are stale by design"
    assert torch.count_nonzero(opt.flat_g) == 0 and net.mlp_refreshes == 3 and opt.step_count == 3
    opt.gather_master_params(world)
    assert torch.allclose(opt.flat_p[:n], ref_p.detach(), rtol=1e-5, atol=1e-7), "masters after the fp32 all-gather"
    # the fp16 image is identical on every rank
    img = [torch.empty(n, dtype=torch.float16) for _ in range(world)]
    dist.all_gather(img, opt.flat_h[:n].clone())
    assert all(torch.equal(img[0], x) for x in img)
    # ---- overflow on ONE rank, in an element owned by the OTHER rank's shard: everybody skips ----
    before = opt.flat_p[:n].clone()
    g_enc.zero_(); g_col.zero_()
    g_enc[5] = 1.0
    if rank == 1:
        g_enc[7] = float("inf")  # element 7 belongs to rank 0's shard
    opt.step(scaler, world)
    assert scaler.found_inf.item() == 1.0, "every rank sees the overflow"
    assert torch.equal(opt.flat_p[:n], before) and opt.step_count == 3
    scaler.update()
    assert torch.count_nonzero(torch.nan_to_num(opt.flat_g, nan=1.0)) == 0
    ret[rank] = True
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_adam_world_size_2_gloo():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret.get(0) and ret.get(1)


def test_shard_layout():
    from instantavatar_b200.optim import shard_layout
    n = 13036208
    for w in (1, 2, 3, 4, 8, 16):
        S, L = shard_layout(n, w)
        assert S % 4 == 0 and S * w == L and n <= L <= n + 64
