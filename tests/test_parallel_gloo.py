"""CPU (gloo, world_size 2): the host-side multi-GPU logic -- ray sharding, the gradient all-reduce that is the one
collective of a training step, the replicated optimiser update, and image gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    from instantavatar_b200 import parallel
    r, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n_rays = 4096
    # --- train sharding: shards are disjoint and cover the step's rays ---
    sl = parallel.shard_train_rays(n_rays, rank, world)
    mine = torch.zeros(n_rays); mine[sl] = 1
    dist.all_reduce(mine)
    assert torch.all(mine == 1)
    # --- per-ray losses are means over the LOCAL rays; sum-all-reduce / world == the single-process gradient ---
    torch.manual_seed(0)
    target = torch.randn(n_rays)
    p = torch.zeros(8, requires_grad=True)
    feats = torch.randn(n_rays, 8)
    loss_local = ((feats[sl] @ p - target[sl]) ** 2).mean()
    g_local = torch.autograd.grad(loss_local, p)[0]
    g = g_local.clone()
    parallel.allreduce_sum_([g])
    g = g / world
    g_ref = torch.autograd.grad(((feats @ p - target) ** 2).mean(), p)[0]
    assert torch.allclose(g, g_ref, atol=1e-6)
    # --- replicated optimiser: identical parameters on every rank after the step ---
    q = torch.ones(8) - 1e-2 * g
    gathered = [torch.empty(8) for _ in range(world)]
    dist.all_gather(gathered, q)
    assert all(torch.equal(gathered[0], x) for x in gathered)
    # --- render sharding: round-robin 8192-ray tiles, gathered image equals the unsharded one ---
    n_img = 512 * 512
    idx = parallel.shard_tiles(n_img, rank, world)
    full = torch.arange(n_img, dtype=torch.float32)[:, None].repeat(1, 4)
    img = parallel.gather_image(full[idx], idx, n_img)
    # one-collective variant: every rank ends up with the image (all-gather + tile un-permutation)
    for tile in (parallel.TILE, 2048):
        idx_t = parallel.shard_tiles(n_img, rank, world, tile)
        img2 = parallel.all_gather_image(full[idx_t], n_img, tile)
        assert img2 is not None and torch.equal(img2, full)
    assert parallel.all_gather_image(full[parallel.shard_tiles(1000, rank, world, 300)], 1000, 300) is None  # ragged: caller falls back
    if rank == 0:
        assert torch.equal(img, full)
        ret["ok"] = True
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret.get("ok")


def test_shard_tiles_cover_and_balance():
    from instantavatar_b200 import parallel
    for world in (1, 2, 4, 8):
        seen = torch.zeros(512 * 512)
        counts = []
        for r in range(world):
            idx = parallel.shard_tiles(512 * 512, r, world)
            seen[idx] += 1
            counts.append(len(idx))
        assert torch.all(seen == 1)
        assert max(counts) - min(counts) <= parallel.TILE
