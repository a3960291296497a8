"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the box, gloo in the CPU tests).

The reference has no distributed code.  The path shards by rays:
  * render: whole frames (or 8192-ray tiles dealt round-robin) per rank, no data-path collective;
  * training: each rank marches its shard of the step's rays, then ONE gradient collective: reduce-scatter (sum) of the
    flat fp32 gradient [encoder.params | color_net.params | pad] (13 036 208 elements, 52.1 MB), Adam on this rank's 1/G
    of the parameters (dividing by the world size), all-gather of the updated fp16 image (26 MB) -- optim.FusedAdam.
    The density-grid refresh is replicated bit-identically (same jitter on every rank).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

TILE = 8192  # BASELINE.json config 3: 8192 rays per batch


def init_from_env(backend: str | None = None, device: torch.device | None = None):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


_TILE_CACHE = {}


def shard_tiles_cached(n_rays: int, rank: int, world: int, tile: int, device, dtype=torch.int64) -> torch.Tensor:
    """device-resident copy of shard_tiles (built once: no per-frame host->device index upload)"""
    key = (n_rays, rank, world, tile, str(device), dtype)
    if key not in _TILE_CACHE:
        _TILE_CACHE[key] = shard_tiles(n_rays, rank, world, tile).to(device=device, dtype=dtype)
    return _TILE_CACHE[key]


def shard_tiles(n_rays: int, rank: int, world: int, tile: int = TILE) -> torch.Tensor:
    """indices of the rays this rank renders: tiles of `tile` consecutive rays dealt round-robin (the body sits in
    the image centre, so contiguous blocks would be badly imbalanced)"""
    n_tiles = (n_rays + tile - 1) // tile
    mine = torch.arange(rank, n_tiles, world)
    idx = (mine[:, None] * tile + torch.arange(tile)[None]).reshape(-1)
    return idx[idx < n_rays]


def shard_train_rays(n_rays: int, rank: int, world: int) -> slice:
    """contiguous equal shards of the step's rays (whole 32x32 patches while world <= 4)"""
    per = n_rays // world
    return slice(rank * per, (rank + 1) * per if rank < world - 1 else n_rays)


def allreduce_sum_(buffers, group=None):
    """the single collective of a training step: in-place sum of the flat gradient buffers"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for b in buffers:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    return buffers


def gather_image(local: torch.Tensor, idx: torch.Tensor, n_rays: int, group=None, tile: int = TILE) -> torch.Tensor | None:
    """assemble a ray-sharded render on rank 0: local [n_local, C] rows at global positions idx"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = local.new_zeros((n_rays, local.shape[1]))
        out[idx.to(local.device)] = local
        return out
    rank = dist.get_rank(group)
    if rank != 0:
        dist.gather(local.contiguous(), None, dst=0, group=group)
        return None
    idxs = [shard_tiles_cached(n_rays, r, world, tile, local.device) for r in range(world)]
    bufs = [local.new_empty((len(i), local.shape[1])) for i in idxs]
    dist.gather(local.contiguous(), bufs, dst=0, group=group)
    out = local.new_empty((n_rays, local.shape[1]))
    for i, b in zip(idxs, bufs):
        out[i] = b
    return out


def _has_native_scatter(group=None) -> bool:
    return dist.get_backend(group) == "nccl"


def reduce_scatter_sum(out: torch.Tensor, flat: torch.Tensor, group=None):
    """out = this rank's equal slice of sum_over_ranks(flat).  NCCL: one reduce-scatter over NVLink; gloo (CPU tests):
    all-reduce + slice."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert flat.numel() == out.numel() * world
    if _has_native_scatter(group):
        dist.reduce_scatter_tensor(out, flat, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out.copy_(flat[rank * out.numel():(rank + 1) * out.numel()])
    return out


def all_gather_inplace(flat: torch.Tensor, group=None):
    """every rank contributes its own equal slice of `flat` (already in place) and receives the others'"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = flat.numel() // world
    assert per * world == flat.numel()
    if _has_native_scatter(group):
        dist.all_gather_into_tensor(flat, flat[rank * per:(rank + 1) * per], group=group)
    else:
        parts = [torch.empty(per, dtype=flat.dtype) for _ in range(world)]
        dist.all_gather(parts, flat[rank * per:(rank + 1) * per].clone(), group=group)
        flat.copy_(torch.cat(parts))
    return flat


def all_gather_image(local: torch.Tensor, n_rays: int, tile: int, group=None) -> torch.Tensor | None:
    """Assemble a ray-sharded render on EVERY rank with one collective: ranks hold the rows of their round-robin tiles
    (shard_tiles) in tile order; all-gather -> [world, tiles_per_rank, tile, C] -> one permuting copy into image order.
    Needs n_rays % (tile * world) == 0 (else None: the caller falls back to gather_image)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or n_rays % (tile * world) != 0:
        return None
    per = n_rays // (tile * world)
    assert local.shape[0] == per * tile
    C = local.shape[1]
    buf = local.new_empty((world, per, tile, C))
    if _has_native_scatter(group):
        dist.all_gather_into_tensor(buf.view(-1), local.contiguous().view(-1), group=group)
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        buf.copy_(torch.stack(parts).view(world, per, tile, C))
    return buf.permute(1, 0, 2, 3).reshape(n_rays, C)  # tile (t, r) is global tile t * world + r


class PeerFrame:
    """Symmetric (peer-mapped) buffers for ONE cooperatively rendered frame (BASELINE.json config 3) on the GPUs of one
    NVLink box: every rank's 64^3 density grid and [n_pixels, 4] RGBA image are mapped into every other rank's address
    space (torch symmetric memory: cuMem handles exchanged through the process group's store).  The fused kernels then

      * max-reduce their shard of the occupancy queries into EVERY rank's density grid with NVLink atomics
        (ia_occupancy_query_peer: positive densities only, ~2 % of the cells) -- instead of a 1 MB max-all-reduce, and
      * store the RGBA of their rays straight into EVERY rank's image (ia_render_fwd_peer) -- instead of gather +
        scatter / all-gather + permute after the kernel,

    so the frame's only cross-GPU synchronisation is two signal-pad barriers.  Construction is collective; it raises if the
    platform cannot map peer memory (the caller falls back to the NCCL collectives)."""

    def __init__(self, n_pixels: int, device, group=None, grid: int = 64):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.density = symm.empty((grid, grid, grid), dtype=torch.float32, device=device)
        self.image = symm.empty((n_pixels, 4), dtype=torch.float32, device=device)
        self.h_density = symm.rendezvous(self.density, group)
        self.h_image = symm.rendezvous(self.image, group)
        self.density.zero_(); self.image.zero_()
        torch.cuda.synchronize(device)
        self.h_density.barrier(channel=0)   # every rank's buffers are zeroed before anyone writes into them
        torch.cuda.synchronize(device)

    @property
    def density_ptrs(self):
        return (self.h_density.buffer_ptrs_dev, self.world)

    def image_ptrs(self, pixel_index):
        return (pixel_index, self.h_image.buffer_ptrs_dev, self.world)

    def barrier_density(self):
        """all ranks' occupancy-query kernels (which write into every rank's grid) have finished"""
        self.h_density.barrier(channel=0)

    def barrier_image(self):
        """all ranks' render kernels have finished storing into every rank's image"""
        self.h_image.barrier(channel=0)
