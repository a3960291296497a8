"""GPU x2 (NCCL): one frame rendered cooperatively by two ranks equals the single-GPU frame bit for bit, and the
ray-sharded training gradient (one all-reduce) equals the single-GPU gradient.  Skipped on single-GPU boxes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from instantavatar_b200 import parallel, synthetic
    from instantavatar_b200.models.dnerf import DNeRFModel
    H = W = 256
    model = DNeRFModel(smpl_data=synthetic.smpl_dict_cached(0), device=dev).eval()
    pose = synthetic.load_pose(0)
    o, d = synthetic.demo_camera_rays(512, 512)
    idx = (np.arange(0, 512, 2)[:, None] * 512 + np.arange(0, 512, 2)[None]).ravel()
    batch = {"rays_o": torch.from_numpy(o[idx][None]).to(dev), "rays_d": torch.from_numpy(d[idx][None]).to(dev),
             "near": torch.zeros((1, len(idx)), device=dev), "far": torch.ones((1, len(idx)), device=dev)}
    batch.update({k: torch.from_numpy(v).to(dev) for k, v in pose.items()})
    model.deformer.prepare_deformer(batch)
    model.net_coarse.initialize(model.deformer.bbox)
    bbox = model.deformer.bbox.cpu().numpy().astype(np.float64)
    enc, col = synthetic.analytic_avatar_params(model.deformer.joints_cano[0].cpu().numpy(), (bbox[0] + bbox[1]) / 2, bbox[1] - bbox[0])
    model.net_coarse.load_flat_params(torch.from_numpy(enc).to(dev), torch.from_numpy(col).to(dev))
    torch.manual_seed(7)
    jit = torch.rand((5, 64, 64, 64, 3), device=dev)
    # ---- cooperative frame vs single-GPU frame ----
    img = model.render_image_sharded(dict(batch), (H, W), rank, world, jit, tile=1024)
    rgb, depth, alpha, counter = model.render_image_fast(dict(batch), (H, W), jit)
    full = torch.cat([rgb.reshape(-1, 3), alpha.reshape(-1, 1)], dim=1)
    if rank == 0:
        ret["frame_equal"] = bool(torch.equal(img, full))
        ret["hit"] = int((alpha > 0.5).sum().item())
    # ---- the same frame with both exchanges inside the kernels over NVLink peer memory (two frames: a second pose checks
    #      the buffer-reuse protocol), every rank must end up with the single-GPU image ----
    try:
        pf = parallel.PeerFrame(H * W, dev)
        peer_ok = torch.ones(1, device=dev)
    except Exception as exc:
        pf, peer_ok = None, torch.zeros(1, device=dev)
        if rank == 0:
            ret["peer_error"] = f"{type(exc).__name__}: {exc}"[:300]
    dist.all_reduce(peer_ok, op=dist.ReduceOp.MIN)
    if peer_ok.item() == 1:
        eq = torch.ones(1, device=dev)
        pose2 = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.load_pose(57).items()}
        for b_ in (batch, {**batch, **pose2}, batch):
            ref_rgb, _, ref_a, _ = model.render_image_fast(dict(b_), (H, W), jit)
            ref = torch.cat([ref_rgb.reshape(-1, 3), ref_a.reshape(-1, 1)], dim=1)
            got = model.render_image_sharded(dict(b_), (H, W), rank, world, jit, tile=1024, peer=pf)
            torch.cuda.synchronize()
            if not torch.equal(got, ref):
                eq.zero_()
            dist.barrier()  # `got` is the symmetric image buffer: everyone has compared before the next frame overwrites it
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret["peer_frames_equal"] = bool(eq.item() == 1)
    elif rank == 0:
        ret["peer_frames_equal"] = None
    # ---- sharded training gradient vs single-GPU gradient ----
    n = 1024
    pick = torch.arange(100 * W + 96, 100 * W + 96 + n, device=dev)
    tb = dict(batch)
    for k in ("rays_o", "rays_d", "near", "far"):
        tb[k] = batch[k][:, pick].contiguous()
    torch.manual_seed(11)
    tb["rgb"] = torch.rand((1, n, 3), device=dev); tb["alpha"] = torch.ones((1, n), device=dev); tb["bg_color"] = torch.rand((1, n, 3), device=dev)
    jitter = torch.rand((n, 256), device=dev); noise = torch.zeros((n, 256), device=dev)
    model.train(); model.global_step = 2001   # no grid refresh on this step

    def grads(b, jit_, noi_, world_size):
        from instantavatar_b200 import ops
        from instantavatar_b200.autograd import GRAD_SCALE
        from instantavatar_b200.models.dnerf import Rays
        g_enc, g_col = model.net_coarse.grad_buffers(); g_enc.zero_(); g_col.zero_()
        rays = Rays(o=b["rays_o"], d=b["rays_d"], near=b["near"], far=b["far"])
        model.deformer.transform_rays_w2s(rays)
        grid = model.renderer.density_grid_train
        if grid.aabb is None:
            grid.aabb = model.renderer.aabb
        grid.set_field(torch.ones((64, 64, 64), dtype=torch.bool, device=dev))
        scene = model.deformer.scene(model.net_coarse, grid.occupancy_bits(), grid.aabb6())
        o_, d_ = rays.o.reshape(-1, 3).contiguous(), rays.d.reshape(-1, 3).contiguous()
        ne, fa = rays.near.reshape(-1).contiguous(), rays.far.reshape(-1).contiguous()
        bg = b["bg_color"].reshape(-1, 3).contiguous()
        out, saved = ops.train_fwd(scene, o_, d_, ne, fa, bg, jit_, noi_)
        losses, g_rgb, g_alpha, g_w = ops.nerf_loss(out, b["rgb"], b["alpha"])
        l = ops.composite_bwd(ne, fa, bg, noi_, saved, g_rgb, None, g_alpha, g_w)
        ops.ngp_backward(scene, l[0], l[1], l[2], l[3], g_enc, g_col, GRAD_SCALE)
        ge, gc = g_enc.clone(), g_col.clone()
        g_enc.zero_(); g_col.zero_()
        if world_size > 1:
            parallel.allreduce_sum_([ge, gc])
        return ge / world_size, gc / world_size

    full_enc, full_col = grads(tb, jitter, noise, 1)
    sl = parallel.shard_train_rays(n, rank, world)
    sb = dict(tb)
    for k in ("rays_o", "rays_d", "near", "far", "rgb", "alpha", "bg_color"):
        sb[k] = tb[k][:, sl].contiguous()
    sh_enc, sh_col = grads(sb, jitter[sl].contiguous(), noise[sl].contiguous(), world)
    if rank == 0:
        rel = lambda a, b: float((a - b).norm() / b.norm())
        ret["rel_enc"] = rel(sh_enc, full_enc); ret["rel_col"] = rel(sh_col, full_col)
        ret["gnorm"] = float(full_enc.norm())
    # ---- sharded optimiser: 3 training steps on ray shards (reduce-scatter, Adam on 1/G, all-gather of the fp16 image)
    #      against 3 single-process steps on the full batch from the same initial state ----
    def fresh_model():
        m = DNeRFModel(smpl_data=synthetic.smpl_dict_cached(0), device=dev)
        m.deformer.prepare_deformer(batch)
        m.net_coarse.initialize(m.deformer.bbox)
        m.net_coarse.load_flat_params(torch.from_numpy(enc).to(dev), torch.from_numpy(col).to(dev))
        m.global_step = 2001
        m.renderer.density_grid_train.set_field(torch.ones((64, 64, 64), dtype=torch.bool, device=dev))  # march everything
        return m

    def steps(m, b, jit_, noi_, k=3):
        for _ in range(k):
            m.training_step(dict(b), jitter=jit_, noise_tensor=noi_)
        return m

    single = steps(fresh_model(), tb, jitter, noise)
    sharded_m = fresh_model(); sharded_m.world_size = world
    steps(sharded_m, sb, jitter[sl].contiguous(), noise[sl].contiguous())
    torch.cuda.synchronize()
    # the same steps with the NCCL collectives instead of the peer-memory kernels (both paths must agree; with two ranks the
    # two-term sums are identical whatever the order)
    nccl_m = fresh_model(); nccl_m.world_size = world; nccl_m.optimizer.use_peer = False
    steps(nccl_m, sb, jitter[sl].contiguous(), noise[sl].contiguous())
    torch.cuda.synchronize()
    if rank == 0:
        ret["peer_optimizer_used"] = bool(sharded_m.optimizer._peer)
        ret["peer_vs_nccl_fp16_equal_frac"] = float((sharded_m.optimizer.flat_h[: single.optimizer.n] == nccl_m.optimizer.flat_h[: single.optimizer.n]).float().mean())
    h_single, h_shard = single.optimizer.flat_h[: single.optimizer.n].float(), sharded_m.optimizer.flat_h[: single.optimizer.n].float()
    gathered = [torch.empty_like(h_shard) for _ in range(world)]
    dist.all_gather(gathered, h_shard)
    sharded_m.optimizer.gather_master_params(world)
    p_single, p_shard = single.optimizer.flat_p[: single.optimizer.n], sharded_m.optimizer.flat_p[: single.optimizer.n]
    if rank == 0:
        ret["fp16_image_same_on_ranks"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
        ret["adam_steps"] = (single.optimizer.step_count, sharded_m.optimizer.step_count)
        ret["param_max_diff"] = float((p_single - p_shard).abs().max()); ret["param_mean_diff"] = float((p_single - p_shard).abs().mean())
        ret["param_moved"] = float((p_single - torch.cat([torch.from_numpy(enc), torch.from_numpy(col)]).to(dev)).abs().max())
        ret["fp16_frac_diff"] = float((h_single != h_shard).float().mean())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_frame_and_gradient():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["hit"] > 1000
    assert ret["frame_equal"], "cooperative frame differs from the single-GPU frame"
    assert ret.get("peer_frames_equal") is not False, "peer-memory frame differs from the single-GPU frame"
    if ret.get("peer_frames_equal") is None:
        print("peer memory unavailable on this box:", ret.get("peer_error"))
    assert ret["gnorm"] > 0 and ret["rel_enc"] < 1e-3 and ret["rel_col"] < 1e-3, dict(ret)
    # sharded optimiser == replicated optimiser up to the summation order of the gradient (Adam normalises the update, so
    # an entry whose gradient is pure rounding noise can move by a full lr step in OPPOSITE directions in the two runs:
    # the worst case is 2 x 3 steps x lr = 6e-2; the mean difference must be tiny)
    assert ret["fp16_image_same_on_ranks"] and ret["adam_steps"] == (3, 3), dict(ret)
    assert ret["param_moved"] > 1e-3 and ret["param_max_diff"] <= 6.1e-2 and ret["param_mean_diff"] < 1e-4, dict(ret)
    # fraction of fp16 entries that differ AT ALL after 3 steps: entries with noise-level gradients, whose normalised update
    # depends on the float-atomic summation order (run-to-run variable, observed 0.8-1.4 %)
    assert ret["fp16_frac_diff"] < 5e-2, dict(ret)
    print("sharded optimiser:", {k: ret[k] for k in ("peer_optimizer_used", "peer_vs_nccl_fp16_equal_frac", "param_max_diff", "param_mean_diff", "fp16_frac_diff")})
    assert ret["peer_vs_nccl_fp16_equal_frac"] > 0.9999, dict(ret)   # entries whose fp16 image differs at all (noise-level gradients)
