"""GPU: SMPLDeformer path (second half; deformers/smpl_deformer.py, `fit.py deformer=smpl`) -- the
nearest-vertex kernel against brute-force PyTorch, the deformer against a literal PyTorch restatement, and the
kernel-for-kernel render / training paths with a differentiable `model(pts, None)`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _deformer():
    import torch
    from instantavatar_b200 import synthetic
    from instantavatar_b200.deformers.smpl_deformer import SMPLDeformer
    d = SMPLDeformer(smpl_data=synthetic.smpl_dict_cached(0))
    pose = {k: torch.from_numpy(v).cuda() for k, v in synthetic.load_pose(0).items()}
    d.body_model = d.body_model.cuda()
    d.prepare_deformer(pose)
    return d, pose


def _net(d, betas):
    import torch
    from instantavatar_b200 import synthetic
    from instantavatar_b200.models.networks.ngp import NeRFNGPNet
    net = NeRFNGPNet(None).cuda()
    net.initialize(d.bbox)
    bbox = d.bbox.cpu().numpy().astype(np.float64)
    enc, col = synthetic.analytic_avatar_params(_template_joints(d, betas), (bbox[0] + bbox[1]) / 2, bbox[1] - bbox[0])
    net.load_flat_params(torch.from_numpy(enc).cuda(), torch.from_numpy(col).cuda())
    return net


def _template_joints(d, betas):
    import math
    import torch
    body_pose_t = torch.zeros((1, 69), device="cuda")
    body_pose_t[:, 2] = math.pi / 6; body_pose_t[:, 5] = -math.pi / 6
    return d.body_model(betas=betas[:1], body_pose=body_pose_t).joints[0].cpu().numpy()


def test_knn1_matches_bruteforce():
    import torch
    from instantavatar_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    verts = torch.rand((6890, 3), device="cuda", generator=g) * 2 - 1
    for n in (1, 33, 10007):
        pts = torch.rand((n, 3), device="cuda", generator=g) * 2.4 - 1.2
        d2, idx = ops.knn1(pts, verts)
        diff = pts[:, None, :] - verts[None]
        ref = (diff * diff).sum(-1)
        rd, ri = ref.min(dim=1)
        np.testing.assert_allclose(d2.cpu().numpy(), rd.cpu().numpy(), rtol=1e-6, atol=1e-12)
        # the index may differ only where two vertices are equidistant to rounding
        srt = ref.sort(dim=1).values
        clear = (srt[:, 1] - srt[:, 0]) > 1e-6 * srt[:, 1]
        assert torch.equal(idx[clear], ri[clear])
    d2, idx = ops.knn1(torch.zeros((0, 3), device="cuda"), verts)
    assert d2.numel() == 0 and idx.numel() == 0
    # more vertices than one shared-memory tile, exact duplicate vertices: the earlier one wins
    big = torch.rand((20000, 3), device="cuda", generator=g)
    big[15000] = big[3]
    d2, idx = ops.knn1(big[3:4] + 0.0, big)
    assert idx.item() == 3 and d2.item() == 0.0


def test_deform_matches_torch_restatement():
    import torch
    d, pose = _deformer()
    g = torch.Generator(device="cuda").manual_seed(1)
    v = d.vertices[0]
    pick = torch.randint(0, v.shape[0], (5000,), device="cuda", generator=g)
    near = v[pick] + 0.01 * torch.randn((5000, 3), device="cuda", generator=g)
    far = torch.rand((3000, 3), device="cuda", generator=g) * 2 - 1
    pts = torch.cat([near, far])
    cano, valid = d.deform(pts)
    # literal restatement of smpl_deformer.py:87-110 with a brute-force nearest neighbour
    diff = pts[:, None, :] - v[None]
    dist_sq, idx = (diff * diff).sum(-1).min(dim=1)
    ref_valid = dist_sq < d.threshold ** 2
    Tv = d.T_inv[0][idx]
    ref = (Tv[:, :3, :3] @ pts[:, :, None]).squeeze(-1) + Tv[:, :3, 3]
    same = torch.isclose(dist_sq, torch.full_like(dist_sq, d.threshold ** 2), rtol=1e-5)   # on the threshold: either way
    assert torch.equal(valid[~same], ref_valid[~same])
    assert valid[:5000].float().mean() > 0.95 and valid[5000:].float().mean() < 0.5
    # points whose nearest vertex is unambiguous map identically
    both = valid & ref_valid
    assert ((cano - ref).abs().max(-1).values[both] > 1e-5).float().mean() < 1e-3
    # a posed vertex maps to its template vertex
    c2, v2 = d.deform(v)
    assert v2.all() and (c2 - d.vs_template[0]).abs().max() < 1e-4


def test_render_and_training_paths_with_smpl_deformer():
    """Raymarcher on the kernel-for-kernel path (raymarch_test/composite_test window loop, raymarch_train + torch
    compositing) with SMPLDeformer + NeRFNGPNet: finite image with hits, and a training step whose loss gradient reaches
    the network parameters and -- through T_inv and the network's input gradient -- the SMPL pose."""
    import torch
    from instantavatar_b200 import synthetic
    from instantavatar_b200.deformers.smpl_deformer import SMPLDeformer
    from instantavatar_b200.models.dnerf import Rays
    from instantavatar_b200.renderers.raymarcher_acc import BoundModel, Raymarcher
    d, pose = _deformer()
    net = _net(d, pose["betas"])
    o, dd = synthetic.demo_camera_rays(512, 512)
    ys, xs = np.arange(128, 384, 8), np.arange(192, 320, 4)
    idx = (ys[:, None] * 512 + xs[None]).ravel()
    def make_rays():
        r = Rays(o=torch.from_numpy(o[idx][None]).cuda(), d=torch.from_numpy(dd[idx][None]).cuda(),
                 near=torch.zeros((1, len(idx)), device="cuda"), far=torch.ones((1, len(idx)), device="cuda"))
        d.transform_rays_w2s(r)
        return r
    rm = Raymarcher(256, 291600, device="cuda")
    rm.initialize(1)
    # ---- eval ----
    with torch.no_grad():
        rm.density_grid_test.initialize(d, net)
        out = rm(make_rays(), BoundModel(d, net, True), eval_mode=True)
    a = out["alpha_coarse"].reshape(-1)
    assert torch.isfinite(out["rgb_coarse"]).all() and a.min() >= -1e-6 and a.max() <= 1 + 1e-5
    assert (a > 0.5).sum() > 50, int((a > 0.5).sum())
    assert rm.density_grid_test.density_field.sum() > 100
    # ---- train ----
    p = {k: v.clone() for k, v in pose.items()}
    p["body_pose"].requires_grad_(True)
    d.prepare_deformer(p)
    with torch.no_grad():
        rm.density_grid_train.update(d, net, 0)
    assert rm.density_grid_train.density_field.sum() > 100
    net.zero_grad(set_to_none=True)
    pred = rm(make_rays(), BoundModel(d, net, False), eval_mode=False, noise=0, bg_color=None)
    assert pred["weight_coarse"].shape[-1] == 256
    target = out["rgb_coarse"].detach()
    loss = ((pred["rgb_coarse"] - target) ** 2).mean() + 0.1 * ((pred["alpha_coarse"] - out["alpha_coarse"].detach()) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and loss.item() < 0.05, loss.item()   # the two paths render the same avatar
    g_enc = net.encoder.params.grad
    assert g_enc is not None and torch.isfinite(g_enc).all() and g_enc.abs().sum() > 0
    g_pose = p["body_pose"].grad
    assert g_pose is not None and torch.isfinite(g_pose).all() and g_pose.abs().sum() > 0
