"""CPU: host-side mirror classes that need no GPU -- losses and the SMPL parameter embedding
(instant_avatar/utils/loss.py, models/structures/body_model_param.py)."""
import numpy as np
import pytest
import torch

from instantavatar_b200.models.structures.body_model_param import SMPLParamEmbedding
from instantavatar_b200.utils_loss import NeRFLoss, NGPLoss
from oracle import torch_ref


def _predictions(shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = {"rgb_coarse": torch.rand(*shape, 3, generator=g), "alpha_coarse": torch.rand(*shape, generator=g),
         "depth_coarse": torch.rand(*shape, generator=g) * 3, "weight_coarse": torch.rand(*shape, 16, generator=g) * 0.2}
    t = {"rgb": torch.rand(*shape, 3, generator=g), "alpha": (torch.rand(*shape, generator=g) > 0.4).float()}
    return p, t


def test_nerf_loss_matches_oracle_restatement():
    p, t = _predictions((1, 512))
    got = NeRFLoss()(p, t)["loss"]
    ref = torch_ref.nerf_loss(p["rgb_coarse"], p["alpha_coarse"], p["weight_coarse"], t["rgb"], t["alpha"])
    assert abs(got.item() - ref.item()) < 1e-6


def test_ngp_loss_reduces_to_nerf_loss_and_adds_depth_term_on_patches():
    p, t = _predictions((1, 4, 8, 8))
    base = NeRFLoss()(p, t)["loss"]
    assert abs(NGPLoss()(p, t)["loss"].item() - base.item()) < 1e-6      # SNARF_NGP_refine.yaml weights
    out = NGPLoss(w_depth_reg=0.01)(p, t)
    a, d = p["alpha_coarse"], p["depth_coarse"]
    avg = (d * a).sum(dim=(-1, -2)) / (a.sum(dim=(-1, -2)) + 1e-3)
    expect = (a * (d - avg[..., None, None]).abs()).mean()
    assert abs(out["loss_depth_reg"].item() - expect.item()) < 1e-7
    assert abs(out["loss"].item() - (base + 0.01 * expect).item()) < 1e-6
    # per-ray (non-patch) predictions skip the patch terms, as in the reference
    p1, t1 = _predictions((1, 256))
    assert "loss_depth_reg" not in NGPLoss(w_depth_reg=0.01)(p1, t1)
    with pytest.raises(RuntimeError):
        NGPLoss(w_lpips=0.01)
    stub = lambda x, y: (x - y).abs().mean(dim=(1, 2, 3))
    out = NGPLoss(w_lpips=0.5, lpips=stub)(p, t)
    assert out["loss_lpips"].item() > 0 and out["loss"].item() > base.item()


def test_smpl_param_embedding_lookup_and_tv_loss():
    F_ = 5
    g = torch.Generator().manual_seed(1)
    init = {"betas": torch.rand(1, 10, generator=g), "global_orient": torch.rand(F_, 3, generator=g),
            "body_pose": torch.rand(F_, 69, generator=g), "transl": torch.rand(F_, 3, generator=g)}
    emb = SMPLParamEmbedding(**init)
    idx = torch.tensor([3])
    out = emb(idx)
    assert out["betas"].shape == (1, 10) and torch.equal(out["betas"], init["betas"])
    for k in ("global_orient", "body_pose", "transl"):
        assert torch.equal(out[k], init[k][3:4])
        assert getattr(emb, k).weight.requires_grad
    # temporal smoothness: squared differences to both neighbours, clamped at the ends
    expect = sum(((init[k][3] - init[k][2]) ** 2).mean() + ((init[k][4] - init[k][3]) ** 2).mean()
                 for k in ("global_orient", "body_pose", "transl"))
    assert abs(emb.tv_loss(idx).item() - expect.item()) < 1e-6
    end = torch.tensor([F_ - 1])
    expect_end = sum(((init[k][F_ - 1] - init[k][F_ - 2]) ** 2).mean() for k in ("global_orient", "body_pose", "transl"))
    assert abs(emb.tv_loss(end).item() - expect_end.item()) < 1e-6
    out["body_pose"].sum().backward()
    assert emb.body_pose.weight.grad[3].abs().sum() > 0 and emb.body_pose.weight.grad[0].abs().sum() == 0


def test_smpl_deformer_inverse_transforms_map_posed_vertices_to_the_template():
    """SMPLDeformer.prepare_deformer (smpl_deformer.py:50-76) on the CPU: T_inv of vertex i takes the posed vertex (root
    frame) to the template-pose vertex, blend shapes removed and re-applied; the ray transform is the inverse root
    transform"""
    from instantavatar_b200 import synthetic
    from instantavatar_b200.deformers.smpl_deformer import SMPLDeformer
    d = SMPLDeformer(smpl_data=synthetic.smpl_dict_cached(0))
    pose = {k: torch.from_numpy(v) for k, v in synthetic.load_pose(57).items()}
    d.prepare_deformer(pose)
    v = d.vertices[0]
    cano = (d.T_inv[0][:, :3, :3] @ v[..., None]).squeeze(-1) + d.T_inv[0][:, :3, 3]
    assert (cano - d.vs_template[0]).abs().max() < 1e-5
    out = d.body_model(betas=pose["betas"], body_pose=pose["body_pose"], global_orient=pose["global_orient"], transl=pose["transl"])
    assert torch.allclose(d.w2s @ out.A[:, 0], torch.eye(4)[None], atol=1e-5)
    bb = d.get_bbox_deformed()
    assert bb.shape == (2, 3) and (bb[1] > bb[0]).all()
    with pytest.raises(ValueError):
        SMPLDeformer(smpl_data=synthetic.smpl_dict_cached(0), k=3)


def test_smpl_deformer_host_logic_against_literal_restatement(monkeypatch):
    """SMPLDeformer.deform / __call__ with the CUDA nearest-vertex kernel replaced by a brute-force torch search: the
    gathered inverse transforms, the validity rule, the train / eval fill values and the gradient path to the pose equal
    a literal restatement of smpl_deformer.py:60-132"""
    from instantavatar_b200 import ops, synthetic
    from instantavatar_b200.deformers.smpl_deformer import SMPLDeformer

    def knn1_bruteforce(pts, verts):
        diff = pts[:, None, :] - verts[None]
        return (diff * diff).sum(-1).min(dim=1)

    monkeypatch.setattr(ops, "knn1", knn1_bruteforce)
    d = SMPLDeformer(smpl_data=synthetic.smpl_dict_cached(0), threshold=0.05)
    pose = {k: torch.from_numpy(v) for k, v in synthetic.load_pose(20).items()}
    pose["body_pose"].requires_grad_(True)
    d.prepare_deformer(pose)
    # literal restatement of prepare_deformer
    out = d.body_model(betas=pose["betas"], body_pose=pose["body_pose"], global_orient=pose["global_orient"], transl=pose["transl"])
    s2w = out.A[:, 0]
    w2s = torch.inverse(s2w)
    T_inv = torch.inverse(out.T.float()).clone() @ s2w[:, None]
    T_inv[..., :3, 3] += d.pose_offset_t - out.pose_offsets
    T_inv[..., :3, 3] += d.shape_offset_t - out.shape_offsets
    T_inv = d.T_template @ T_inv
    verts = (out.vertices @ w2s[:, :3, :3].permute(0, 2, 1)) + w2s[:, None, :3, 3]
    assert torch.allclose(d.T_inv, T_inv, atol=1e-6) and torch.allclose(d.vertices, verts, atol=1e-6)
    g = torch.Generator().manual_seed(0)
    v = verts[0].detach()
    pts = torch.cat([v[torch.randint(0, v.shape[0], (400,), generator=g)] + 0.01 * torch.randn(400, 3, generator=g),
                     torch.rand(100, 3, generator=g) * 2 - 1])
    # literal deform
    dist_sq, idx = knn1_bruteforce(pts, v)
    valid_ref = dist_sq < 0.05 ** 2
    Tv = T_inv[0][idx]
    cano_ref = (Tv[..., :3, :3] @ pts[..., None]).squeeze(-1) + Tv[..., :3, 3]
    cano, valid = d.deform(pts)
    assert torch.equal(valid, valid_ref) and torch.allclose(cano, cano_ref, atol=1e-6)
    assert valid[:400].float().mean() > 0.9 and not valid.all()

    def model(x, _):
        return torch.sigmoid(x), x.sum(-1) * 10

    rgb_t, sig_t = d(pts, model, eval_mode=False)
    rgb_e, sig_e = d(pts, model, eval_mode=True)
    assert torch.all(sig_t[~valid] == -1e5) and torch.all(sig_e[~valid] == 0) and torch.all(rgb_t[~valid] == 0)
    assert torch.allclose(sig_t[valid], cano_ref[valid].sum(-1) * 10, atol=1e-5) and torch.allclose(rgb_e[valid], torch.sigmoid(cano_ref[valid]), atol=1e-6)
    # non-finite network outputs count as empty space in training mode only
    def bad_model(x, _):
        s = x.sum(-1)
        s = torch.where(torch.arange(len(s)) == 0, torch.full_like(s, float("nan")), s)
        return torch.sigmoid(x), s
    _, sig_bad = d(pts, bad_model, eval_mode=False)
    first = valid.nonzero()[0, 0]
    assert sig_bad[first] == -1e5
    sig_t[valid].sum().backward()
    assert pose["body_pose"].grad is not None and pose["body_pose"].grad.abs().sum() > 0


def test_snarf_deformer_frame_state_matches_oracle(monkeypatch):
    """SNARFDeformer.initialize / prepare_deformer (torch SMPL path) against the numpy oracle's SubjectOracle /
    prepare_frame (snarf_deformer.py:41-107): canonical bbox, voxel-grid normalisation, bone transforms, root transform,
    posed vertices; the CUDA field kernel is replaced by the oracle's C restatement."""
    from instantavatar_b200 import ops, synthetic
    from instantavatar_b200.deformers.snarf_deformer import SNARFDeformer
    from oracle import capi
    from oracle import testing as scene_util

    sc = scene_util.oracle_scene(0)
    subj, fr = sc["subj"], sc["frame"]

    def precompute(voxel_w, tfs, offset_k, scale_k, want_voxel_d=True):
        w = voxel_w.reshape(24, *voxel_w.shape[-3:]).numpy()
        vd, vJ = capi.precompute(w, tfs.detach().reshape(24, 4, 4).numpy(), offset_k.reshape(3).numpy(), scale_k.reshape(3).numpy(), *w.shape[1:])
        v = vd.reshape(3, -1)
        return torch.from_numpy(vJ), torch.from_numpy(vd), torch.from_numpy(np.concatenate([v.min(1), v.max(1)]))

    monkeypatch.setattr(ops, "precompute", precompute)
    d = SNARFDeformer(None, "male", {"cano_pose": "A_pose", "resolution": 128}, smpl_data=synthetic.smpl_dict_cached(0))
    d.fast_prepare = False   # torch SMPL forward (the one-launch kernel needs a GPU)
    pose = {k: torch.from_numpy(v) for k, v in sc["pose"].items()}
    d.initialize(pose["betas"], torch.device("cpu"), lbs_voxel=torch.from_numpy(subj.lbs_voxel))
    d.initialized = True
    d.prepare_deformer(pose)
    np.testing.assert_allclose(d.bbox.numpy(), subj.bbox, atol=1e-6)
    np.testing.assert_allclose(d.deformer.offset_kernel.reshape(3).numpy(), subj.offset_kernel, atol=1e-7)
    np.testing.assert_allclose(d.deformer.scale_kernel.reshape(3).numpy(), subj.scale_kernel, rtol=1e-6)
    np.testing.assert_allclose(d.tfs[0].numpy(), fr["tfs"], atol=5e-6)
    np.testing.assert_allclose(d.w2s[0].numpy(), fr["w2s"], atol=5e-6)
    np.testing.assert_allclose(d.vertices[0].numpy(), fr["vertices"], atol=1e-5)
    lo, hi = d.get_bbox_deformed()
    np.testing.assert_allclose(torch.stack([lo, hi]).numpy(), fr["bbox_deformed"], atol=1e-5)
