"""GPU: training kernels (forward, compositing backward, network backward, Adam) against the CPU oracle and its
plain-PyTorch fp32 autograd reference."""
import numpy as np
import pytest

from oracle import render as orender
from oracle import scene as oscene
from oracle import testing as scene_util
from oracle import torch_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    return scene_util.oracle_scene(0)


@pytest.fixture(scope="module")
def dev(sc):
    import torch
    scene, extra = scene_util.upload(sc)
    torch.cuda.synchronize()
    return scene, extra


def patch_rays(sc, seed=0):
    """two 16x16 pixel patches on the body of the 512x512 demo camera"""
    fr = sc["frame"]
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    idx = []
    for (y0, x0) in ((200, 240), (300, 250)):
        ys, xs = np.arange(y0, y0 + 16), np.arange(x0, x0 + 16)
        idx.append((ys[:, None] * 512 + xs[None]).ravel())
    idx = np.concatenate(idx)
    rng = np.random.default_rng(seed)
    jitter = rng.random((len(idx), 256), dtype=np.float32)
    noise = rng.normal(0, 1, (len(idx), 256)).astype(np.float32)
    bg = rng.random((len(idx), 3), dtype=np.float32)
    return o[idx], d[idx], near[idx], far[idx], jitter, noise, bg


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_ngp_backward_matches_torch_autograd(sc, dev):
    import torch
    from instantavatar_b200 import ops
    scene, _ = dev
    net = sc["net"]
    rng = np.random.default_rng(11)
    v = sc["subj"].verts_cano
    n = 3001  # not a multiple of 32
    x = (v[rng.integers(0, len(v), n)] + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    dsig = (rng.normal(0, 1, n) * 1e-3).astype(np.float32)
    drgb = (rng.normal(0, 1, (n, 3)) * 1e-2).astype(np.float32)
    enc = torch.from_numpy(net.enc).requires_grad_(True); col = torch.from_numpy(net.col).requires_grad_(True)
    s, c = torch_ref.ngp_forward(torch.from_numpy(x), net.center, net.scale, enc, col, True)
    ((s * torch.from_numpy(dsig)).sum() + (c * torch.from_numpy(drgb)).sum()).backward()
    g_enc_ref, g_col_ref = enc.grad.numpy(), col.grad.numpy()

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g_enc = torch.zeros(net.enc.size, device="cuda"); g_col = torch.zeros(net.col.size, device="cuda")
    count = torch.tensor([n], device="cuda", dtype=torch.int32)
    ops.ngp_backward(scene, t(x), t(dsig), t(drgb), count, g_enc, g_col, 128.0)
    torch.cuda.synchronize()
    g_enc, g_col = g_enc.cpu().numpy(), g_col.cpu().numpy()
    # MLP weight gradients (fp16 dgrad chain: ~1e-3 relative per term)
    assert rel_err(g_col, g_col_ref) < 2e-2, rel_err(g_col, g_col_ref)
    assert rel_err(g_enc[:3072], g_enc_ref[:3072]) < 2e-2, rel_err(g_enc[:3072], g_enc_ref[:3072])
    # hash-grid gradients
    gg, gr = g_enc[3072:], g_enc_ref[3072:]
    assert rel_err(gg, gr) < 2e-2, rel_err(gg, gr)
    assert np.array_equal(gg != 0, gr != 0) or np.mean((gg != 0) != (gr != 0)) < 1e-4
    # the unused output rows of W5 (3..15) receive no gradient
    assert np.all(g_col[5120 + 3 * 64:] == 0)


def _oracle_train(sc, rays):
    o, d, near, far, jitter, noise, bg = rays
    fr = sc["frame"]
    return orender.render_train(o, d, near, far, sc["occ"], fr["bbox_deformed"][0], fr["bbox_deformed"][1],
                                scene_util.oracle_model_aux(sc, False), jitter, noise, bg, return_aux=True)


def test_train_fwd_matches_oracle(sc, dev):
    import torch
    from instantavatar_b200 import ops
    scene, _ = dev
    rays = patch_rays(sc)
    ref = _oracle_train(sc, rays)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o, d, near, far, jitter, noise, bg = rays
    stats = ops.new_stats("cuda")
    out, saved = ops.train_fwd(scene, t(o), t(d), t(near), t(far), t(bg), t(jitter), t(noise), stats)
    torch.cuda.synchronize()
    st = ops.stats_dict(stats)
    assert st["samples"] == int(ref["mask"].sum())
    cnt = saved["count"].cpu().numpy()
    np.testing.assert_array_equal(cnt, ref["mask"].sum(-1))
    # sample depths / positions are bit-exact
    z = saved["z"].cpu().numpy()
    for r in range(0, len(cnt), 37):
        np.testing.assert_array_equal(z[r, :cnt[r]], ref["z"][r][ref["mask"][r]])
    w = out["weights"].cpu().numpy()
    assert np.abs(w - ref["weights"]).max() < 2e-3
    assert np.abs(out["rgb"].cpu().numpy() - ref["rgb"]).max() < 2e-3
    assert np.abs(out["alpha"].cpu().numpy() - ref["alpha"]).max() < 2e-3
    assert np.abs(out["depth"].cpu().numpy() - ref["depth"]).max() < 1e-2
    assert (ref["alpha"] > 0.5).sum() > 100


def test_train_fwd_split_equals_fused_bit_for_bit(sc, dev):
    """ia_train_fwd_split (march -> sample list -> point query -> compositing; the default) and the one-kernel ia_train_fwd
    must agree in every output and every saved-for-backward value, for every tile shape of the fused kernel"""
    import torch
    from instantavatar_b200 import ops
    scene, _ = dev
    rays = patch_rays(sc, seed=4)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o, d, near, far, jitter, noise, bg = rays
    res = {}
    try:
        for name, split, trpw, k in (("split", 1, 2, 1), ("split_k2", 1, 2, 2), ("split_k4", 1, 2, 4), ("split_auto", 1, 2, 0),
                                     ("fused2", 0, 2, 0), ("fused1", 0, 1, 0)):
            ops.set_option("train_split", split); ops.set_option("train_rays_per_warp", trpw)
            ops.set_option("query_lanes_per_sample", k)   # lanes sharing a sample's 13 root finds in the split form's query
            stats = ops.new_stats("cuda")
            out, saved = ops.train_fwd(scene, t(o), t(d), t(near), t(far), t(bg), t(jitter), t(noise), stats)
            torch.cuda.synchronize()
            res[name] = (out, saved, ops.stats_dict(stats))
    finally:
        ops.set_option("train_split", 1); ops.set_option("train_rays_per_warp", 2); ops.set_option("query_lanes_per_sample", 0)
    out0, saved0, st0 = res["split"]
    assert st0["samples"] > 1000
    cnt = saved0["count"].long()
    live = torch.arange(saved0["sigma"].shape[1], device="cuda")[None] < cnt[:, None]   # slots the forward filled
    for name in ("split_k2", "split_k4", "split_auto", "fused2", "fused1"):
        out1, saved1, st1 = res[name]
        assert st1["samples"] == st0["samples"] and st1["net_evals"] == st0["net_evals"] and st1["field_loads"] == st0["field_loads"]
        for k in out0:
            assert torch.equal(out0[k], out1[k]), (name, k)
        assert torch.equal(saved0["count"], saved1["count"]) and torch.equal(saved0["best"], saved1["best"])
        for k in ("sigma", "z"):
            assert torch.equal(saved0[k][live], saved1[k][live]), (name, k)
        for k in ("rgb", "xc"):
            assert torch.equal(saved0[k][live], saved1[k][live]), (name, k)


def test_train_backward_matches_oracle(sc, dev):
    import torch
    from instantavatar_b200 import ops
    scene, _ = dev
    net = sc["net"]
    rays = patch_rays(sc, seed=1)
    o, d, near, far, jitter, noise, bg = rays
    ref = _oracle_train(sc, rays)
    rng = np.random.default_rng(3)
    tgt_rgb = rng.random((len(o), 3), dtype=np.float32); tgt_a = (rng.random(len(o)) > 0.3).astype(np.float32)
    # ---- reference gradients: oracle forward (numpy) + differentiable tail in plain PyTorch ----
    enc = torch.from_numpy(net.enc).requires_grad_(True); col = torch.from_numpy(net.col).requires_grad_(True)
    step = ((far - near) / np.float32(256)).astype(np.float32)
    outs = torch_ref.render_train_torch(ref, net, enc, col, step, bg, noise)
    loss_ref = torch_ref.nerf_loss(outs["rgb"], outs["alpha"], outs["weights"], torch.from_numpy(tgt_rgb), torch.from_numpy(tgt_a))
    loss_ref.backward()
    # ---- CUDA path through the custom autograd Function ----
    from instantavatar_b200.autograd import _RenderTrain
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    enc_g = t(net.enc).requires_grad_(True); col_g = t(net.col).requires_grad_(True)
    rgb, depth, alpha, weights = _RenderTrain.apply(enc_g, col_g, scene, t(o), t(d), t(near), t(far), t(bg), t(jitter), t(noise), None)
    loss = torch_ref.nerf_loss(rgb, alpha, weights, t(tgt_rgb), t(tgt_a))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * max(1.0, abs(loss_ref.item()))
    g_enc, g_col = enc_g.grad.cpu().numpy(), col_g.grad.cpu().numpy()
    g_enc_ref, g_col_ref = enc.grad.numpy(), col.grad.numpy()
    assert np.linalg.norm(g_col_ref) > 0 and np.linalg.norm(g_enc_ref[3072:]) > 0
    assert rel_err(g_col, g_col_ref) < 5e-2, rel_err(g_col, g_col_ref)
    assert rel_err(g_enc[:3072], g_enc_ref[:3072]) < 5e-2, rel_err(g_enc[:3072], g_enc_ref[:3072])
    assert rel_err(g_enc[3072:], g_enc_ref[3072:]) < 5e-2, rel_err(g_enc[3072:], g_enc_ref[3072:])


def test_adam_matches_torch(dev):
    import torch
    from instantavatar_b200 import ops
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n, device="cuda"); p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p = p0.clone(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    found = torch.zeros(1, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * (10.0 ** -step)
        g[::7] = 0  # untouched entries keep decaying momentum (dense semantics)
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g * 1024.0, m, v, 1e-2, (0.9, 0.99), 1e-15, step, 1.0 / 1024.0, found)
    torch.cuda.synchronize()
    assert torch.allclose(p, p_ref.detach(), rtol=1e-5, atol=1e-6), (p - p_ref.detach()).abs().max()
    # inf gradients: GradScaler semantics -> step skipped
    g = torch.randn(n, device="cuda"); g[5] = float("inf")
    ops.grad_check_finite(g, found)
    before = p.clone()
    ops.adam_step(p, g, m, v, 1e-2, (0.9, 0.99), 1e-15, 6, 1.0, found)
    torch.cuda.synchronize()
    assert found.item() == 1.0 and torch.equal(p, before)


def test_adam_device_state_matches_torch(dev):
    """graph-friendly variant: step count / bias corrections / 1/scale on the device, fused zero-grad + fp16 refresh"""
    import torch
    from instantavatar_b200 import ops
    torch.manual_seed(1)
    n = 40000
    p0 = torch.randn(n, device="cuda"); p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p = p0.clone(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    state = torch.tensor([1e-2, 0.9, 0.99, 1e-15, 0, 1, 1, 1], dtype=torch.float32).cuda()
    found = torch.zeros(1, device="cuda"); scale = torch.full((1,), 1024.0, device="cuda")
    half = torch.zeros(n - 8, device="cuda", dtype=torch.float16)
    for step in range(1, 5):
        g = torch.randn(n, device="cuda") * 0.1
        p_ref.grad = g.clone(); opt.step()
        gs = g * 1024.0 * 2  # scaled loss, summed over 2 ranks
        ops.adam_prepare(state, 0.5, scale, found)
        ops.adam_step_dev(p, gs, m, v, state, found, half, 8)
        assert not gs.any()  # gradient consumed and zeroed
    torch.cuda.synchronize()
    assert state[4].item() == 4.0
    assert torch.allclose(p, p_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(half, p[8:].half())
    found.fill_(1.0)
    before = p.clone()
    ops.adam_prepare(state, 0.5, scale, found)
    ops.adam_step_dev(p, torch.ones(n, device="cuda"), m, v, state, found, half, 8)
    torch.cuda.synchronize()
    assert torch.equal(p, before) and state[4].item() == 4.0
