"""GPU: the in-repo `tinycudann` module (tcnn.NetworkWithInputEncoding / tcnn.Network on libia_b200.so) used exactly as the
reference's network does (models/networks/ngp.py:73-83): forward identical to the fused NeRFNGPNet kernels, gradients against
the plain-PyTorch fp32 autograd reference (oracle/torch_ref.py) and against the fused backward."""
import numpy as np
import pytest

from oracle import testing as scene_util
from oracle import torch_ref

pytestmark = pytest.mark.gpu

ENC_CFG = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 1.5}
NET1 = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1}
NET2 = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64, "n_hidden_layers": 2}


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


class RefShapedNet:
    """the forward of the reference's NeRFNGPNet (ngp.py:73-83) on the two tcnn modules"""

    def __init__(self, net, device="cuda"):
        import torch
        import tinycudann as tcnn
        self.encoder = tcnn.NetworkWithInputEncoding(n_input_dims=3, n_output_dims=16, encoding_config=ENC_CFG, network_config=NET1).to(device)
        self.color_net = tcnn.Network(n_input_dims=15, n_output_dims=3, network_config=NET2).to(device)
        with torch.no_grad():
            self.encoder.params.copy_(torch.from_numpy(net.enc)); self.color_net.params.copy_(torch.from_numpy(net.col))
        self.center, self.scale = torch.from_numpy(net.center).to(device), torch.from_numpy(net.scale).to(device)

    def __call__(self, x):
        x = (x - self.center) / self.scale + 0.5
        x = x.clamp(min=0, max=1)
        x = self.encoder(x)
        sigma = x[..., 0]
        color = self.color_net(x[..., 1:]).float()
        return color, sigma.float()


def test_shim_forward_equals_fused_kernels():
    import torch
    from instantavatar_b200 import ops
    sc = scene_util.oracle_scene(0)
    scene, _ = scene_util.upload(sc)
    net = sc["net"]
    m = RefShapedNet(net)
    assert m.encoder.params.dtype == torch.float32 and m.encoder.params.numel() == 3072 + 2 * 6513496 and m.color_net.params.numel() == 6144
    rng = np.random.default_rng(5)
    v = sc["subj"].verts_cano
    for n in (1, 31, 4097):
        x = torch.from_numpy((v[rng.integers(0, len(v), n)] + rng.normal(0, 0.03, (n, 3))).astype(np.float32)).cuda()
        color, sigma = m(x)
        out16 = m.encoder(((x - m.center) / m.scale + 0.5).clamp(0, 1))
        assert out16.dtype == torch.float16 and out16.shape == (n, 16) and color.shape == (n, 3)
        c_ref, s_ref = ops.ngp_forward(scene, x)
        assert torch.equal(sigma, s_ref) and torch.equal(color, c_ref)   # the same arithmetic, cut at the module boundary
    e = m.encoder(torch.empty((0, 3), device="cuda"))
    assert e.shape == (0, 16)
    # batched leading dims as the reference passes them
    xb = torch.rand((2, 7, 3), device="cuda")
    assert m.encoder(xb).shape == (2, 7, 16) and m.color_net(torch.rand((2, 7, 15), device="cuda")).shape == (2, 7, 3)


def test_shim_gradients_match_torch_autograd_and_fused_backward():
    import torch
    from instantavatar_b200 import ops
    sc = scene_util.oracle_scene(0)
    scene, _ = scene_util.upload(sc)
    net = sc["net"]
    m = RefShapedNet(net)
    rng = np.random.default_rng(11)
    v = sc["subj"].verts_cano
    n = 3001
    x = (v[rng.integers(0, len(v), n)] + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    dsig = (rng.normal(0, 1, n) * 1e-3).astype(np.float32)
    drgb = (rng.normal(0, 1, (n, 3)) * 1e-2).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # ---- through the shim, with autograd as the reference's training loop uses it ----
    xs = t(x).requires_grad_(True)
    color, sigma = m(xs)
    ((sigma * t(dsig)).sum() + (color * t(drgb)).sum()).backward()
    g_enc, g_col, g_x = m.encoder.params.grad.cpu().numpy(), m.color_net.params.grad.cpu().numpy(), xs.grad.cpu().numpy()
    # ---- fp32 PyTorch reference ----
    enc = torch.from_numpy(net.enc).requires_grad_(True); col = torch.from_numpy(net.col).requires_grad_(True)
    xr = torch.from_numpy(x).requires_grad_(True)
    s, c = torch_ref.ngp_forward(xr, net.center, net.scale, enc, col, True, pos_grad=True)
    ((s * torch.from_numpy(dsig)).sum() + (c * torch.from_numpy(drgb)).sum()).backward()
    assert rel_err(g_col, col.grad.numpy()) < 2e-2
    assert rel_err(g_enc[:3072], enc.grad.numpy()[:3072]) < 2e-2
    assert rel_err(g_enc[3072:], enc.grad.numpy()[3072:]) < 2e-2
    assert rel_err(g_x, xr.grad.numpy()) < 5e-2
    assert np.all(g_col[5120 + 3 * 64:] == 0)
    # ---- the fused backward on the same inputs (same dgrad chain, no module boundary) ----
    f_enc = torch.zeros(net.enc.size, device="cuda"); f_col = torch.zeros(net.col.size, device="cuda")
    ops.ngp_backward(scene, t(x), t(dsig), t(drgb), torch.tensor([n], device="cuda", dtype=torch.int32), f_enc, f_col, 128.0)
    assert rel_err(g_col, f_col.cpu().numpy()) < 2e-3 and rel_err(g_enc, f_enc.cpu().numpy()) < 5e-3
    # parameter updates are seen by the next forward (fp16 copies refresh on `params` version change)
    with torch.no_grad():
        m.color_net.params.mul_(0.5)
    color2, _ = m(t(x))
    assert not torch.equal(color2, color.detach())


def test_shim_rejects_unbuilt_configurations():
    import tinycudann as tcnn
    with pytest.raises(NotImplementedError):
        tcnn.NetworkWithInputEncoding(3, 16, dict(ENC_CFG, n_levels=8), NET1)
    with pytest.raises(NotImplementedError):
        tcnn.Network(15, 3, dict(NET2, n_neurons=128))
    with pytest.raises(NotImplementedError):
        tcnn.Network(16, 3, NET2)
