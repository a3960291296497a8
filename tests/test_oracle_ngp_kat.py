"""CPU: hash-grid + MLP known-answer test (SURVEY.md §8c golden vector 3).  The oracle's three rounding modes are pinned
to tests/golden/ngp_kat_golden.npz (generator next to it); the GPU product is compared with mode 1 on the same points in
tests/test_gpu_parity.py::test_ngp_kat_matches_oracle_mode1."""
import hashlib
import os

import numpy as np
import pytest

from oracle import capi
from oracle import scene as oscene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngp_kat_golden.npz")


def kat_points(subj, n=65536, seed=20260923):
    rng = np.random.default_rng(seed)
    lo, hi = subj.bbox[0], subj.bbox[1]
    x = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    j = subj.joints_cano.reshape(-1, 3)[:24]
    near = (j[rng.integers(0, 24, n // 4)] + rng.normal(0, 0.06, (n // 4, 3))).astype(np.float32)
    x[: n // 4] = np.clip(near, lo, hi)
    return x


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_oracle_reproduces_kat(mode):
    z = np.load(GOLD)
    subj = oscene.build_subject()
    net = oscene.build_net(subj)
    x = kat_points(subj, int(z["n_points"]), int(z["seed"]))
    np.testing.assert_array_equal(x[::8], z["x"])
    sigma, rgb = capi.ngp_forward(x, net.center, net.scale, net.enc, net.col, emulate=mode)
    if mode == 0:  # fp32 everywhere: expf may differ by an ulp between libm builds
        np.testing.assert_allclose(sigma[::8], z["sigma_mode0"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(rgb[::8], z["rgb_mode0"], rtol=0, atol=1e-6)
    else:  # outputs are fp16 values: exact
        np.testing.assert_array_equal(sigma[::8], z[f"sigma_mode{mode}"])
        np.testing.assert_array_equal(rgb[::8], z[f"rgb_mode{mode}"])
        digest = np.frombuffer(hashlib.sha256(sigma.tobytes() + rgb.tobytes()).digest(), np.uint8)
        np.testing.assert_array_equal(digest, z[f"sha256_mode{mode}"])


def test_mode_gaps_are_what_design_quotes():
    """DESIGN.md §3 quotes these bounds for 'how far could real tiny-cuda-nn be' at the network-output level"""
    z = np.load(GOLD)
    d21, d10 = z["delta_2_vs_1"], z["delta_1_vs_0"]
    assert d21[3] < 5e-3 and d21[1] < 6e-2      # rgb L-inf, relative sigma L-inf between tcnn-like and product rounding
    assert d10[3] < 2.5e-3 and d10[1] < 2e-2    # product rounding vs fp32
