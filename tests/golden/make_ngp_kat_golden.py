"""Writes tests/golden/ngp_kat_golden.npz: the hash-grid + MLP known-answer test of SURVEY.md §8c (golden vector 3).

65 536 seeded canonical points through `orc_ngp_forward` (oracle/ia_oracle.c, restating models/networks/ngp.py:27-57,
73-83 and tiny-cuda-nn v1.6) with the deterministic synthetic network of the bench scene, in the three rounding modes:
  0  fp32 everywhere,
  1  fp16 tables / weights / activations, fp32 accumulation  (the CUDA product's model),
  2  tiny-cuda-nn-like: fp16 corner terms + fp16 running sum in the hash interpolation, fp16 accumulator fragments
     (rounded per 16-wide k-block) in every MLP layer  [TCNN-MEM].
Stored: every 8th point's outputs for the three modes (fp32), SHA-256 of the full 65 536-point outputs per mode, and the
mode-to-mode deviation statistics that DESIGN.md §3 quotes.  Run from the repository root (a few seconds)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import capi  # noqa: E402
from oracle import scene as oscene  # noqa: E402

N = 65536
subj = oscene.build_subject()
net = oscene.build_net(subj)
rng = np.random.default_rng(20260923)
# points: 3/4 uniform in the network's bounding box, 1/4 concentrated near the body surface (joints + noise)
lo, hi = subj.bbox[0], subj.bbox[1]
x = rng.uniform(lo, hi, (N, 3)).astype(np.float32)
j = subj.joints_cano.reshape(-1, 3)[:24]
near = (j[rng.integers(0, 24, N // 4)] + rng.normal(0, 0.06, (N // 4, 3))).astype(np.float32)
x[: N // 4] = np.clip(near, lo, hi)
out = {"x": x[::8].copy(), "n_points": np.array(N), "seed": np.array(20260923)}
res = {}
for mode in (0, 1, 2):
    sigma, rgb = capi.ngp_forward(x, net.center, net.scale, net.enc, net.col, emulate=mode)
    res[mode] = (sigma, rgb)
    out[f"sigma_mode{mode}"] = sigma[::8].copy()
    out[f"rgb_mode{mode}"] = rgb[::8].copy()
    out[f"sha256_mode{mode}"] = np.frombuffer(hashlib.sha256(sigma.tobytes() + rgb.tobytes()).digest(), np.uint8)
for a, b in ((1, 0), (2, 1), (2, 0)):
    ds = np.abs(res[a][0] - res[b][0]); dc = np.abs(res[a][1] - res[b][1]).max(-1)
    scale = np.maximum(np.abs(res[b][0]), 1.0)
    out[f"delta_{a}_vs_{b}"] = np.array([ds.max(), (ds / scale).max(), np.median(ds), dc.max(), np.median(dc)], np.float64)
    print(f"mode {a} vs {b}: sigma max|d| {ds.max():.4g} (rel {(ds / scale).max():.3g}, median {np.median(ds):.3g}); "
          f"rgb max|d| {dc.max():.4g} (median {np.median(dc):.3g})")
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ngp_kat_golden.npz"), **out)
