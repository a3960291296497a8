"""ctypes binding of oracle/libia_oracle.so (built by `make -C oracle`)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
f32 = np.float32


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libia_oracle.so")
    src = os.path.join(_HERE, "ia_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libia_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _c(a, dt=f32):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(C.c_int(n))


def precompute(voxel_w, tfs, offset, scale, D, H, W):
    voxel_w, tfs, offset, scale = _c(voxel_w), _c(tfs), _c(offset), _c(scale)
    vd = np.empty((3, D, H, W), f32); vJ = np.empty((12, D, H, W), f32)
    lib().orc_precompute(_p(voxel_w), _p(tfs), _p(offset), _p(scale), C.c_int(D), C.c_int(H), C.c_int(W), _p(vd), _p(vJ))
    return vd, vJ


def broyden(xd, voxel_J, tfs, bone_ids, offset, scale, cvg=1e-5, dvg=1e-1, want_jinv=True):
    xd, voxel_J, tfs, offset, scale = _c(xd).reshape(-1, 3), _c(voxel_J), _c(tfs), _c(offset), _c(scale)
    bone_ids = _c(bone_ids, np.int32)
    M, I = len(xd), len(bone_ids)
    _, D, H, W = voxel_J.shape
    xc = np.empty((M, I, 3), f32); valid = np.empty((M, I), np.uint8); iters = np.empty((M, I), np.int32)
    jinv = np.empty((M, I, 3, 3), f32) if want_jinv else None
    lib().orc_broyden(_p(xd), C.c_long(M), _p(voxel_J), C.c_int(D), C.c_int(H), C.c_int(W), _p(tfs),
                      _p(bone_ids, C.c_int), C.c_int(I), _p(offset), _p(scale), C.c_float(cvg), C.c_float(dvg),
                      _p(xc), _p(jinv) if want_jinv else None, _p(valid, C.c_uint8), _p(iters, C.c_int32))
    return xc, jinv, valid.astype(bool), iters


def filter_roots(xc, valid):
    xc = _c(xc); v = _c(valid, np.uint8)
    M, I = v.shape
    out = np.empty((M, I), np.uint8)
    lib().orc_filter(_p(xc), _p(v, C.c_uint8), C.c_long(M), C.c_int(I), _p(out, C.c_uint8))
    return out.astype(bool)


def hashgrid_layout():
    res = np.empty(16, np.uint32); size = np.empty(16, np.uint32); off = np.empty(16, np.uint32)
    ls = np.empty(16, f32); tot = C.c_uint32(0)
    lib().orc_hashgrid_layout(_p(res, C.c_uint32), _p(ls), _p(size, C.c_uint32), _p(off, C.c_uint32), C.byref(tot))
    return {"res": res, "scale": ls, "size": size, "offset": off, "total": int(tot.value)}


def ngp_forward(x, center, scale, enc_params, col_params, emulate=True, want_feat=False):
    """emulate: 0/False fp32, 1/True fp16 values + fp32 accumulation (the product's model), 2 tiny-cuda-nn-like fp16 accumulation"""
    x = _c(x).reshape(-1, 3); P = len(x)
    center, scale, enc_params, col_params = _c(center), _c(scale), _c(enc_params), _c(col_params)
    sigma = np.empty(P, f32); rgb = np.empty((P, 3), f32)
    feat = np.empty((P, 16), f32) if want_feat else None
    lib().orc_ngp_forward(_p(x), C.c_long(P), _p(center), _p(scale), _p(enc_params), _p(col_params),
                          C.c_int(int(emulate)), _p(sigma), _p(rgb), _p(feat) if want_feat else None)
    return (sigma, rgb, feat) if want_feat else (sigma, rgb)


def raymarch_train(rays_o, rays_d, nears, fars, grid, scale, offset, step_size, N_steps):
    rays_o, rays_d, nears, fars, scale, offset, step_size = map(_c, (rays_o, rays_d, nears, fars, scale, offset, step_size))
    g = _c(grid, np.uint8); N = len(rays_o)
    depths = np.empty((N, N_steps), f32)
    lib().orc_raymarch_train(_p(rays_o), _p(rays_d), _p(nears), _p(fars), C.c_long(N), _p(g, C.c_uint8),
                             C.c_int(g.shape[0]), _p(scale), _p(offset), _p(step_size), C.c_int(N_steps), _p(depths))
    return depths


def raymarch_test(rays_o, rays_d, nears, fars, alive, grid, scale, offset, step_size, N_steps):
    """nears is mutated in place (must be a contiguous float32 array)."""
    assert nears.dtype == f32 and nears.flags.c_contiguous
    rays_o, rays_d, fars, scale, offset, step_size = map(_c, (rays_o, rays_d, fars, scale, offset, step_size))
    alive = _c(alive, np.int64); g = _c(grid, np.uint8); A = len(alive)
    pts = np.empty((A, N_steps, 3), f32); deltas = np.empty((A, N_steps), f32); depths = np.empty((A, N_steps), f32)
    lib().orc_raymarch_test(_p(rays_o), _p(rays_d), _p(nears), _p(fars), _p(alive, C.c_int64), C.c_long(A),
                            _p(g, C.c_uint8), C.c_int(g.shape[0]), _p(scale), _p(offset), _p(step_size),
                            C.c_int(N_steps), _p(pts), _p(deltas), _p(depths))
    return pts, deltas, depths


def composite_test(rgb_vals, sigma_vals, delta_vals, depth_vals, alive, color, depth, nohit, thresh):
    """color/depth/nohit mutated in place."""
    for a in (color, depth, nohit):
        assert a.dtype == f32 and a.flags.c_contiguous
    rgb_vals, sigma_vals, delta_vals, depth_vals = map(_c, (rgb_vals, sigma_vals, delta_vals, depth_vals))
    alive = _c(alive, np.int64); A = len(alive); N_steps = sigma_vals.shape[1] if A else 0
    lib().orc_composite_test(_p(rgb_vals), _p(sigma_vals), _p(delta_vals), _p(depth_vals), _p(alive, C.c_int64),
                             C.c_long(A), C.c_int(N_steps), _p(color), _p(depth), _p(nohit), C.c_float(thresh))
