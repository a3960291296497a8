"""Builds the synthetic scene with the ORACLE on the CPU and uploads it for the CUDA product
(test infrastructure: used by tests/, smoke() and bench.py only; product code never imports oracle/)."""
import numpy as np

from . import render as orender
from . import scene as oscene

_CACHE = {}

# every 4th pixel of the 512x512 demo camera: the ray set of tests/golden/oracle_frames_golden.npz
GOLDEN_PIXELS = (np.arange(0, 512, 4)[:, None] * 512 + np.arange(0, 512, 4)[None]).ravel()


def load_golden_frame(frame_idx):
    """committed end-to-end oracle image of one pose (tests/golden/make_oracle_frames_golden.py) -> dict like
    oracle.render.render_test's result, plus the occupancy field it was rendered with"""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_frames_golden.npz")
    z = np.load(path)
    assert frame_idx in list(z["frames"]), (frame_idx, z["frames"])
    assert np.array_equal(z["pixel_index"], GOLDEN_PIXELS)
    ref = {k: z[f"{frame_idx}/{k}"] for k in ("rgb", "alpha", "depth", "counter")}
    ref["occ"] = np.unpackbits(z[f"{frame_idx}/occ_bits"])[:64 ** 3].reshape(64, 64, 64).astype(bool)
    return ref


def oracle_scene(frame_idx=0, track="male-3-casual", sigma_in=100.0):
    key = (frame_idx, track, sigma_in)
    if key in _CACHE:
        return _CACHE[key]
    from instantavatar_b200 import synthetic
    subj = oscene.build_subject(track=track)
    pose = synthetic.load_pose(frame_idx, track)
    fr = subj.prepare_frame(pose)
    net = oscene.build_net(subj, sigma_in=sigma_in)
    field, density, jit = oscene.build_occupancy(subj, fr, net)
    sc = {"subj": subj, "pose": pose, "frame": fr, "net": net, "occ": field, "occ_density": density, "occ_jitter": jit}
    _CACHE[key] = sc
    return sc


def upload(sc, device="cuda"):
    """oracle scene -> instantavatar_b200.ops.Scene (same inputs on the device)."""
    import torch
    from instantavatar_b200 import ops
    subj, fr, net = sc["subj"], sc["frame"], sc["net"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    offset_k, scale_k, tfs = t(subj.offset_kernel), t(subj.scale_kernel), t(fr["tfs"])
    fld, vd, aabb = ops.precompute(t(subj.lbs_voxel), tfs, offset_k, scale_k)
    table_h, mlp_h = ops.params_to_half(t(net.enc), t(net.col))
    occ_bits = ops.pack_occupancy(t(sc["occ"]))
    scene = ops.Scene(field=fld, offset_k=offset_k, scale_k=scale_k, tfs=tfs, table_h=table_h, mlp_h=mlp_h,
                      net_center=t(net.center), net_scale=t(net.scale), occ_bits=occ_bits,
                      occ_aabb=t(fr["bbox_deformed"].reshape(6)))
    return scene, {"voxel_d": vd, "aabb": aabb}


def oracle_model(sc, eval_mode=True):
    return lambda p: orender.deform_query(p, sc["frame"], sc["subj"], sc["net"], eval_mode)


def oracle_model_aux(sc, eval_mode=False):
    """model callable for oracle.render.render_train(return_aux=True)"""
    def f(p, want_aux=False):
        return orender.deform_query(p, sc["frame"], sc["subj"], sc["net"], eval_mode, return_aux=want_aux)
    return f
