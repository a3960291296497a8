"""CPU: pins the oracle's Python-level restatements against golden vectors produced by executing the
reference's own Python (tests/golden/make_ref_python_golden.py)."""
import os

import numpy as np
import pytest

from instantavatar_b200 import synthetic
from oracle import frame as oframe
from oracle import render as orender
from oracle.smpl_np import SMPLNumpy


@pytest.fixture(scope="module")
def smpl_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "smpl_golden.npz"))


@pytest.fixture(scope="module")
def py_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "pyfuncs_golden.npz"))


@pytest.mark.parametrize("track,frame", [("male-3-casual", 0), ("male-3-casual", 20), ("male-3-casual", 57),
                                         ("male-3-casual", 100), ("female-4-casual", 0), ("female-4-casual", 40)])
def test_smpl_forward_matches_reference(smpl_golden, track, frame):
    smpl = SMPLNumpy(synthetic.smpl_dict_cached(0))
    p = synthetic.load_pose(frame, track)
    out = smpl.forward(p["betas"], p["body_pose"], p["global_orient"], p["transl"])
    key = f"{track}/{frame}"
    np.testing.assert_allclose(out["A"], smpl_golden[key + "/A"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(out["vertices"], smpl_golden[key + "/vertices"], atol=5e-6, rtol=1e-5)
    np.testing.assert_allclose(out["joints"], smpl_golden[key + "/joints24"], atol=5e-6, rtol=1e-5)


def test_canonical_pose_and_bbox(smpl_golden):
    smpl = SMPLNumpy(synthetic.smpl_dict_cached(0))
    bp = oframe.rest_pose_a()
    np.testing.assert_array_equal(bp, smpl_golden["cano/body_pose"])
    out = smpl.forward(synthetic.load_pose(0)["betas"], bp)
    np.testing.assert_allclose(out["A"], smpl_golden["cano/A"], atol=2e-6)
    np.testing.assert_allclose(out["vertices"], smpl_golden["cano/vertices"], atol=5e-6)
    np.testing.assert_allclose(oframe.bbox_from_smpl(smpl_golden["cano/vertices"]), smpl_golden["cano/bbox"], atol=1e-6)


def test_composite_train_matches_reference(py_golden):
    w, tr = orender.composite_train(py_golden["composite/sigma"], py_golden["composite/dists"])
    np.testing.assert_allclose(w, py_golden["composite/w"], atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(tr, py_golden["composite/trans"], atol=1e-6, rtol=1e-5)


def test_connected_component_matches_reference(py_golden):
    dens = py_golden["grid/density"]
    field = orender._field_from_density(dens)
    np.testing.assert_array_equal(field, py_golden["grid/field"])
    f0 = (1 - np.exp(np.float32(0.01) * -dens)).astype(np.float32)
    f0 = orender.max_pool3(f0)
    f0 = f0 > min(f0.mean(dtype=np.float32), np.float32(0.01))
    np.testing.assert_array_equal(orender.max_connected_component(f0), py_golden["grid/mcc"])


def test_nerf_loss_matches_reference(py_golden):
    pred = {"rgb": py_golden["loss/rgb_coarse"], "alpha": py_golden["loss/alpha_coarse"], "weights": py_golden["loss/weight_coarse"]}
    L = orender.nerf_loss(pred, py_golden["loss/tgt_rgb"], py_golden["loss/tgt_alpha"])
    for k in ["mse_loss", "loss_alpha_coarse", "reg_alpha", "reg_density", "loss"]:
        assert abs(L[k] - float(py_golden["loss/out_" + k])) < 2e-6, k
