"""Deterministic synthetic inputs (data, not algorithms).

The licensed SMPL pickles, the PeopleSnapshot images and any trained
checkpoint are absent, so benchmarks and tests run on:

* a synthetic SMPL-compatible body model ("capsule-man") with exactly the keys
  the reference loader reads (deformers/smplx/body_models.py:127-262),
* real SMPL pose frames copied from the reference's pose tracks into
  tests/golden/poses.npz,
* an analytic "trained-like" NGP parameter set whose density is the capsule-man
  occupancy, so that a render has a realistic sample/termination profile,
* the demo pinhole camera of novel_view.py:27-44 scaled to 512x512.

Both the oracle (tests) and the product (bench, smoke) consume these arrays;
nothing here is on the hot path.
"""
from __future__ import annotations

import os

import numpy as np

SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int64
)

# rest-pose joint locations of a ~1.7 m T-posed body, metres, y up, x = subject's left
_REST_JOINTS = np.array(
    [
        [0.00, -0.24, 0.03], [0.07, -0.33, 0.02], [-0.07, -0.33, 0.02], [0.00, -0.12, 0.00],
        [0.10, -0.71, 0.02], [-0.10, -0.71, 0.02], [0.00, 0.02, 0.02], [0.09, -1.11, -0.02],
        [-0.09, -1.11, -0.02], [0.00, 0.07, 0.04], [0.12, -1.17, 0.10], [-0.12, -1.17, 0.10],
        [0.00, 0.28, 0.00], [0.08, 0.19, 0.00], [-0.08, 0.19, 0.00], [0.00, 0.37, 0.03],
        [0.18, 0.23, -0.01], [-0.18, 0.23, -0.01], [0.44, 0.22, -0.03], [-0.44, 0.22, -0.03],
        [0.69, 0.23, -0.03], [-0.69, 0.23, -0.03], [0.78, 0.22, -0.04], [-0.78, 0.22, -0.04],
    ],
    dtype=np.float64,
)

# capsule radius around each joint's bone (joint -> first child, or a blob at leaves)
_RADII = np.array(
    [0.13, 0.085, 0.085, 0.13, 0.06, 0.06, 0.13, 0.045, 0.045, 0.13, 0.04, 0.04, 0.06, 0.07, 0.07,
     0.10, 0.055, 0.055, 0.04, 0.04, 0.035, 0.035, 0.04, 0.04],
    dtype=np.float64,
)

N_VERTS = 6890
N_FACES = 13776


def _bone_segments(joints: np.ndarray):
    """(a, b) end points of the capsule attached to each joint."""
    a = joints.copy()
    b = joints.copy()
    for j in range(24):
        children = np.nonzero(SMPL_PARENTS == j)[0]
        if len(children):
            b[j] = joints[children[0]]
        else:  # leaf: short stub continuing the parent's direction
            d = joints[j] - joints[SMPL_PARENTS[j]]
            b[j] = joints[j] + 0.6 * d
    return a, b


def _seg_dist(p: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """distance from points p [N,3] to segments a[J,3]-b[J,3] -> [N,J]"""
    ab = b - a
    ap = p[:, None, :] - a[None]
    t = (ap * ab[None]).sum(-1) / np.maximum((ab * ab).sum(-1), 1e-12)[None]
    t = np.clip(t, 0.0, 1.0)
    c = a[None] + t[..., None] * ab[None]
    return np.linalg.norm(p[:, None, :] - c, axis=-1)


def make_smpl_dict(seed: int = 0) -> dict:
    """Synthetic SMPL-compatible model dict."""
    rng = np.random.default_rng(seed)
    a, b = _bone_segments(_REST_JOINTS)
    per = N_VERTS // 24
    counts = np.full(24, per)
    counts[: N_VERTS - per * 24] += 1
    verts = []
    for j in range(24):
        n = counts[j]
        t = rng.uniform(0.0, 1.0, n)
        axis = b[j] - a[j]
        L = np.linalg.norm(axis)
        axis = axis / max(L, 1e-9)
        # orthonormal frame
        ref = np.array([0.0, 0.0, 1.0]) if abs(axis[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
        u = np.cross(axis, ref); u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        ang = rng.uniform(0, 2 * np.pi, n)
        # extend a little beyond the end points so capsule caps are covered
        s = (t * 1.2 - 0.1) * L
        r = _RADII[j] * np.sqrt(np.clip(1.0 - np.clip(np.abs(t * 1.2 - 0.1 - 0.5) - 0.5, 0, None) ** 2 * 25, 0.05, 1))
        verts.append(a[j] + s[:, None] * axis + (r * np.cos(ang))[:, None] * u + (r * np.sin(ang))[:, None] * v)
    v_template = np.concatenate(verts, 0)
    assert v_template.shape == (N_VERTS, 3)

    d = _seg_dist(v_template, a, b)  # [V,24]
    w = np.exp(-((d / 0.06) ** 2))
    # keep the 4 strongest bones per vertex
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0)
    w = w / w.sum(1, keepdims=True)

    # joint regressor: normalised mean of the 64 vertices closest to each rest joint
    J_regressor = np.zeros((24, N_VERTS))
    for j in range(24):
        dj = np.linalg.norm(v_template - _REST_JOINTS[j], axis=1)
        idx = np.argsort(dj)[:64]
        J_regressor[j, idx] = 1.0 / 64

    shapedirs = rng.normal(0.0, 1e-3, (N_VERTS, 3, 10))
    posedirs = rng.normal(0.0, 1e-4, (N_VERTS, 3, 207))
    faces = rng.integers(0, N_VERTS, (N_FACES, 3)).astype(np.uint32)
    kintree = np.stack([SMPL_PARENTS.copy(), np.arange(24)]).astype(np.int64)
    kintree[0, 0] = 2 ** 32 - 1
    return {
        "v_template": v_template,
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": J_regressor,
        "kintree_table": kintree,
        "weights": w,
        "f": faces,
    }


_SMPL_CACHE = {}


def smpl_dict_cached(seed: int = 0) -> dict:
    if seed not in _SMPL_CACHE:
        _SMPL_CACHE[seed] = make_smpl_dict(seed)
    return _SMPL_CACHE[seed]


def load_pose(frame: int = 0, track: str = "male-3-casual") -> dict:
    """One real SMPL pose frame (copied from the reference's
    data/PeopleSnapshot/<track>/poses/anim_nerf_train.npz into tests/golden/poses.npz)."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "poses.npz")
    z = np.load(path)
    frames = list(z[f"{track}/frames"])
    i = frames.index(frame)
    return {
        "betas": z[f"{track}/betas"].astype(np.float32).reshape(1, 10),
        "global_orient": z[f"{track}/global_orient"][i : i + 1].astype(np.float32),
        "body_pose": z[f"{track}/body_pose"][i : i + 1].astype(np.float32),
        "transl": z[f"{track}/transl"][i : i + 1].astype(np.float32),
    }


def track_frames(track: str = "male-3-casual") -> list:
    """frame numbers of `track` held by tests/golden/poses.npz"""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "poses.npz")
    return [int(f) for f in np.load(path)[f"{track}/frames"]]


def demo_camera_rays(H: int = 512, W: int = 512):
    """Rays of the reference demo camera (novel_view.py:27-44: f=2000 px at 1080^2, c2w = I)
    rescaled to HxW, generated as datasets/peoplesnapshot.py:12-25 does."""
    f = 2000.0 * H / 1080.0
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], dtype=np.float64)
    x, y = np.meshgrid(np.arange(W), np.arange(H), indexing="xy")
    xy = np.stack([x, y, np.ones_like(x)], axis=-1).reshape(-1, 3).astype(np.float32)
    d_c = xy @ np.linalg.inv(K).T
    d_w = d_c / np.linalg.norm(d_c, axis=1, keepdims=True)
    o_w = np.zeros_like(d_w)
    return o_w.astype(np.float32), d_w.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# hash-grid layout (host copy; the device code and the oracle each compute their own)
# ----------------------------------------------------------------------------------------------
def hashgrid_layout():
    res, scale, size, off = [], [], [], []
    o = 0
    for l in range(16):
        s = np.float32(np.exp2(np.float32(l) * np.log2(np.float32(1.5)))) * np.float32(16.0) - np.float32(1.0)
        r = int(np.ceil(s)) + 1
        n = min((r ** 3 + 7) // 8 * 8, 1 << 19)
        res.append(r); scale.append(float(s)); size.append(n); off.append(o)
        o += n
    return {"res": res, "scale": scale, "size": size, "offset": off, "total": o}


N_ENC_MLP = 64 * 32 + 16 * 64  # 3072
N_COL_MLP = 64 * 16 + 64 * 64 + 16 * 64  # 6144


def analytic_avatar_params(cano_joints: np.ndarray, center: np.ndarray, scale: np.ndarray, seed: int = 1337,
                           sigma_in: float = 100.0):
    """Flat fp32 `encoder.params` / `color_net.params` (tcnn ordering) whose density
    is ~ +sigma_in inside the canonical capsule-man and ~ -sigma_in outside.

    cano_joints [24,3]: joints of the canonical (A-pose) body; center/scale: NeRFNGPNet bbox
    normalisation (ngp.py:64-77) so that grid vertex v of a level maps to
    x = (v/scale_l - 0.5) * scale + center.
    """
    rng = np.random.default_rng(seed)
    lay = hashgrid_layout()
    grid = rng.uniform(-1e-4, 1e-4, (lay["total"], 2)).astype(np.float32)
    a, b = _bone_segments(cano_joints.astype(np.float64))

    def sdf(p):
        out = np.empty(len(p))
        for s in range(0, len(p), 65536):
            d = _seg_dist(p[s : s + 65536], a, b) - _RADII[None]
            out[s : s + 65536] = d.min(1)
        return out

    # dense levels 0..3 carry the occupancy in feature 0 and a smooth colour code in feature 1
    for l in range(4):
        r = lay["res"][l]
        n = lay["size"][l]
        idx = np.arange(r ** 3)
        vx, vy, vz = idx % r, (idx // r) % r, idx // (r * r)
        v = np.stack([vx, vy, vz], -1).astype(np.float64)
        # inverse of pos = x*scale + 0.5  ->  x = (v - 0.5) / scale
        x01 = (v - 0.5) / lay["scale"][l]
        p = (x01 - 0.5) * scale[None] + center[None]
        f0 = np.clip(-sdf(p) / 0.02, -1.0, 1.0) * 0.25
        f1 = 0.25 * np.sin(7.0 * p[:, 0] + 3.0 * l) * np.cos(5.0 * p[:, 1]) + 0.1 * np.sin(9.0 * p[:, 2])
        dst = idx % n
        grid[lay["offset"][l] + dst, 0] = f0
        grid[lay["offset"][l] + dst, 1] = f1
    # hashed levels: small high-frequency detail
    for l in range(4, 16):
        sl = slice(lay["offset"][l], lay["offset"][l] + lay["size"][l])
        grid[sl, 0] = rng.normal(0, 2e-3, lay["size"][l])
        grid[sl, 1] = rng.normal(0, 2e-2, lay["size"][l])

    W1 = np.zeros((64, 32), np.float32)
    for k in range(32):
        W1[k, k] = 1.0
        W1[k + 32, k] = -1.0
    W2 = np.zeros((16, 64), np.float32)
    for l in range(16):
        W2[0, 2 * l] = sigma_in
        W2[0, 2 * l + 32] = -sigma_in
    for o in range(1, 16):
        col = rng.normal(0, 1.0, 16).astype(np.float32) * 4.0
        for l in range(16):
            W2[o, 2 * l + 1] = col[l]
            W2[o, 2 * l + 1 + 32] = -col[l]
    lim = lambda fi, fo: np.sqrt(6.0 / (fi + fo))
    W3 = rng.uniform(-lim(16, 64), lim(16, 64), (64, 16)).astype(np.float32) * 2
    W4 = rng.uniform(-lim(64, 64), lim(64, 64), (64, 64)).astype(np.float32) * 2
    W5 = rng.uniform(-lim(64, 16), lim(64, 16), (16, 64)).astype(np.float32) * 2
    enc = np.concatenate([W1.ravel(), W2.ravel(), grid.ravel()]).astype(np.float32)
    col = np.concatenate([W3.ravel(), W4.ravel(), W5.ravel()]).astype(np.float32)
    assert enc.size == N_ENC_MLP + lay["total"] * 2 and col.size == N_COL_MLP
    return enc, col


def random_params(seed: int = 1337, table_scale: float = 1e-4):
    """tcnn-style random init (uniform tables, Xavier-uniform weights)."""
    rng = np.random.default_rng(seed)
    lay = hashgrid_layout()
    grid = rng.uniform(-table_scale, table_scale, lay["total"] * 2).astype(np.float32)
    lim = lambda fi, fo: np.sqrt(6.0 / (fi + fo))
    mk = lambda o, i: rng.uniform(-lim(i, o), lim(i, o), (o, i)).astype(np.float32).ravel()
    enc = np.concatenate([mk(64, 32), mk(16, 64), grid])
    col = np.concatenate([mk(64, 16), mk(64, 64), mk(16, 64)])
    return enc.astype(np.float32), col.astype(np.float32)
