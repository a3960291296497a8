"""GPU: end-to-end comparison with the REFERENCE'S OWN KERNELS.  oracle/_ref/*.so are the reference's raymarcher,
fuse_broyden, filter and precompute extensions built from /root/reference for sm_100; oracle/ref_structure.py drives
them with the reference's host loop (only tiny-cuda-nn is replaced, by ia_ngp_forward).  The fused kernel must
reproduce that pipeline's image."""
import numpy as np
import pytest

from oracle import ref_structure
from oracle import scene as oscene
from oracle import testing as scene_util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_structure.available(), reason="oracle/_ref not built")]


def test_fused_render_matches_reference_kernels_pipeline():
    import torch
    from instantavatar_b200 import ops
    sc = scene_util.oracle_scene(0)
    scene, extra = scene_util.upload(sc)
    subj, fr = sc["subj"], sc["frame"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rs = ref_structure.RefStructure(t(subj.lbs_voxel)[None], t(subj.offset_kernel), t(subj.scale_kernel),
                                    lambda x: ops.ngp_forward(scene, x))
    rs.precompute(t(fr["tfs"])[None])
    # the reference's precompute kernel vs ours (voxel-major, padded)
    vJ_ref = rs.voxel_J[0].permute(1, 2, 3, 0).contiguous()
    assert torch.equal(scene.field[..., :12], vJ_ref)
    # occupancy grid through the reference structure with the same jitter
    jit = t(sc["occ_jitter"])
    field_ref = rs.density_grid_initialize(jit)
    dens = ops.occupancy_query(scene, jit, t(fr["bbox_deformed"].reshape(6)))
    field, bits = ops.occupancy_build(dens)
    assert (field != field_ref).float().mean().item() < 5e-4
    # render a 128x192 crop with BOTH grids equal to the reference's
    o, d, near, far = oscene.camera_rays(fr, 512, 512)
    ys, xs = np.arange(160, 352), np.arange(192, 320)
    idx = (ys[:, None] * 512 + xs[None]).ravel()
    ref = rs.render_test(t(o[idx]), t(d[idx]), t(near[idx]), t(far[idx]))
    import dataclasses
    scene2 = dataclasses.replace(scene, occ_bits=ops.pack_occupancy(field_ref), occ_aabb=torch.cat(rs.aabb).contiguous())
    out = ops.render_fwd(scene2, t(o[idx]), t(d[idx]), t(near[idx]), t(far[idx]), None, 128)
    torch.cuda.synchronize()
    e_rgb = (out["rgb"] - ref["rgb"]).abs().max(-1).values
    e_a = (out["alpha"] - ref["alpha"]).abs()
    assert (ref["alpha"] > 0.5).sum().item() > 3000
    # the reference's Broyden differs from the oracle's in the last bits (nvcc fma contraction), which moves roots by
    # <= 1e-5 and, rarely, flips an fp16 rounding inside the network: allow the north_star tolerance on all but a few rays
    assert (e_rgb > 1e-3).float().mean().item() < 2e-3, ((e_rgb > 1e-3).sum().item(), e_rgb.max().item())
    assert (e_a > 1e-3).float().mean().item() < 2e-3
    assert e_rgb.max().item() < 5e-2
