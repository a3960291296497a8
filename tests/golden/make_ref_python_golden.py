"""Generates golden vectors by IMPORTING / EXECUTING the reference's own Python (CPU) in the build
container (needs /root/reference; it does not exist on the GPU box, hence committed fixtures).

 * smpl_golden.npz   : reference deformers/smplx SMPL.forward on the synthetic body model
                       (A, vertices, joints) for several real pose frames + the A-pose.
 * pyfuncs_golden.npz: pure-torch reference functions executed from their source via ast
                       (module imports of hydra/kaolin/tinycudann would fail):
                       raymarcher_acc.composite, density_grid.max_connected_component,
                       snarf_deformer.get_bbox_from_smpl / get_predefined_rest_pose,
                       utils/loss.NeRFLoss.forward, deformer_torch.query_weights_smpl smoothing loop.
Only numeric inputs/outputs are stored, no reference source.
"""
import ast, importlib.util, os, sys, types
import numpy as np, torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from instantavatar_b200 import synthetic


def import_ref_smplx():
    pkg_dir = f"{REF}/instant_avatar/deformers/smplx"
    spec = importlib.util.spec_from_file_location("ref_smplx", f"{pkg_dir}/__init__.py", submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_smplx"] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_functions(path, names, extra_globals=None):
    """exec only the named top-level functions/classes of a reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    g = {"torch": torch, "F": torch.nn.functional, "nn": torch.nn, "np": np}
    g.update(extra_globals or {})
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), g)
    return g


def main():
    torch.manual_seed(0)
    smplx = import_ref_smplx()
    data = synthetic.smpl_dict_cached(0)
    Struct = type("Struct", (), {})
    st = Struct()
    for k, v in data.items():
        setattr(st, k, v)
    model = smplx.SMPL("unused", data_struct=st, gender="male")
    out = {}
    poses = [("male-3-casual", f) for f in (0, 20, 57, 100)] + [("female-4-casual", f) for f in (0, 40)]
    for track, f in poses:
        p = synthetic.load_pose(f, track)
        t = {k: torch.from_numpy(v) for k, v in p.items()}
        o = model(betas=t["betas"], body_pose=t["body_pose"], global_orient=t["global_orient"], transl=t["transl"])
        key = f"{track}/{f}"
        out[key + "/A"] = o.A[0].detach().numpy(); out[key + "/vertices"] = o.vertices[0].detach().numpy()
        out[key + "/joints24"] = o.joints[0, :24].detach().numpy()
    # canonical A-pose, as SNARFDeformer.initialize does (snarf_deformer.py:41-52)
    sd = ref_functions(f"{REF}/instant_avatar/deformers/snarf_deformer.py", {"get_predefined_rest_pose", "get_bbox_from_smpl"})
    bp = sd["get_predefined_rest_pose"]("A_pose", device="cpu")
    betas = torch.from_numpy(synthetic.load_pose(0)["betas"])
    o = model(betas=betas, body_pose=bp)
    out["cano/body_pose"] = bp.numpy(); out["cano/A"] = o.A[0].detach().numpy()
    out["cano/vertices"] = o.vertices[0].detach().numpy()
    out["cano/bbox"] = sd["get_bbox_from_smpl"](o.vertices.detach()).numpy()
    np.savez_compressed(f"{HERE}/smpl_golden.npz", **{k: v.astype(np.float32) for k, v in out.items()})
    print("smpl_golden:", len(out), "arrays")

    py = {}
    rng = np.random.default_rng(7)
    # composite (raymarcher_acc.py:25-36)
    rm = ref_functions(f"{REF}/instant_avatar/renderers/raymarcher_acc.py", {"composite"})
    sig = (rng.normal(0, 40, (64, 256))).astype(np.float32); sig[rng.random((64, 256)) < 0.6] = -1e3
    dists = np.full((64, 256), 2 / 256, np.float32)
    w, tr = rm["composite"](torch.from_numpy(sig), torch.from_numpy(dists))
    py["composite/sigma"] = sig; py["composite/dists"] = dists; py["composite/w"] = w.numpy(); py["composite/trans"] = tr.numpy()
    # max_connected_component + the field post-processing (density_grid.py:104-125)
    dg = ref_functions(f"{REF}/instant_avatar/models/structures/density_grid.py", {"max_connected_component"})
    G = 32
    dens = np.zeros((G, G, G), np.float32)
    dens[4:14, 5:20, 6:12] = rng.random((10, 15, 6)) * 50; dens[20:26, 20:28, 20:25] = rng.random((6, 8, 5)) * 80
    dens[rng.random((G, G, G)) < 0.002] = 30.0
    d = torch.from_numpy(dens)
    field = 1 - torch.exp(0.01 * -d)
    field = torch.nn.functional.max_pool3d(field[None, None], kernel_size=3, stride=1, padding=1)[0, 0]
    field = field > torch.clamp(field.mean(), max=0.01)
    mcc = dg["max_connected_component"](field)
    label = torch.mode(mcc[field], 0).values
    py["grid/density"] = dens; py["grid/mcc"] = mcc.numpy(); py["grid/field"] = (mcc == label).numpy()
    # NeRFLoss (utils/loss.py:53-79)
    ls = ref_functions(f"{REF}/instant_avatar/utils/loss.py", {"NeRFLoss"})
    opt = types.SimpleNamespace(w_rgb=1.0, w_alpha=0.1, w_reg=0.1)
    loss_fn = ls["NeRFLoss"](opt)
    pred = {"rgb_coarse": torch.rand(1, 128, 3), "alpha_coarse": torch.rand(1, 128), "weight_coarse": torch.rand(1, 128, 256) * 0.05}
    tgt = {"rgb": torch.rand(1, 128, 3), "alpha": (torch.rand(1, 128) > 0.5).float()}
    for k, v in pred.items(): v.requires_grad_(True)
    L = loss_fn(pred, tgt)
    L["loss"].backward()
    for k, v in pred.items(): py["loss/" + k] = v.detach().numpy(); py["loss/grad_" + k] = v.grad.numpy()
    for k, v in tgt.items(): py["loss/tgt_" + k] = v.numpy()
    for k, v in L.items(): py["loss/out_" + k] = np.float64(v.item())
    np.savez_compressed(f"{HERE}/pyfuncs_golden.npz", **py)
    print("pyfuncs_golden:", len(py), "arrays")


if __name__ == "__main__":
    main()
