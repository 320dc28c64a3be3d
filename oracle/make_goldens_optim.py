"""Generates tests/golden/contact_loss.npz from the UNMODIFIED reference: the body of ObjPose_Opt.contact_loss is taken
from /root/reference/optim/optimizer.py with `ast` (the module itself cannot be imported here: tensorboard / pytorch3d /
matplotlib are absent) and executed as is, in float32 and float64, with autograd for the gradient.
Run in the build container only (needs /root/reference):  python -m oracle.make_goldens_optim"""
import ast
import textwrap
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/optim/optimizer.py")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "contact_loss.npz"


def reference_contact_loss():
    src = REF.read_text()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "ObjPose_Opt")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "contact_loss")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    ns = {"torch": torch}
    exec(code, ns)
    return ns["contact_loss"]


def inputs(seed, n_obj, n_hum):
    g = np.random.default_rng(seed)
    obj = g.normal(size=(n_obj, 3)) * 0.3 + np.array([0.2, 0.0, 0.5])
    hum = g.normal(size=(n_hum, 3)) * np.array([0.25, 0.6, 0.15])
    p = np.clip(g.beta(0.4, 1.5, size=n_obj), 0, 1)
    q = np.clip(g.beta(0.3, 2.0, size=n_hum), 0, 1)
    return obj.astype(np.float32), hum.astype(np.float32), p.astype(np.float32), q.astype(np.float32)


CASES = {"small": (0, 257, 300), "smplx": (1, 3000, 10475), "coincident": (2, 64, 64)}


def reference_icp():
    """Imports the reference's optim/icp/icp.py UNMODIFIED.  Its only pytorch3d dependencies are `knn_points` (stubbed by a
    brute-force K = 1 search with pytorch3d's return convention) and `pytorch3d.structures.utils` (only touched for list
    weights, unused here); optim/icp/utils.py vendors the rest."""
    import collections
    import sys
    import types

    KNN = collections.namedtuple("KNN", "dists idx knn")

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False, **kw):
        assert K == 1
        d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
        dist, idx = d.min(-1)
        nn = torch.gather(p2, 1, idx[..., None].expand(-1, -1, p2.shape[-1]))
        return KNN(dist[..., None], idx[..., None], nn[:, :, None, :] if return_nn else None)

    mods = {n: types.ModuleType(n) for n in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.structures",
                                             "pytorch3d.structures.utils")}
    mods["pytorch3d.ops"].knn_points = knn_points
    mods["pytorch3d.ops.knn"].knn_points = knn_points
    mods["pytorch3d.structures"].utils = mods["pytorch3d.structures.utils"]
    saved = {n: sys.modules.get(n) for n in list(mods) + ["optim", "optim.icp", "optim.icp.icp", "optim.icp.utils"]}
    sys.modules.update(mods)
    sys.path.insert(0, "/root/reference")
    try:
        for n in ("optim", "optim.icp", "optim.icp.icp", "optim.icp.utils"):
            sys.modules.pop(n, None)
        import importlib

        icp = importlib.import_module("optim.icp.icp")
    finally:
        sys.path.remove("/root/reference")
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return icp


def icp_inputs(seed, n_obj, n_hum):
    """Two contact patches that roughly face each other, with unit normals, and an initial similarity transform."""
    g = np.random.default_rng(seed)
    hum = np.stack([g.uniform(-0.3, 0.3, n_hum), g.uniform(-0.3, 0.3, n_hum), np.zeros(n_hum)], -1)
    hum[:, 2] = 0.15 * np.sin(4 * hum[:, 0]) * np.cos(3 * hum[:, 1])
    hn = np.stack([-0.6 * np.cos(4 * hum[:, 0]) * np.cos(3 * hum[:, 1]), 0.45 * np.sin(4 * hum[:, 0]) * np.sin(3 * hum[:, 1]), np.ones(n_hum)], -1)
    hn /= np.linalg.norm(hn, axis=1, keepdims=True)
    sel = g.choice(n_hum, n_obj, replace=n_obj > n_hum)
    ang = 0.35
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    obj = (hum[sel] + g.normal(size=(n_obj, 3)) * 0.01) @ Rz.T * 1.15 + np.array([0.05, -0.08, 0.12])
    on = (-hn[sel]) @ Rz.T + g.normal(size=(n_obj, 3)) * 0.05
    on /= np.linalg.norm(on, axis=1, keepdims=True)
    R0 = np.eye(3)[None]
    T0 = np.array([[0.01, 0.02, -0.03]])
    s0 = np.array([1.0])
    f = lambda a: a.astype(np.float32)
    return f(obj)[None], f(on)[None], f(hum)[None], f(hn)[None], f(R0), f(T0), f(s0)


def reference_utils():
    """rot6d_to_matrix, matrix_to_rot6d, apply_transformation, calculate_centroid out of optim/utils.py (the module imports
    trimesh / matplotlib, absent here) and mask_loss_iou out of optim/optimizer.py, executed as they are."""
    ns = {"torch": torch, "F": torch.nn.functional}
    src = (REF.parent / "utils.py").read_text()
    for n in ast.parse(src).body:
        if isinstance(n, ast.FunctionDef) and n.name in ("rot6d_to_matrix", "matrix_to_rot6d", "apply_transformation", "calculate_centroid"):
            exec(textwrap.dedent(ast.get_source_segment(src, n)), ns)
    src = REF.read_text()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "ObjPose_Opt")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "mask_loss_iou")
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns


def utils_inputs():
    g = torch.Generator().manual_seed(11)
    return dict(rot6d=torch.randn(6, generator=g), verts=torch.randn(50, 3, generator=g), trans=torch.randn(3, generator=g),
                scale=torch.tensor(1.3), mask=(torch.rand(20, 30, generator=g) > 0.6).float() * torch.rand(20, 30, generator=g),
                target=(torch.rand(20, 30, generator=g) > 0.5).float(), mat=torch.linalg.qr(torch.randn(3, 3, generator=g))[0])


ICP_CASES = {"patch": (3, 400, 700, False), "patch_scale": (4, 900, 600, True)}


def main():
    icp = reference_icp()
    out_icp = {}
    for name, (seed, n_obj, n_hum, est_scale) in ICP_CASES.items():
        obj, on, hum, hn, R0, T0, s0 = icp_inputs(seed, n_obj, n_hum)
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            t = lambda a: torch.from_numpy(a).to(dt)
            sol = icp.ICP(t(obj), t(hum), init_transform=icp.SimilarityTransform(t(R0), t(T0), t(s0)), max_iterations=30,
                          estimate_scale=est_scale, obj_contact_normals=t(on), hum_contact_normals=t(hn))
            out_icp[f"{name}_{tag}_R"] = sol.RTs.R.numpy()
            out_icp[f"{name}_{tag}_T"] = sol.RTs.T.numpy()
            out_icp[f"{name}_{tag}_s"] = sol.RTs.s.numpy()
            out_icp[f"{name}_{tag}_Xt"] = sol.Xt.numpy()
            out_icp[f"{name}_{tag}_rmse"] = sol.rmse.numpy()
            out_icp[f"{name}_{tag}_iters"] = np.array(len(sol.t_history))
            out_icp[f"{name}_{tag}_converged"] = np.array(bool(sol.converged))
    ns, u = reference_utils(), utils_inputs()

    class Self:
        target_mask = u["target"]
    out_icp["utils_rot"] = ns["rot6d_to_matrix"](u["rot6d"]).numpy()
    out_icp["utils_rot6d"] = ns["matrix_to_rot6d"](u["mat"]).numpy()
    out_icp["utils_transformed"] = ns["apply_transformation"](u["verts"], u["rot6d"], u["trans"], u["scale"]).numpy()
    out_icp["utils_centroid"] = ns["calculate_centroid"](u["mask"]).numpy()
    out_icp["utils_centroid_empty"] = ns["calculate_centroid"](torch.zeros(6, 8)).numpy()
    out_icp["utils_mask_loss"] = ns["mask_loss_iou"](Self, u["mask"]).numpy()
    np.savez_compressed(OUT.with_name("icp.npz"), **out_icp)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out_icp.items() if k.endswith(("_s", "_rmse", "_iters", "_T"))})
    fn = reference_contact_loss()
    out = {}
    for name, (seed, n_obj, n_hum) in CASES.items():
        obj, hum, p, q = inputs(seed, n_obj, n_hum)
        if name == "coincident":
            hum[:16] = obj[:16]  # zero distances: cdist's subgradient is 0 there
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            class Self:
                object_contact_probs = torch.from_numpy(p).to(dt)
                human_contact_probs = torch.from_numpy(q).to(dt)
            o = torch.from_numpy(obj).to(dt).requires_grad_(True)
            loss = fn(Self, o, torch.from_numpy(hum).to(dt))
            loss.backward()
            out[f"{name}_{tag}_loss"] = loss.detach().numpy()
            out[f"{name}_{tag}_grad"] = o.grad.numpy()
    np.savez_compressed(OUT, **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
