#!/usr/bin/env python
"""Pose refinement of an object against a human (reference: `python -m optim.fit`, optim/fit.py) on the B200 path with a
synthetic scene: contact ICP for the initial transform, then the Adam loop over rotation / translation with the soft
silhouette (mask + centroid terms) and the contact term -- the kernels behind ivlm_knn1, ivlm_soft_silhouette(+_backward) and
ivlm_contact_loss.  Needs a B200.

    python examples/demo_synthetic_fit.py [--iters 120] [--size 256]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from interactvlm_b200 import optim as PO  # noqa: E402
from interactvlm_b200 import synthetic as S  # noqa: E402

LOSS_WEIGHTS = {"mask_loss": {"w": 5.0, "kick_in": 0}, "centroid_loss": {"w": 1.0e-4, "kick_in": 0},   # optim/cfg/fit.yaml
                "contact_loss": {"w": 10.0, "kick_in": 0}}


def make_scene(size=128, seed=0, device="cuda"):
    """Human = a large blob, object = a small blob resting against it at the ground-truth pose; contacts = the vertices of
    each within 4 cm of the other; target mask = the object's silhouette at the ground-truth pose."""
    hv, hf = S.make_test_mesh("blob", n_lat=24, n_lon=40, seed=seed)
    ov, of_ = S.make_test_mesh("blob", n_lat=14, n_lon=24, seed=seed + 1)
    hv = hv * 1.6 + np.array([0.0, 0.0, 3.0], np.float32)
    gt_t = np.array([0.62, 0.1, 2.75], np.float32)
    ov = ov * 0.5
    og = ov + gt_t
    d = np.linalg.norm(og[:, None, :] - hv[None, :, :], axis=-1)
    o_contact = (d.min(1) < 0.06).astype(np.float32)
    h_contact = (d.min(0) < 0.06).astype(np.float32)
    t = lambda a, dt=torch.float32: torch.as_tensor(a, dtype=dt, device=device)
    cam = {"focal_length": torch.tensor([1.1 * size, 1.1 * size]), "principal_point": torch.tensor([size / 2.0, size / 2.0])}
    ren = PO.SSRenderer((size, size), None, t(of_, torch.int64), cam, device=device)
    with torch.no_grad():
        mask = (ren.render(t(og))[0][0, ..., 3] > 0.5).float()
    human = {"vertices": t(hv), "contact_verts": t(h_contact), "centroid_offset": torch.zeros(3, device=device)}
    obj = {"vertices": t(ov), "contact_verts": t(o_contact), "mask": mask}
    return human, obj, ren, t(gt_t), (hf, of_)


def run(iters=120, size=128, use_icp=True, device="cuda"):
    human, obj, ren, gt_t, _ = make_scene(size, device=device)
    rot0 = PO.matrix_to_rot6d(torch.eye(3, device=device)[None])[0] + torch.tensor([0.0, 0.08, -0.05, 0.06, 0.0, 0.04], device=device)
    t0 = gt_t + torch.tensor([0.18, -0.12, 0.10], device=device)
    info = {}
    if use_icp and int(obj["contact_verts"].sum()) >= 4 and int(human["contact_verts"].sum()) >= 4:
        oc, hc = obj["contact_verts"] > 0.3, human["contact_verts"] > 0.5          # fit.py:109-114 thresholds
        sol = PO.ICP(obj["vertices"][oc][None], human["vertices"][hc][None],
                     init_transform=PO.SimilarityTransform(PO.rot6d_to_matrix(rot0[None]), t0[None], torch.ones(1, device=device)),
                     max_iterations=10, estimate_scale=False)
        rot0, t0 = PO.matrix_to_rot6d(sol.RTs.R)[0], sol.RTs.T[0]                    # fit.py:192-193
        info["icp_rmse"] = float(sol.rmse)
    model = PO.ObjPose_Opt(rot0, t0, torch.tensor(1.0), human, obj, ren, vars=("pose",)).to(device)
    info["t_err_start"] = float((model.translation.detach() - gt_t).norm())
    hist = PO.fit(model, LOSS_WEIGHTS, max_iter=iters)
    info["t_err_end"] = float((model.translation.detach() - gt_t).norm())
    return model, hist, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=120)
    ap.add_argument("--size", type=int, default=256)
    args = ap.parse_args()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model, hist, info = run(args.iters, args.size)
    e1.record()
    torch.cuda.synchronize()
    print(f"{len(hist)} iterations at {args.size}^2 in {e0.elapsed_time(e1):.0f} ms ({e0.elapsed_time(e1) / len(hist):.2f} ms/iter incl. set-up)")
    print("first:", {k: round(v, 4) for k, v in hist[0].items()})
    print("last: ", {k: round(v, 4) for k, v in hist[-1].items()})
    print(info)


if __name__ == "__main__":
    main()
