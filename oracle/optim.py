"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the contact term of the reference's pose refinement,
ObjPose_Opt.contact_loss (optim/optimizer.py:80-96): weighted mean of the pairwise object-human vertex distances with the
outer product of the two contact-probability vectors as weights.  Pinned: tests/golden/contact_loss.npz holds outputs of
the reference's own function body (extracted from /root/reference/optim/optimizer.py by oracle/make_goldens_optim.py and
executed unmodified, value and autograd gradient); tests/test_optim_cpu.py checks this restatement against them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module."""
from __future__ import annotations

import numpy as np


def contact_loss(obj_verts, human_verts, obj_probs, human_probs, dtype=np.float64):
    """-> (loss, d loss / d obj_verts).  optimizer.py:82-96: dist = cdist(obj, human); w = outer(p_obj, p_human);
    loss = (dist * w).sum() / w.sum().  Direct differences (no |a|^2 + |b|^2 - 2ab expansion), `dtype` arithmetic."""
    o, h = np.asarray(obj_verts, dtype=dtype), np.asarray(human_verts, dtype=dtype)
    p, q = np.asarray(obj_probs, dtype=dtype), np.asarray(human_probs, dtype=dtype)
    loss_num = 0.0
    grad = np.zeros_like(o)
    for i0 in range(0, len(o), 1024):  # row blocks keep the pair matrix small
        diff = o[i0:i0 + 1024, None, :] - h[None, :, :]
        dist = np.sqrt((diff * diff).sum(-1))
        loss_num += float((dist * q[None, :]).sum(1) @ p[i0:i0 + 1024])
        inv = np.divide(1.0, dist, out=np.zeros_like(dist), where=dist > 0)
        grad[i0:i0 + 1024] = ((q[None, :] * inv)[..., None] * diff).sum(1) * p[i0:i0 + 1024, None]
    denom = p.sum() * q.sum()
    return loss_num / denom, grad / denom


def corresponding_points_alignment(X, Y, weights=None, estimate_scale=False, allow_reflection=False, eps=1e-9):
    """optim/icp/icp.py:274-417 (Umeyama): R, T, s with s X R + T ~ Y in the weighted least-squares sense; float64 numpy,
    single cloud ([n,3] inputs)."""
    X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
    n, dim = X.shape
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64)
    wsum = max(w.sum(), eps)
    Xmu, Ymu = (X * w[:, None]).sum(0) / wsum, (Y * w[:, None]).sum(0) / wsum
    Xc, Yc = (X - Xmu) * w[:, None], (Y - Ymu) * w[:, None]
    cov = Xc.T @ Yc / wsum
    U, S, Vt = np.linalg.svd(cov)
    E = np.eye(dim)
    if not allow_reflection:
        E[-1, -1] = np.linalg.det(U @ Vt)
    R = U @ E @ Vt
    if estimate_scale:
        s = (np.diag(E) * S).sum() / max((Xc * Xc).sum() / wsum, eps)
    else:
        s = 1.0
    T = Ymu - s * (Xmu @ R)
    return R, T, s


def icp(obj_pts, hum_pts, init_transform=None, max_iterations=100, relative_rmse_thr=1e-6, estimate_scale=False,
        allow_reflection=False, obj_normals=None, hum_normals=None, min_scale=None, scale_penalty=10.0):
    """optim/icp/icp.py:38-268 for one pair of clouds.  As in the reference, the nearest-neighbour query cloud (initially
    transformed object points ++ object normals) is built ONCE before the loop (icp.py:176-185) and never refreshed, so
    every iteration finds the same neighbours and the loop stops at its second pass; restated as is.
    -> dict(R, T, s, Xt, rmse, iters, converged)."""
    X0, Y = np.asarray(obj_pts, dtype=np.float64), np.asarray(hum_pts, dtype=np.float64)
    if init_transform is not None:
        R, T, s = (np.asarray(a, dtype=np.float64) for a in init_transform)
        Xt = float(s) * (X0 @ R) + T
    else:
        R, T, s = np.eye(3), np.zeros(3), 1.0
        Xt = X0.copy()
    q = Xt if obj_normals is None else np.concatenate([Xt, np.asarray(obj_normals, dtype=np.float64)], -1)
    t = Y if hum_normals is None else np.concatenate([Y, -np.asarray(hum_normals, dtype=np.float64)], -1)
    prev, rmse, converged, iters = None, None, False, 0
    for _ in range(max_iterations):
        iters += 1
        d2 = ((q[:, None, :] - t[None, :, :]) ** 2).sum(-1)
        nn = t[d2.argmin(1)]
        nn_pts, nn_normals = nn[:, :3], -nn[:, 3:]
        R, T, s = corresponding_points_alignment(X0, nn_pts, np.ones(len(X0)), estimate_scale, allow_reflection)
        Xt = s * (X0 @ R) + T
        rmse = np.sqrt(((Xt - nn_pts) ** 2).sum(1).mean())
        combined = rmse
        if nn_normals.shape[1]:
            nt = s * (nn_normals @ R)
            combined = rmse + (1 - (nt * nn_normals).sum(1))     # per-point vector, like the reference (icp.py:219-224)
        if min_scale is not None:
            combined = combined + scale_penalty * max(s - min_scale, 0)
        rel = np.ones_like(combined) if prev is None else (combined - prev) / prev
        if np.all(rel <= relative_rmse_thr):
            converged = True
            break
        prev = combined
    return dict(R=R, T=T, s=s, Xt=Xt, rmse=rmse, iters=iters, converged=converged)
