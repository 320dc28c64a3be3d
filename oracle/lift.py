"""Oracle (test infrastructure): numpy restatement of the reference's 2D->3D lifting predictors.

Follows model/components.py of the reference:
  lift_human        <- HumanContact3DPredictor.forward/_process_view   (components.py:220-277)
  lift_object_mesh  <- ObjectMeshContact3DPredictor._process_view/forward_inference (components.py:392-489)
  lift_points       <- ObjectPCAfford3DPredictor.forward/_process_view  (components.py:289-347)
  convert_contacts  <- utils/utils.py:428-443
np.add.at is the sequential equivalent of torch's CPU scatter_add_ (same accumulation order), fp32 throughout.
Pinned against the reference by tests/golden/lift_*.npz (oracle/make_goldens.py).
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x: np.ndarray) -> np.ndarray:
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def _mesh_view(values, valid_px, p2v, bary, n):
    """One view: votes/counts per vertex from pixels `valid_px` (bool [H*W])."""
    verts = p2v.reshape(-1, 3)
    w = bary.reshape(-1, 3).astype(np.float32)
    ok = ((verts >= 0) & (verts < n)).all(1) & valid_px
    verts, w, vals = verts[ok], w[ok], values[ok]
    votes = np.zeros(n, np.float32)
    cnt = np.zeros(n, np.float32)
    if verts.size == 0:
        return votes, cnt, False
    for k in range(3):  # components.py:253-255 / 474-476: three scatter_add_ passes
        np.add.at(votes, verts[:, k], (w[:, k] * vals).astype(np.float32))
        np.add.at(cnt, verts[:, k], w[:, k])
    seen = cnt > 0
    votes[seen] = votes[seen] / cnt[seen]
    return votes, seen.astype(np.float32), True


def lift_human(masks: np.ndarray, p2v: np.ndarray, bary: np.ndarray, n_verts: int) -> np.ndarray:
    """masks [B,V,H,W] fp32 logits -> [B,n_verts] (components.py:220-277)."""
    B, V = masks.shape[:2]
    out = np.zeros((B, n_verts), np.float32)
    views = np.zeros((B, n_verts), np.float32)
    for b in range(B):
        for v in range(V):
            x = np.clip(masks[b, v], -20.0, 20.0)  # :241
            p = _sigmoid(x).reshape(-1)            # :242
            votes, seen, any_px = _mesh_view(p, np.ones(p.shape, bool), p2v[v], bary[v], n_verts)
            if not any_px:
                continue
            out[b] += votes
            views[b] += seen
    ok = views > 0
    out[ok] = out[ok] / views[ok]
    return np.clip(out, 0.0, 1.0)  # :231


def lift_object_mesh(masks: np.ndarray, p2v: np.ndarray, bary: np.ndarray, n_verts: int, thr: float = 0.3):
    """masks [1,V,H,W] fp32 logits -> [1,n_verts]; only pixels with sigmoid > thr vote (components.py:446-489)."""
    B, V = masks.shape[:2]
    out = np.zeros((B, n_verts), np.float32)
    views = np.zeros((B, n_verts), np.float32)
    for b in range(B):
        for v in range(V):
            p = _sigmoid(masks[b, v]).reshape(-1)
            votes, seen, any_px = _mesh_view(p, p > np.float32(thr), p2v[v], bary[v], n_verts)
            if not any_px:
                continue
            out[b] += votes
            views[b] += seen
    ok = views > 0
    out[ok] = out[ok] / views[ok]
    return out


def lift_points(values: np.ndarray, p2p: np.ndarray, n_points: int) -> np.ndarray:
    """values [B,V,H,W] fp32 (already sigmoid-ed heat maps for HM view types) -> [B,n_points]
    (components.py:289-347; numpy twin in preprocess_data/utils_obj_pc.py:47-86)."""
    B, V = values.shape[:2]
    out = np.zeros((B, n_points), np.float32)
    views = np.zeros((B, n_points), np.float32)
    for b in range(B):
        for v in range(V):
            m = p2p[v].reshape(-1)
            ok = m != -1
            votes = np.zeros(n_points, np.float32)
            cnt = np.zeros(n_points, np.float32)
            np.add.at(votes, m[ok], values[b, v].reshape(-1)[ok])
            np.add.at(cnt, m[ok], np.float32(1.0))
            seen = cnt > 0
            votes[seen] /= cnt[seen]
            out[b] += votes
            views[b] += seen
    ok = views > 0
    out[ok] /= views[ok]
    return out


def convert_contacts(contact: np.ndarray, mapping: np.ndarray) -> np.ndarray:
    """contact [B,6890] x mapping [10475,6890] -> [B,10475] (utils/utils.py:428-443: bmm(mapping, contact[...,None]))."""
    return (mapping.astype(np.float32) @ contact.astype(np.float32).T).T


def f1_metrics(pred: np.ndarray, gt: np.ndarray, thr: float = 0.5, eps: float = 1e-10):
    """get_h_contact_metrics (utils/eval_utils.py:63-94): batch-mean F1/precision/recall; pred >= thr, gt > 0
    (pass an already-binarised gt, e.g. the reference's own thresholded prediction for agreement-F1)."""
    p = (pred >= thr).astype(np.float32)
    g = (gt > 0).astype(np.float32)
    tp = (p * g).sum(-1)
    prec = tp / (p.sum(-1) + eps)
    rec = tp / (g.sum(-1) + eps)
    f1 = 2 * prec * rec / (prec + rec + eps)
    return float(f1.mean()), float(prec.mean()), float(rec.mean())
