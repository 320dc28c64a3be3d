"""Evaluation-side reports that sit directly behind the hot path in `evaluate.py`: the geodesic contact
error, the DAMON semantic / category / binary reports, the affordance metrics and the `{dataset}_results.pkl` file.
Host-side bookkeeping over per-sample prediction vectors (numpy / torch on whatever device the vectors live on); the hot
path itself is `model.evaluate()`.

  h_geo_metric            utils/eval_utils.py:127-150  (DIST_MATRIX = ./data/smpl_neutral_geodesic_dist.npy, [6890,6890])
  o_affordance_metrics    utils/eval_utils.py:152-211  (SIM, MAE, AUC, aIoU over 20 thresholds)
  damon_semantic_contact  evaluate.py:355-427
  damon_binary_contact    evaluate.py:430-468
  save_results            evaluate.py:594-597
"""
from __future__ import annotations

import os

import numpy as np
import torch

# preprocess_data/constants.py:388-409
DAMON_CATEGORIES_MAPPING = {
    "transport": ["motorcycle", "bicycle", "boat", "car", "truck", "bus", "train", "airplane"],
    "accessory": ["backpack", "tie", "handbag", "baseball_glove"],
    "furniture": ["bench", "chair", "couch", "bed", "toilet", "dining_table"],
    "everyday-objects": ["book", "umbrella", "cell_phone", "laptop", "kite", "suitcase", "bottle", "remote", "toothbrush",
                         "teddy_bear", "scissors", "keyboard", "hair drier", "traffic light", "fire_hydrant", "stop sign", "tv", "vase",
                         "parking meter", "clock", "potted plant", "mouse"],
    "sports": ["frisbee", "sports_ball", "tennis_racket", "baseball_bat", "skateboard", "snowboard", "skis", "surfboard"],
    "food": ["banana", "cake", "apple", "carrot", "pizza", "donut", "hot_dog", "sandwich", "broccoli", "orange"],
    "kitchen": ["knife", "spoon", "cup", "wine_glass", "oven", "fork", "bowl", "refrigerator", "toaster", "sink", "microwave"],
}


def load_dist_matrix(data_root="./data"):
    """The SMPL geodesic distance matrix the reference loads at import time (utils/eval_utils.py:15)."""
    return torch.tensor(np.load(os.path.join(data_root, "smpl_neutral_geodesic_dist.npy")))


def h_geo_metric(pred, gt, dist_matrix):
    """get_h_geo_metric: mean geodesic distance from every predicted contact vertex to the nearest ground-truth contact vertex
    (false-positive distance) and from every ground-truth vertex to the nearest predicted one (false-negative distance),
    averaged over the batch.  With no ground-truth (or no predicted) contact the full matrix stands in, like the reference."""
    gt, pred = torch.as_tensor(gt).detach(), torch.as_tensor(pred).detach()
    D = dist_matrix.to(pred.device)
    fp, fn = torch.zeros(gt.shape[0]), torch.zeros(gt.shape[0])
    for b in range(gt.shape[0]):
        g, p = gt[b] == 1, pred[b] >= 0.5
        cols = D[:, g] if bool(g.any()) else D
        err = cols[p, :] if bool(p.any()) else cols
        fp[b] = err.min(dim=1)[0].mean()
        fn[b] = err.min(dim=0)[0].mean()
    return fp.mean().item(), fn.mean().item()


def _sim(a, b, eps=1e-12):
    a, b = a / (a.sum() + eps), b / (b.sum() + eps)
    return torch.min(a, b).sum()


def _auc(labels: np.ndarray, scores: np.ndarray) -> float:
    """Area under the ROC curve = Mann-Whitney U statistic with average ranks for ties (what sklearn.roc_auc_score returns)."""
    order = np.argsort(scores, kind="mergesort")
    ranks = np.empty(len(scores), np.float64)
    s = scores[order]
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    pos = labels == 1
    n1, n0 = int(pos.sum()), int((~pos).sum())
    return float((ranks[pos].sum() - n1 * (n1 + 1) / 2.0) / (n1 * n0))


def o_affordance_metrics(contact_gt, contact_pred):
    """get_o_affordance_metrics -> (SIM, MAE, AUC, aIoU, valid samples); samples whose thresholded ground truth holds one class
    only do not count towards AUC / aIoU."""
    gt, pr = torch.as_tensor(contact_gt).float(), torch.as_tensor(contact_pred).float()
    B = gt.shape[0]
    thresholds = np.linspace(0, 1, 20)
    sim_t = mae_t = auc_t = iou_t = 0.0
    valid = B
    for b in range(B):
        sim_t += _sim(gt[b], pr[b]).item()
        mae_t += (torch.sum(torch.abs(gt[b] - pr[b])) / 2048).item()
        g = (gt[b] >= 0.5).int()
        if len(torch.unique(g)) == 1:
            valid -= 1
            continue
        auc_t += _auc(g.cpu().numpy(), pr[b].cpu().numpy().astype(np.float64))
        ious = []
        for t in thresholds:
            p = (pr[b] >= t).int()
            ious.append(1.0 * torch.sum(p & g) / torch.sum(p | g))
        iou_t += torch.tensor(ious).mean().item()
    return sim_t / B, mae_t / B, auc_t / max(1, valid), iou_t / max(1, valid), valid


def _group_stats(saved, indices):
    preds, gts = [saved["pred"][i] for i in indices], [saved["gt"][i] for i in indices]
    tp = sum(np.sum(np.logical_and(p, g)) for p, g in zip(preds, gts))
    pred_pos, gt_pos = sum(np.sum(p) for p in preds), sum(np.sum(g) for g in gts)
    return {"num_samples": len(indices), "avg_f1": np.mean([saved["f1"][i] for i in indices]),
            "precision": tp / pred_pos if pred_pos > 0 else 0, "recall": tp / gt_pos if gt_pos > 0 else 0,
            "geo": np.mean([saved["geo"][i] for i in indices])}


def damon_semantic_contact(saved, verbose=True):
    """get_damon_semantic_contact: per-object and per-category statistics of the saved DAMON results (`pred` / `gt` rows,
    per-sample `f1` / `geo`, `objnames` as the dataloader nests them: [[name]]).  Note that, like the reference, precision
    and recall use the RAW prediction values as weights (`np.logical_and(p, g)` on probabilities), not thresholded ones.
    -> dict(objects, weighted_f1, weighted_geo, categories)."""
    names = [o[0][0].lower() for o in saved["objnames"]]
    by_obj = {}
    for i, n in enumerate(names):
        by_obj.setdefault(n, []).append(i)
    objects = {n: _group_stats(saved, idx) for n, idx in by_obj.items()}
    total = sum(r["num_samples"] for r in objects.values())
    wf1 = sum(r["avg_f1"] * r["num_samples"] for r in objects.values()) / total
    wgeo = sum(r["geo"] * r["num_samples"] for r in objects.values()) / total
    cats = {}
    for cat, members in DAMON_CATEGORIES_MAPPING.items():
        idx = [i for i, n in enumerate(names) if n in members]
        if idx:
            cats[cat] = _group_stats(saved, idx)
    if verbose:
        print("\n[DAMON-HCONTACT - Semantic Contact]")
        print(f"Weighted F1: {wf1:.4f}, Weighted Geo: {wgeo:.4f}")
        print("\n[DAMON-HCONTACT - Semantic Contact Category Summary]")
        print(f"{'Category':20} | {'Samples':>7} | {'F1':>6} | {'Prec':>6} | {'Recall':>6} | {'Geo':>6}")
        print("-" * 70)
        for cat, m in cats.items():
            print(f"{cat:20} | {m['num_samples']:7d} | {m['avg_f1']:.4f} | {m['precision']:.4f} | {m['recall']:.4f} | {m['geo']:.4f}")
    return {"objects": objects, "weighted_f1": wf1, "weighted_geo": wgeo, "categories": cats}


def damon_binary_contact(saved, threshold=0.5, verbose=True):
    """get_damon_binary_contact: the per-object predictions of one image are OR-ed (prediction >= threshold, gt > 0, geo = max),
    then F1 is averaged over images and precision / recall are pooled."""
    img = {}
    for i, name in enumerate(saved["imgnames"]):
        key = name[0]
        p, g = np.asarray(saved["pred"][i]) >= threshold, np.asarray(saved["gt"][i]) > 0
        if key not in img:
            img[key] = {"pred": p, "gt": g, "geo": saved["geo"][i]}
        else:
            img[key] = {"pred": np.logical_or(img[key]["pred"], p), "gt": np.logical_or(img[key]["gt"], g),
                        "geo": max(img[key]["geo"], saved["geo"][i])}
    f1s, geos, tp, pp, gp = [], [], 0, 0, 0
    for v in img.values():
        tpi, ppi, gpi = np.sum(np.logical_and(v["pred"], v["gt"])), np.sum(v["pred"]), np.sum(v["gt"])
        prec, rec = (tpi / ppi if ppi else 0), (tpi / gpi if gpi else 0)
        f1s.append(2 * prec * rec / (prec + rec) if (prec + rec) else 0)
        geos.append(v["geo"])
        tp, pp, gp = tp + tpi, pp + ppi, gp + gpi
    out = {"f1": np.mean(f1s), "precision": tp / pp if pp else 0, "recall": tp / gp if gp else 0, "geo": np.mean(geos),
           "num_images": len(img)}
    if verbose:
        print(f"\n[DAMON-HCONTACT - Binary Contact @ threshold={threshold}]")
        print(f"Global F1: {out['f1']:.4f}, Precision: {out['precision']:.4f}, Recall: {out['recall']:.4f}, Geo: {out['geo']:.4f}")
    return out


def collect_hcontact_results(preds, gts, imgnames, objnames, dist_matrix=None, threshold=0.5):
    """The `saved_results_hC` dictionary evaluate.py:131-200 assembles sample by sample, from batched predictions [n,6890]."""
    from .harness import h_contact_metrics

    preds, gts = torch.as_tensor(preds).float(), torch.as_tensor(gts).float()
    saved = {"imgnames": [], "objnames": [], "pred": [], "gt": [], "f1": [], "geo": []}
    f1s, ps, rs, geos = [], [], [], []
    for i in range(preds.shape[0]):
        f1, p, r = h_contact_metrics(gts[i:i + 1], preds[i:i + 1], threshold)
        geo = h_geo_metric(preds[i:i + 1], gts[i:i + 1], dist_matrix)[0] if dist_matrix is not None else 0.0
        saved["imgnames"].append([imgnames[i]])
        saved["objnames"].append([[objnames[i]]])
        saved["pred"].append(preds[i:i + 1].cpu().numpy())
        saved["gt"].append(gts[i:i + 1].cpu().numpy())
        saved["f1"].append(f1)
        saved["geo"].append(geo)
        f1s.append(f1); ps.append(p); rs.append(r); geos.append(geo)
    saved["pred"], saved["gt"] = np.vstack(saved["pred"]), np.vstack(saved["gt"])
    saved.update(avg_f1=float(np.mean(f1s)), avg_precision=float(np.mean(ps)), avg_recall=float(np.mean(rs)), avg_geo=float(np.mean(geos)))
    return saved


def save_results(saved, log_dir, dataset_name):
    """`{log_dir}/{dataset}_results.pkl` (evaluate.py:594-597, joblib)."""
    import joblib

    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"{dataset_name}_results.pkl")
    joblib.dump(saved, path)
    return path
