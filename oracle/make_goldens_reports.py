"""Generates tests/golden/reports.npz from the UNMODIFIED reference: get_damon_semantic_contact / get_damon_binary_contact are
cut out of evaluate.py and get_h_geo_metric / get_o_affordance_metrics / SIM out of utils/eval_utils.py with `ast` (neither
module imports here: deepspeed is absent, eval_utils loads ./data at import) and executed as they are on seeded inputs; their
printed report lines are captured too.  Run in the build container only:  python -m oracle.make_goldens_reports"""
import ast
import contextlib
import io
import json
import textwrap
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "reports.npz"


def _functions(path, names, ns):
    src = path.read_text()
    for n in ast.parse(src).body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            exec(textwrap.dedent(ast.get_source_segment(src, n)), ns)
    return ns


def make_saved(seed=0, n=40, nv=300):
    g = np.random.default_rng(seed)
    objs = ["chair", "motorcycle", "cell_phone", "Cup", "skateboard", "banana", "backpack", "unknown_thing"]
    imgs = [f"img_{i // 2:03d}.jpg" for i in range(n)]           # two objects per image
    pred = g.beta(0.4, 0.9, size=(n, nv)).astype(np.float32)
    gt = (g.random((n, nv)) < 0.15).astype(np.float32)
    gt[3] = 0                                                     # a sample without contact
    pred[5] = 0.1                                                 # a sample with no predicted contact
    return dict(pred=pred, gt=gt, imgnames=[[i] for i in imgs], objnames=[[[objs[k % len(objs)]]] for k in range(n)],
                f1=list(g.random(n)), geo=list(g.random(n) * 0.3))


def dist_matrix(nv=300, seed=1):
    g = np.random.default_rng(seed)
    p = g.normal(size=(nv, 3))
    return np.linalg.norm(p[:, None] - p[None], axis=-1).astype(np.float32)


if __name__ == "__main__":
    from sklearn.metrics import roc_auc_score

    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_constants", REF / "preprocess_data" / "constants.py")
    const = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(const)
    ns = {"np": np, "torch": torch, "DAMON_CATEGORIES_MAPPING": const.DAMON_CATEGORIES_MAPPING}
    _functions(REF / "evaluate.py", {"get_damon_semantic_contact", "get_damon_binary_contact"}, ns)
    D = dist_matrix()
    ns2 = {"np": np, "torch": torch, "DIST_MATRIX": torch.tensor(D), "roc_auc_score": roc_auc_score}
    _functions(REF / "utils" / "eval_utils.py", {"get_h_geo_metric", "get_o_affordance_metrics", "SIM"}, ns2)
    saved = make_saved()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ns["get_damon_semantic_contact"](saved)
        ns["get_damon_binary_contact"](saved)
        ns["get_damon_binary_contact"](saved, threshold=0.3)
    geo = ns2["get_h_geo_metric"](torch.from_numpy(saved["pred"][:8]), torch.from_numpy(saved["gt"][:8]))
    g = np.random.default_rng(7)
    agt, apr = g.random((6, 2048)).astype(np.float32), g.random((6, 2048)).astype(np.float32)
    agt[2] = 0.9                                                  # one-class ground truth -> not a valid AUC sample
    apr[4, :50] = apr[4, 50:100]                                  # tied scores
    with contextlib.redirect_stdout(io.StringIO()):
        aff = ns2["get_o_affordance_metrics"](torch.from_numpy(agt), torch.from_numpy(apr))
    np.savez_compressed(OUT, report=np.array(buf.getvalue()), geo=np.array(geo), aff=np.array(aff, dtype=np.float64))
    print(buf.getvalue())
    print("geo", geo, "aff", aff)
