"""interactvlm_b200/reports.py against goldens produced by the reference's own functions (oracle/make_goldens_reports.py cuts
get_damon_semantic_contact / get_damon_binary_contact out of evaluate.py and get_h_geo_metric / get_o_affordance_metrics out of
utils/eval_utils.py and runs them unmodified): the printed report text must match character for character."""
import contextlib
import io
from pathlib import Path

import joblib
import numpy as np
import torch

from interactvlm_b200 import reports as R
from oracle.make_goldens_reports import dist_matrix, make_saved

GOLD = np.load(Path(__file__).parent / "golden" / "reports.npz")


def test_damon_reports_print_what_the_reference_prints():
    saved = make_saved()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        sem = R.damon_semantic_contact(saved)
        b5 = R.damon_binary_contact(saved)
        R.damon_binary_contact(saved, threshold=0.3)
    assert buf.getvalue() == str(GOLD["report"])
    assert set(sem["categories"]) == set(R.DAMON_CATEGORIES_MAPPING) and b5["num_images"] == 20
    assert "unknown_thing" in sem["objects"] and "cup" in sem["objects"]          # names are lower-cased, unknown ones kept per object


def test_geodesic_and_affordance_metrics():
    saved = make_saved()
    D = torch.tensor(dist_matrix())
    fp, fn = R.h_geo_metric(torch.from_numpy(saved["pred"][:8]), torch.from_numpy(saved["gt"][:8]), D)
    assert abs(fp - float(GOLD["geo"][0])) < 1e-6 and abs(fn - float(GOLD["geo"][1])) < 1e-6
    g = np.random.default_rng(7)
    agt, apr = g.random((6, 2048)).astype(np.float32), g.random((6, 2048)).astype(np.float32)
    agt[2] = 0.9
    apr[4, :50] = apr[4, 50:100]
    got = R.o_affordance_metrics(agt, apr)
    assert np.allclose(np.array(got, dtype=np.float64), GOLD["aff"], rtol=0, atol=1e-6), (got, GOLD["aff"])


def test_collect_and_save_results(tmp_path):
    g = np.random.default_rng(1)
    pred, gt = g.random((5, 300)).astype(np.float32), (g.random((5, 300)) < 0.2).astype(np.float32)
    saved = R.collect_hcontact_results(pred, gt, [f"i{k}.jpg" for k in range(5)], ["chair", "cup", "car", "bed", "kite"],
                                       torch.tensor(dist_matrix()))
    assert saved["pred"].shape == (5, 300) and len(saved["f1"]) == 5 and 0 <= saved["avg_f1"] <= 1 and saved["avg_geo"] > 0
    path = R.save_results(saved, tmp_path / "logs", "damon_hcontact")
    assert path.endswith("damon_hcontact_results.pkl")
    back = joblib.load(path)
    assert np.array_equal(back["pred"], saved["pred"]) and back["objnames"][1] == [["cup"]]
    with contextlib.redirect_stdout(io.StringIO()):
        assert R.damon_semantic_contact(back)["objects"]["cup"]["num_samples"] == 1
