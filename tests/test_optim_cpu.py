"""oracle/optim.py against the goldens recorded from the reference's own contact_loss body (tests/golden/contact_loss.npz,
made by oracle/make_goldens_optim.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import optim as OO
from oracle.make_goldens_optim import CASES, inputs

GOLD = np.load(Path(__file__).parent / "golden" / "contact_loss.npz")


def case(name):
    seed, n_obj, n_hum = CASES[name]
    obj, hum, p, q = inputs(seed, n_obj, n_hum)
    if name == "coincident":
        hum[:16] = obj[:16]
    return obj, hum, p, q


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_goldens(name):
    obj, hum, p, q = case(name)
    loss, grad = OO.contact_loss(obj, hum, p, q)
    assert abs(loss - GOLD[f"{name}_f64_loss"]) < 1e-12
    assert np.abs(grad - GOLD[f"{name}_f64_grad"]).max() < 1e-12
    # the reference's own float32 run (cdist through the |a|^2+|b|^2-2ab expansion) is within 2e-6 of its float64 run
    assert abs(GOLD[f"{name}_f32_loss"] - GOLD[f"{name}_f64_loss"]) < 2e-6
    assert np.isfinite(grad).all()


def test_icp_oracle_matches_reference_goldens():
    """oracle/optim.py icp() against outputs of the reference's own ICP (optim/icp/icp.py imported unmodified with a
    brute-force stand-in for pytorch3d's knn_points; oracle/make_goldens_optim.py)."""
    from oracle.make_goldens_optim import ICP_CASES, icp_inputs

    G = np.load(Path(__file__).parent / "golden" / "icp.npz")
    for name, (seed, n_obj, n_hum, est) in ICP_CASES.items():
        obj, on, hum, hn, R0, T0, s0 = icp_inputs(seed, n_obj, n_hum)
        r = OO.icp(obj[0], hum[0], (R0[0], T0[0], s0[0]), 30, estimate_scale=est, obj_normals=on[0], hum_normals=hn[0])
        assert r["iters"] == int(G[f"{name}_f64_iters"]) and r["converged"] == bool(G[f"{name}_f64_converged"])
        for k in ("R", "T", "Xt"):
            assert np.abs(r[k] - G[f"{name}_f64_{k}"][0]).max() < 1e-12
        assert abs(r["s"] - G[f"{name}_f64_s"][0]) < 1e-12 and abs(r["rmse"] - G[f"{name}_f64_rmse"][0]) < 1e-12
        assert abs(np.linalg.det(r["R"]) - 1) < 1e-9


def test_product_wrapper_refuses_cpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interactvlm_b200 import optim as PO

    with pytest.raises((RuntimeError, ValueError)):
        PO.contact_loss(torch.zeros(4, 3), torch.zeros(5, 3), torch.ones(4), torch.ones(5))


def test_pose_utils_match_reference_goldens():
    """rot6d / transformation / centroid / mask-loss helpers of interactvlm_b200.optim (plain torch, run on CPU here) against
    outputs of the reference's own functions (optim/utils.py, optim/optimizer.py:171-174)."""
    from interactvlm_b200 import optim as PO
    from oracle.make_goldens_optim import utils_inputs

    G = np.load(Path(__file__).parent / "golden" / "icp.npz")
    u = utils_inputs()
    assert np.abs(PO.rot6d_to_matrix(u["rot6d"]).numpy() - G["utils_rot"]).max() < 1e-6
    assert np.abs(PO.matrix_to_rot6d(u["mat"]).numpy() - G["utils_rot6d"]).max() == 0
    assert np.abs(PO.apply_transformation(u["verts"], u["rot6d"], u["trans"], u["scale"]).numpy() - G["utils_transformed"]).max() < 1e-6
    assert np.abs(PO.calculate_centroid(u["mask"]).numpy() - G["utils_centroid"]).max() < 1e-5
    assert np.abs(PO.calculate_centroid(torch.zeros(6, 8)).numpy() - G["utils_centroid_empty"]).max() == 0

    class M:
        target_mask = u["target"]
    assert abs(PO.ObjPose_Opt.mask_loss_iou(M, u["mask"]).item() - float(G["utils_mask_loss"])) < 1e-6
