"""CPU: the numpy oracle (oracle/) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_goldens.py, run in the build container with /root/reference imported under ref_shim)."""
from pathlib import Path

import numpy as np
import pytest

from interactvlm_b200 import synthetic as S
from oracle import lift as OL
from oracle.make_goldens import LIFT_SEED, OBJ_NVERTS

GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def gold_lift():
    return np.load(GOLD / "lift.npz")


@pytest.fixture(scope="module")
def logits():
    return S.make_mask_logits(2, seed=LIFT_SEED["logits"])


def test_lift_human_matches_reference(gold_lift, logits):
    p2v, bary = S.make_mesh_lift_maps(seed=LIFT_SEED["maps"])
    out = OL.lift_human(logits, p2v, bary, S.N_SMPL)
    ref = gold_lift["human"]
    assert np.abs(out - ref).max() < 1e-6          # fp32 reduction-order tolerance
    assert np.array_equal(out >= 0.5, ref >= 0.5)  # bit-exact contact vertex set
    assert 0.2 < (ref >= 0.5).mean() < 0.8         # the synthetic case is not degenerate


def test_lift_object_mesh_matches_reference(gold_lift, logits):
    p2v, bary = S.make_mesh_lift_maps(n_verts=OBJ_NVERTS, seed=LIFT_SEED["obj"], coverage=0.25)
    out = OL.lift_object_mesh(logits[:1], p2v, bary, OBJ_NVERTS, thr=0.3)
    ref = gold_lift["object_mesh"]
    assert np.abs(out - ref).max() < 1e-6
    assert np.array_equal(out >= 0.5, ref >= 0.5)


def test_lift_points_matches_reference(gold_lift, logits):
    p2p = S.make_point_lift_maps(seed=LIFT_SEED["points"])
    heat = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    out = OL.lift_points(heat, p2p, 2048)
    assert np.abs(out - gold_lift["points"]).max() < 1e-6


def test_convert_contacts_matches_reference(gold_lift):
    mapping = S.make_smplx_matrix(seed=0)
    out = OL.convert_contacts(gold_lift["human"], mapping)
    assert np.abs(out - gold_lift["smplx"]).max() < 1e-6


def test_lift_edge_cases():
    # empty view (all background), out-of-range ids, a vertex seen by no view
    p2v = np.full((2, 8, 8, 3), -1, np.int64)
    bary = np.zeros((2, 8, 8, 3), np.float32)
    p2v[0, 0, 0] = [0, 1, 2]; bary[0, 0, 0] = [0.5, 0.25, 0.25]
    p2v[0, 0, 1] = [0, 1, 99]; bary[0, 0, 1] = [0.3, 0.3, 0.4]   # 99 >= n_verts -> pixel dropped
    masks = np.zeros((1, 2, 8, 8), np.float32)
    masks[0, 0, 0, 0] = 30.0  # clamped to 20
    out = OL.lift_human(masks, p2v, bary, 4)
    p = 1 / (1 + np.exp(-20.0))
    assert np.allclose(out[0, :3], p, atol=1e-7) and out[0, 3] == 0
    f1, pr, rc = OL.f1_metrics(out, (out >= 0.5).astype(np.float32))
    assert abs(f1 - 1.0) < 1e-6
