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


def test_product_wrapper_refuses_cpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interactvlm_b200 import optim as PO

    with pytest.raises((RuntimeError, ValueError)):
        PO.contact_loss(torch.zeros(4, 3), torch.zeros(5, 3), torch.ones(4), torch.ones(5))
