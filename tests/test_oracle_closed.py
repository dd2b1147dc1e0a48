"""The closed-form numpy step against the reference-minted trajectories (first step of each SGD
case) — keeps the full-size checker honest.  CPU only."""
import numpy as np
import pytest

from helpers import load_train_case
from oracle import closed
from oracle.ref_bpr import resolve_reg


@pytest.mark.parametrize("name", ["sgd_reg3", "sgd_bias_all", "sgd_noreg"])
def test_closed_form_step_matches_reference(name):
    case = load_train_case(name)
    ue, ie = case["init_user"].astype(np.float64), case["init_item"].astype(np.float64)
    ib = case["init_item_bias"].astype(np.float64) if case["bias"] else None
    reg, lr = resolve_reg(case["reg"]), case["opt_kw"]["lr"]
    for s in range(case["triples"].shape[0]):
        t = case["triples"][s]
        bpr, l2, upd = closed.sgd_step(ue, ie, case["coo_user"][t], case["indices"][t], case["negs"][s], lr, reg, ib)
        np.testing.assert_allclose(bpr, case["bpr_loss"][s], rtol=2e-6)
        np.testing.assert_allclose(l2, case["l2_reg"][s], rtol=2e-6, atol=1e-9)
        ue[upd["users"]] = upd["user_rows"]
        ie[upd["items"]] = upd["item_rows"]
        if ib is not None:
            ib[upd["items"]] = upd["bias"]
    np.testing.assert_allclose(ue, case["final_user"], atol=2e-6)
    np.testing.assert_allclose(ie, case["final_item"], atol=2e-6)
    if ib is not None:
        np.testing.assert_allclose(ib, case["final_item_bias"], atol=2e-6)
