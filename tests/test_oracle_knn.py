"""oracle/knn.py (numpy restatement of ItemKNN / FreeItemKNN + hand-written gradients) against
tests/golden/knn.npz minted from the unmodified reference (tests/golden/make_golden_knn.py)."""
import numpy as np
import pytest

from helpers import GOLDEN
from oracle import knn

CASES = ["itemknn_bias", "itemknn_fused", "itemknn_noreg", "freeknn_bias", "freeknn_fused"]


def load(name):
    z = np.load(GOLDEN / "knn.npz")
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


@pytest.mark.parametrize("name", CASES)
def test_forward_and_gradients_match_reference(name):
    g = load(name)
    kind = name.split("_")[0]
    fuse = name.endswith("fused")
    bias = g.get("b0")
    r = knn.bpr_step(kind, g["w0"], bias, g["item"], g["neg"], g["seen"], reg=tuple(g["reg"]), fuse=fuse)
    np.testing.assert_allclose(r["logits_pos"], g["logits_pos"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(r["logits_neg"], g["logits_neg"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(r["bpr_loss"], g["bpr_loss"], rtol=1e-5)
    np.testing.assert_allclose(r["l2_reg"], g["l2_reg"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r["grad_w"], g["grad_w"], rtol=1e-4, atol=1e-5)
    if bias is not None:
        np.testing.assert_allclose(r["grad_bias"], g["grad_b"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", CASES)
def test_eval_logits_with_colliding_ids(name):
    g = load(name)
    fwd = knn.itemknn_forward if name.startswith("itemknn") else knn.freeknn_forward
    out = fwd(g["w0"], g.get("b0"), g["wide"], g["seen"])
    np.testing.assert_allclose(out, g["eval_logits"], rtol=1e-5, atol=1e-5)
    assert (g["wide"][:, 0] == g["seen"][:, 0]).all()  # the collision the mask is about is in the fixture


@pytest.mark.parametrize("name", CASES)
def test_three_sgd_steps_reach_reference_weights(name):
    g = load(name)
    kind, fuse = name.split("_")[0], name.endswith("fused")
    w, b = g["w0"].astype(np.float64), (g["b0"].astype(np.float64) if "b0" in g else None)
    for loss in g["losses"]:
        r = knn.bpr_step(kind, w, b, g["item"], g["neg"], g["seen"], reg=tuple(g["reg"]), fuse=fuse)
        np.testing.assert_allclose(r["bpr_loss"] + r["l2_reg"], loss, rtol=1e-5)
        w = w - float(g["lr"]) * r["grad_w"]
        if b is not None:
            b = b - float(g["lr"]) * r["grad_bias"]
    np.testing.assert_allclose(w, g["w_end"], rtol=1e-4, atol=1e-5)
    if b is not None:
        np.testing.assert_allclose(b, g["b_end"], rtol=1e-4, atol=1e-6)


def test_keep_mask_drops_ids_present_in_item_list():
    item = np.array([[3, 5], [1, 1]])
    seen = np.array([[5, 7, 0], [2, 0, 0]])
    assert knn.keep_mask(item, seen).tolist() == [[False, True, True], [True, True, True]]
