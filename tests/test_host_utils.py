"""CPU checks of small host-side pieces the kernels rely on: the Walker alias table, the
regularisation defaulting rule, the synthetic interaction generator."""
import numpy as np
import pytest


def test_alias_table_reproduces_the_weights_exactly():
    from rbpr.engine import build_alias
    rng = np.random.default_rng(3)
    for n in (2, 7, 64, 1001):
        w = rng.random(n) ** 3
        w[0] = 0.0  # padding item: never drawn
        w[rng.integers(1, n, size=max(1, n // 10))] = 0.0
        if w.sum() == 0:
            w[1] = 1.0
        prob, alias = build_alias(w.copy())
        assert prob.dtype == np.float32 and alias.dtype == np.int32
        # mass that lands on item k: own column with prob[k], plus (1 - prob[c]) from columns aliased to k
        mass = prob.astype(np.float64).copy()
        np.add.at(mass, alias, 1.0 - prob.astype(np.float64))
        np.testing.assert_allclose(mass / n, w / w.sum(), atol=2e-6)
        zero = w <= 0
        assert (prob[zero] == 0).all() and (~zero[alias[zero]]).all()  # a zero-weight column never returns itself


@pytest.mark.parametrize("reg,expected", [
    (None, (0.0, 0.0, 0.0)), ({}, (0.0, 0.0, 0.0)), ({"all": 0.1}, (0.1, 0.1, 0.1)),
    ({"user": 0.2, "item": 0.3}, (0.2, 0.3, 0.3)),             # neg defaults to item
    ({"user": 0.2, "item": 0.3, "neg": 0.4}, (0.2, 0.3, 0.4)),
    ({"all": 0.1, "user": 0.9}, (0.1, 0.1, 0.1)),               # `all` overrides
    ({"neg": 0.5}, (0.0, 0.0, 0.5)),
])
def test_reg_defaulting_rule_matches_oracle(reg, expected):
    from oracle.ref_bpr import resolve_reg as oracle_rule
    from rbpr.engine import resolve_reg
    assert resolve_reg(reg) == pytest.approx(expected)
    assert oracle_rule(reg) == pytest.approx(expected)


def test_synthetic_interactions_have_the_advertised_shape():
    from rbpr import synth
    inter = synth.make("c1-small", seed=13)
    assert inter.num_users == 1001 and inter.num_items == 501
    assert inter.indptr[0] == inter.indptr[1] == 0  # row 0 is the padding user
    deg = np.diff(inter.indptr)[1:]
    assert deg.min() >= 5 and abs(inter.nnz - 20_000) / 20_000 < 0.15
    assert inter.indices.min() >= 1 and inter.indices.max() < inter.num_items
    for u in (1, 17, 1000):  # rows strictly ascending (the kernels' CSR contract)
        row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
        assert (np.diff(row) > 0).all()
    again = synth.make("c1-small", seed=13)
    assert np.array_equal(again.indices, inter.indices)  # deterministic
    users, seen, held = synth.split_heldout(inter, 50)
    for r, u in enumerate(users):
        s = seen[1][seen[0][r]:seen[0][r + 1]]
        h = held[1][held[0][r]:held[0][r + 1]]
        full = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
        assert np.array_equal(np.sort(np.concatenate([s, h])), full) and h.size >= 1
