"""The adaptive-sampler restatement against the vectors minted from the reference's own
AdaptiveSampler (tests/golden/adaptive.npz).  CPU only."""
import numpy as np

from helpers import GOLDEN
from oracle import adaptive


def test_update_stats_matches_reference():
    z = np.load(GOLDEN / "adaptive.npz")
    snap, std = adaptive.update_stats(z["item_emb"])
    np.testing.assert_array_equal(snap, z["snapshot"])
    np.testing.assert_allclose(std, z["factor_std"], rtol=1e-6)


def test_pick_given_reference_draws_matches_reference_items():
    z = np.load(GOLDEN / "adaptive.npz")
    for r, u in enumerate(z["users"]):
        f = int(z["factor"][r])
        got = adaptive.reference_pick(z["snapshot"][f], z["seen"][r], float(z["user_emb"][u, f]), int(z["geom"][r]))
        assert got == int(z["negs"][r]), r
        assert got != 0 and got not in set(z["seen"][r].tolist())


def test_counter_based_sample_is_valid_and_deterministic():
    z = np.load(GOLDEN / "adaptive.npz")
    snap, std = adaptive.update_stats(z["item_emb"])
    rows = [z["seen"][r] for r in range(len(z["users"]))]
    a = adaptive.sample(z["user_emb"], snap, std, z["users"], rows, 1, 0.2, seed=7, step=3)
    b = adaptive.sample(z["user_emb"], snap, std, z["users"], rows, 1, 0.2, seed=7, step=3)
    c = adaptive.sample(z["user_emb"], snap, std, z["users"], rows, 1, 0.2, seed=7, step=4)
    assert (a == b).all() and not (a == c).all()
    for r in range(len(rows)):
        assert a[r, 0] > 0 and a[r, 0] not in set(rows[r].tolist())
