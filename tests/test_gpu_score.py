"""GPU parity tests of full-catalog scoring + top-k + NDCG/Recall through the C ABI against the
oracle (restated reference eval path).  Tolerance: metrics 1e-4 (north_star); scores 1e-5."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _setup(U, I, D, seed, bias=False):
    from rbpr import synth
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    inter = synth.generate("t", U - 1, I - 1, (U - 1) * 14, 12, 5, 0.8, seed)
    g = torch.Generator().manual_seed(seed)
    ue = torch.randn(U, D, generator=g) * 0.3
    ie = torch.randn(I, D, generator=g) * 0.3
    ue[0] = 0
    ie[0] = 0
    ib = torch.randn(I, generator=g) * 0.2 if bias else None
    eng = Engine(ue.to(dev), ie.to(dev), None if ib is None else ib.to(dev))
    users, seen, held = synth.split_heldout(inter, n_eval_users=min(200, U - 1), frac=0.2, seed=seed)
    return eng, inter, ue, ie, ib, users, seen, held


@pytest.mark.parametrize("U,I,D,bias", [(300, 517, 32, False), (260, 1000, 128, True), (150, 131, 20, False)])
def test_topk_and_metrics_match_oracle(U, I, D, bias):
    from oracle import ref_bpr
    eng, inter, ue, ie, ib, users, seen, held = _setup(U, I, D, 11 + D, bias)
    ks = [1, 5, 10, 20, 50, 100]
    out = eng.score_topk(torch.as_tensor(users), (torch.as_tensor(seen[0]), torch.as_tensor(seen[1])),
                         (torch.as_tensor(held[0]), torch.as_tensor(held[1])), ks, k_max=100)
    torch.cuda.synchronize()
    model = ref_bpr.RefModel(ue, ie, ib)
    seen_pad = torch.nn.utils.rnn.pad_sequence(
        [torch.as_tensor(seen[1][seen[0][r]:seen[0][r + 1]], dtype=torch.long) for r in range(len(users))],
        batch_first=True, padding_value=0)
    logits = model.eval_logits(torch.as_tensor(users), seen_pad)
    target = ref_bpr.multi_hot(held[0], held[1], I)
    for q, k in enumerate(ks):
        np.testing.assert_allclose(out["ndcg"][:, q].cpu().numpy(), ref_bpr.ndcg_at_k(logits, target, k).numpy(),
                                   atol=1e-4, err_msg=f"ndcg@{k}")
        np.testing.assert_allclose(out["recall"][:, q].cpu().numpy(), ref_bpr.recall_at_k(logits, target, k).numpy(),
                                   atol=1e-4, err_msg=f"recall@{k}")
    # top-k scores equal the oracle's sorted scores
    kk = min(100, I)
    ref_sorted = torch.sort(logits, dim=-1, descending=True).values[:, :kk]
    np.testing.assert_allclose(out["scores"][:, :kk].cpu().numpy(), ref_sorted.numpy(), atol=1e-5, rtol=1e-5)
    # and the items really have those scores, are unique, unmasked
    items = out["items"].cpu().long()
    got = torch.gather(logits, 1, items[:, :kk])
    np.testing.assert_allclose(got.numpy(), ref_sorted.numpy(), atol=1e-5, rtol=1e-5)
    assert all(len(set(r.tolist())) == kk for r in items[:, :kk])


def test_dense_scores_match_oracle():
    from oracle import ref_bpr
    eng, inter, ue, ie, ib, users, seen, held = _setup(180, 333, 48, 3, True)
    out = eng.score_dense(torch.as_tensor(users), (torch.as_tensor(seen[0]), torch.as_tensor(seen[1])))
    model = ref_bpr.RefModel(ue, ie, ib)
    seen_pad = torch.nn.utils.rnn.pad_sequence(
        [torch.as_tensor(seen[1][seen[0][r]:seen[0][r + 1]], dtype=torch.long) for r in range(len(users))],
        batch_first=True, padding_value=0)
    logits = model.eval_logits(torch.as_tensor(users), seen_pad)
    np.testing.assert_allclose(out.cpu().numpy(), logits.numpy(), atol=1e-5, rtol=1e-5)


def test_metric_kats_via_topk_kernel():
    """SURVEY §4 known answers, pushed through the kernel by planting the scores as a rank-1 model."""
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    z = np.load(GOLDEN / "metrics.npz")
    o, t = z["kat_output"], z["kat_target"]  # (3,5): treat columns as items 1..5
    n, m = o.shape
    D = 8
    ue = torch.zeros(n + 1, D)
    ie = torch.zeros(m + 1, D)
    # user r = e_r, item c has coordinate r equal to o[r,c]
    for r in range(n):
        ue[r + 1, r] = 1.0
        ie[1:, r] = torch.as_tensor(o[r])
    eng = Engine(ue.to(dev), ie.to(dev))
    held_ptr, held_idx = [0], []
    for r in range(n):
        pos = (np.nonzero(t[r])[0] + 1).tolist()
        held_idx += pos
        held_ptr.append(len(held_idx))
    out = eng.score_topk(torch.arange(1, n + 1), None,
                         (torch.tensor(held_ptr), torch.tensor(held_idx + [0], dtype=torch.int32)[:len(held_idx)]),
                         [3], k_max=3)
    np.testing.assert_allclose(out["ndcg"][:, 0].cpu().numpy(), z["kat_ndcg@3"], atol=1e-6)
    np.testing.assert_allclose(out["recall"][:, 0].cpu().numpy(), z["kat_recall@3"], atol=1e-6)


def test_topk_ties_and_short_catalog():
    """k larger than the number of unmasked items and exact score ties: deterministic, in-range."""
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    U, I, D = 6, 9, 4
    ue = torch.ones(U, D)
    ie = torch.ones(I, D)  # every score ties
    eng = Engine(ue.to(dev), ie.to(dev))
    out = eng.score_topk(torch.arange(1, U), None, None, [], k_max=20)
    items = out["items"].cpu().numpy()
    # ties resolve to ascending item id; item 0 (masked) ranks last; beyond I -> -1
    assert items[0, :8].tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    assert items[0, 8] == 0 and (items[0, 9:] == -1).all()
