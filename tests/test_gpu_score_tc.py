"""The tensor-core scoring path (csrc/score_tc.cu: tcgen05 TF32 candidate filter + exact fp32
rescoring) against the dense fp32 path (csrc/score.cu) — by construction the two must agree BIT FOR
BIT on every output (the candidate set is a superset of the exact top-k and the rescoring uses the
dense kernel's arithmetic) — and against the oracle's eval sequence (metrics within 1e-4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(U, I, D, bias, seed, scale=0.3, popular=False):
    from rbpr import synth
    from rbpr.engine import Engine
    inter = synth.generate("t", U - 1, I - 1, (U - 1) * 30, 25, 5, 0.9, seed)
    g = torch.Generator().manual_seed(seed)
    ue = torch.randn(U, D, generator=g) * scale
    ie = torch.randn(I, D, generator=g) * scale
    if popular:  # a few items with much larger norms (trained models look like this): widens the margin
        ie[torch.randint(1, I, (50,), generator=g)] *= 6.0
    ue[0] = 0
    ie[0] = 0
    ib = torch.randn(I, generator=g) * 0.2 if bias else None
    eng = Engine(ue.to(DEV), ie.to(DEV), None if ib is None else ib.to(DEV))
    users, seen, held = synth.split_heldout(inter, n_eval_users=min(700, U - 1), frac=0.2, seed=seed)
    return eng, inter, ue, ie, ib, users, seen, held


def _run(eng, users, seen, held, ks, monkeypatch, tensor):
    if tensor:
        monkeypatch.delenv("RBPR_NO_TC_SCORE", raising=False)
    else:
        monkeypatch.setenv("RBPR_NO_TC_SCORE", "1")
    t = lambda a, b: (torch.as_tensor(a), torch.as_tensor(b))  # noqa: E731
    p0 = eng.score_path_counts()
    top = eng.score_topk(torch.as_tensor(users), t(*seen), t(*held), ks, k_max=max(ks))
    met = eng.score_metrics(torch.as_tensor(users), t(*seen), t(*held), ks, want=("ndcg", "ndcg_linear", "recall", "precision", "map"),
                            want_items=True)
    eng.sync_check()
    p1 = eng.score_path_counts()
    return top, met, (p1[0] - p0[0], p1[1] - p0[1])


@pytest.mark.parametrize("U,I,D,bias,popular", [(900, 9000, 128, False, False), (800, 12345, 128, True, True),
                                                 (600, 8700, 128, False, False),  # odd number of user tiles in pair mode
                                                 (500, 8300, 64, True, False), (400, 8200, 20, False, False),
                                                 (300, 9100, 256, True, False)])
def test_tensor_path_equals_dense_path_bit_for_bit(U, I, D, bias, popular, monkeypatch):
    from oracle import ref_bpr
    eng, inter, ue, ie, ib, users, seen, held = _problem(U, I, D, bias, 7 + D, popular=popular)
    ks = [1, 5, 20, 100]
    top_d, met_d, used_d = _run(eng, users, seen, held, ks, monkeypatch, tensor=False)
    top_t, met_t, used_t = _run(eng, users, seen, held, ks, monkeypatch, tensor=True)
    assert used_d[0] == 0 and used_t[0] == 2  # two calls, one block of users each, all on the tensor path
    assert used_t[1] == 0                      # nobody overflowed
    assert torch.equal(top_t["items"], top_d["items"])
    assert torch.equal(top_t["scores"], top_d["scores"])
    assert torch.equal(met_t["items"], met_d["items"])
    for k in ("ndcg", "ndcg_linear", "recall", "precision", "map"):
        assert torch.equal(met_t[k], met_d[k]), k
    for k in ("ndcg", "recall"):
        assert torch.equal(top_t[k], top_d[k]), k
    # ... and the oracle's eval sequence on a sample of the users
    rows = np.arange(0, len(users), 5)
    model = ref_bpr.RefModel(ue, ie, ib)
    seen_pad = torch.nn.utils.rnn.pad_sequence(
        [torch.as_tensor(seen[1][seen[0][r]:seen[0][r + 1]], dtype=torch.long) for r in rows], batch_first=True)
    logits = model.eval_logits(torch.as_tensor(users[rows]), seen_pad)
    target = torch.zeros(rows.size, I)
    for q, r in enumerate(rows):
        target[q, torch.as_tensor(held[1][held[0][r]:held[0][r + 1]], dtype=torch.long)] = 1.0
    np.testing.assert_allclose(met_t["ndcg"].cpu().numpy()[rows][:, 3], ref_bpr.ndcg_at_k(logits, target, 100).numpy(), atol=1e-4)
    np.testing.assert_allclose(met_t["recall"].cpu().numpy()[rows][:, 2], ref_bpr.recall_at_k(logits, target, 20).numpy(), atol=1e-4)


def test_tensor_path_hands_mass_ties_to_the_dense_path(monkeypatch):
    """Users whose scores tie en masse (an all-zero user row, as in an untrained table) overflow the
    candidate list; they are re-done by the dense path inside the same call and get its results."""
    eng, inter, ue, ie, ib, users, seen, held = _problem(600, 8500, 32, False, 3)
    zero_rows = [3, 77, 300]
    with torch.no_grad():
        for r in zero_rows:
            eng.user_emb[users[r]] = 0
    ks = [10, 100]
    top_d, met_d, _ = _run(eng, users, seen, held, ks, monkeypatch, tensor=False)
    top_t, met_t, used_t = _run(eng, users, seen, held, ks, monkeypatch, tensor=True)
    assert used_t == (2, 2 * len(zero_rows))
    assert torch.equal(top_t["items"], top_d["items"]) and torch.equal(top_t["scores"], top_d["scores"])
    for k in ("ndcg", "recall", "precision", "map"):
        assert torch.equal(met_t[k], met_d[k]), k
    # ties rank by ascending item id: the zero rows list the first unmasked items
    for r in zero_rows:
        s = set(seen[1][seen[0][r]:seen[0][r + 1]].tolist())
        want = [i for i in range(1, 400) if i not in s][:100]
        assert top_t["items"][r].cpu().tolist() == want
