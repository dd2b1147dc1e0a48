"""The oracle restatement reproduces the golden vectors minted from the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, TRAIN_CASES, load_train_case, oracle_run
from oracle import ref_bpr


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_trajectory_matches_reference(name):
    case = load_train_case(name)
    model, outs = oracle_run(case)
    np.testing.assert_allclose([o["bpr_loss"].item() for o in outs], case["bpr_loss"], rtol=1e-6)
    np.testing.assert_allclose([o["l2_reg"].item() for o in outs], case["l2_reg"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose([o["logits"].abs().mean().item() for o in outs],
                               case["logits_abs_mean"], rtol=1e-6)
    np.testing.assert_allclose(model.user_emb.detach().numpy(), case["final_user"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(model.item_emb.detach().numpy(), case["final_item"], rtol=1e-6, atol=1e-8)
    if case["bias"]:
        np.testing.assert_allclose(model.item_bias.detach().numpy(), case["final_item_bias"],
                                   rtol=1e-6, atol=1e-8)


def test_metrics_match_reference():
    z = np.load(GOLDEN / "metrics.npz")
    out, tgt = torch.as_tensor(z["output"]), torch.as_tensor(z["target"])
    for k in (1, 5, 20, 100):
        np.testing.assert_allclose(ref_bpr.ndcg_at_k(out, tgt, k).numpy(), z[f"ndcg@{k}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ref_bpr.recall_at_k(out, tgt, k).numpy(), z[f"recall@{k}"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ref_bpr.ndcg_at_k(out, tgt, 20).mean().item(), z["ndcg@20_stream"], rtol=1e-6)
    o, t = torch.as_tensor(z["kat_output"]), torch.as_tensor(z["kat_target"])
    np.testing.assert_allclose(ref_bpr.ndcg_at_k(o, t, 3).numpy(), [0.38685283, 0.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(ref_bpr.ndcg_at_k(o, t, 3).numpy(), z["kat_ndcg@3"], atol=1e-7)
    np.testing.assert_allclose(ref_bpr.recall_at_k(o, t, 3).numpy(), [0.5, 0.0, 1.0], atol=1e-7)


def test_reference_style_sampler_matches_reference():
    z = np.load(GOLDEN / "sampler.npz")
    if str(z["torch_version"]) != torch.__version__:
        pytest.skip("torch.multinomial stream is only pinned for the torch version that minted it")
    seen = ref_bpr.padded_seen(z["indptr"], z["indices"], torch.as_tensor(z["users"]))
    gen = torch.Generator().manual_seed(int(z["seed"]))
    negs = ref_bpr.reference_style_negatives(torch.ones(int(z["I"])), seen, gen)
    assert negs.squeeze(-1).tolist() == z["negs"].tolist()


def test_sampling_weights_exclude_seen_and_padding():
    z = np.load(GOLDEN / "sampler.npz")
    users = torch.as_tensor(z["users"])
    seen = ref_bpr.padded_seen(z["indptr"], z["indices"], users)
    w = ref_bpr.sampling_weights(torch.ones(int(z["I"])), seen)
    assert (w[:, 0] == 0).all()
    for r, u in enumerate(users.tolist()):
        row = z["indices"][z["indptr"][u]:z["indptr"][u + 1]]
        assert (w[r, torch.as_tensor(row, dtype=torch.long)] == 0).all()
        np.testing.assert_allclose(w[r].sum().item(), 1.0, rtol=1e-6)
