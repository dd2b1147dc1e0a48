"""GPU parity tests of the adaptive sampler through the C ABI and the AdaptiveSampler mirror."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _engine(z):
    from rbpr.engine import Engine
    eng = Engine(torch.as_tensor(z["user_emb"]).to(DEV).contiguous(), torch.as_tensor(z["item_emb"]).to(DEV).contiguous())
    eng.bind_csr(torch.as_tensor(z["indptr"]), torch.as_tensor(z["indices"]))
    return eng


def test_stats_match_reference_snapshot():
    from oracle import adaptive
    z = np.load(GOLDEN / "adaptive.npz")
    eng = _engine(z)
    eng.adaptive_update_stats()
    st = eng.adaptive_stats()
    np.testing.assert_allclose(st["std"].cpu().numpy(), z["factor_std"], rtol=1e-6)
    snap = z["snapshot"]
    order = st["order"].cpu().numpy()
    pos = st["pos"].cpu().numpy()
    for f in range(snap.shape[0]):
        np.testing.assert_array_equal(order[f], np.argsort(-snap[f], kind="stable"))
        np.testing.assert_array_equal(pos[f][order[f]], np.arange(snap.shape[1]))
    del adaptive


@pytest.mark.parametrize("p", [0.2, 0.02, 0.9])
def test_padded_sampling_bit_exact_vs_oracle(p):
    from oracle import adaptive
    z = np.load(GOLDEN / "adaptive.npz")
    eng = _engine(z)
    eng.adaptive_update_stats()
    snap, std = adaptive.update_stats(z["item_emb"])
    std = eng.adaptive_stats()["std"].cpu().numpy()  # same fp32 std on both sides
    users, seen = torch.as_tensor(z["users"]), torch.as_tensor(z["seen"])
    for step in (0, 5):
        got = eng.sample_adaptive_padded(users, seen, 1, p, seed=0xABCDEF0123, step=step).cpu().numpy()
        eng.sync_check()
        exp = adaptive.sample(z["user_emb"], snap, std, z["users"], [z["seen"][r] for r in range(len(z["users"]))],
                              1, p, seed=0xABCDEF0123, step=step)
        assert got.tolist() == exp.tolist()


def test_adaptive_sampler_class_and_fused_path():
    """AdaptiveSampler mirror (update_stats / sample / every) and the fast path (sampler kind
    ADAPTIVE inside rbpr_train_steps): negatives re-derived by the oracle from the model state
    step by step, then the step itself checked against the oracle."""
    from oracle import adaptive, ref_bpr
    from rbpr import native
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    from revisit_bpr.modules import AdaptiveSampler
    z = np.load(GOLDEN / "adaptive.npz")
    U, D = z["user_emb"].shape
    I = z["item_emb"].shape[0]
    model = BPR(MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0)),
                reg_alphas={"user": 0.01, "item": 0.02})
    with torch.no_grad():
        model.logits_model._user_emb.weight.copy_(torch.as_tensor(z["user_emb"]))
        model.logits_model._item_emb.weight.copy_(torch.as_tensor(z["item_emb"]))
    model = model.to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(21)
    s = AdaptiveSampler(model, I, 0.2, gen, every=2)
    with pytest.raises(AttributeError):
        s.sample({"user": torch.as_tensor(z["users"]).to(DEV), "item": torch.zeros(96, 1, dtype=torch.long, device=DEV),
                  "seen_items": torch.as_tensor(z["seen"]).to(DEV)})
    s.update_stats()
    batch = {"user": torch.as_tensor(z["users"]).to(DEV), "item": torch.zeros(96, 1, dtype=torch.long, device=DEV),
             "seen_items": torch.as_tensor(z["seen"]).to(DEV)}
    a = s.sample(batch)
    assert a.shape == (96, 1)
    for r in range(96):
        assert a[r, 0].item() > 0 and a[r, 0].item() not in set(z["seen"][r].tolist())
    # ---- fast path: triple ids, adaptive negatives drawn from the CURRENT user rows each step
    eng = model.logits_model.engine()
    eng.bind_csr(torch.as_tensor(z["indptr"]), torch.as_tensor(z["indices"]))
    eng.set_reg({"user": 0.01, "item": 0.02})
    eng.set_sgd(0.05)
    eng.set_adaptive(0.2, every=2)
    eng.adaptive_update_stats()
    ref = ref_bpr.RefModel(torch.as_tensor(z["user_emb"]), torch.as_tensor(z["item_emb"]), None, {"user": 0.01, "item": 0.02})
    opt = ref_bpr.make_optimizer(ref, "sgd", lr=0.05)
    coo = np.repeat(np.arange(U), np.diff(z["indptr"]))
    nnz = z["indices"].size
    B, steps, seed = 64, 5, 99
    perm = torch.randperm(nnz, generator=torch.Generator().manual_seed(1))[:B * steps]
    stats, negs = eng.train_steps(perm.to(DEV), B, seed, 0, want_neg=True)
    eng.sync_check()
    negs = negs.cpu().numpy()
    snap, std = adaptive.update_stats(ref.item_emb.detach().numpy())
    for st in range(steps):
        t = perm[st * B:(st + 1) * B].numpy()
        rows = [z["indices"][z["indptr"][u]:z["indptr"][u + 1]] for u in coo[t]]
        std32 = std
        exp = adaptive.sample(ref.user_emb.detach().numpy(), snap, std32, coo[t], rows, 1, 0.2, seed, st,
                              subsequences=t)[:, 0]
        assert negs[st * B:(st + 1) * B].tolist() == exp.tolist(), st
        if (st + 1) % 2 == 0:  # refresh after the draw, before the update
            snap, std = adaptive.update_stats(ref.item_emb.detach().numpy())
        out = ref_bpr.train_step(ref, opt, torch.as_tensor(coo[t]), torch.as_tensor(z["indices"][t], dtype=torch.long),
                                 torch.as_tensor(exp))
        np.testing.assert_allclose(stats[st, 0].item(), out["bpr_loss"].item(), rtol=1e-4)
    np.testing.assert_allclose(eng.user_emb.cpu().numpy(), ref.user_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), ref.item_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    del native
