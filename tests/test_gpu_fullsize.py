"""BASELINE.json full-size checks (config 2: ML-20M shape, D=128; config 3 shape MSD, D=256, Adam)
through size-independent properties and the vectorised numpy oracles:
  * negatives of a full 65 536-triple step: bit-exact vs the CPU restatement, never seen / padding;
  * the step itself vs the closed-form numpy step (loss 1e-4 relative, touched rows 1e-5);
  * untouched rows are bit-identical before and after (scatter update touches only its rows);
  * with no L2, column sums of the item table are conserved by a step (∂i⁺ + ∂i⁻ = 0 per triple);
  * sampler idempotence: same (seed, step, triples) -> same negatives, regardless of wave splits.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REG = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}


@pytest.fixture(scope="module")
def ml20m():
    from rbpr import synth
    return synth.make("ml-20m", seed=13)


def _tables(inter, D, scale):
    torch.manual_seed(13)
    ue = (torch.rand(inter.num_users, D) - 0.5) / D * scale
    ie = (torch.rand(inter.num_items, D) - 0.5) / D * scale
    ue[0] = 0
    ie[0] = 0
    return ue, ie


def test_full_size_sgd_step_parity_and_properties(ml20m):
    from oracle import closed, philox
    from oracle.ref_bpr import resolve_reg
    from rbpr import native
    from rbpr.engine import Engine
    inter, D, B = ml20m, 128, 65536
    ue, ie = _tables(inter, D, 40.0)
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(REG)
    eng.set_sgd(0.05)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(13))
    t = perm[:B]
    stats, negs = eng.train_steps(t.to(DEV), B, seed=13, step0=7, want_neg=True)
    eng.sync_check()
    coo = inter.coo_users()
    tn = t.numpy()
    exp = philox.sample_negatives(inter.indptr, inter.indices, coo, tn, 13, 7, inter.num_items)
    got = negs.cpu().numpy()
    assert (got == exp).all()
    assert got.min() >= 1 and got.max() < inter.num_items
    u, i = coo[tn], inter.indices[tn].astype(np.int64)
    bpr, l2, upd = closed.sgd_step(ue.numpy(), ie.numpy(), u, i, exp, 0.05, resolve_reg(REG))
    st = stats.cpu().numpy()[0]
    assert st[3] == B
    np.testing.assert_allclose(st[0], bpr, rtol=1e-4)
    np.testing.assert_allclose(st[1], l2, rtol=1e-4)
    new_u, new_i = eng.user_emb.cpu().numpy(), eng.item_emb.cpu().numpy()
    np.testing.assert_allclose(new_u[upd["users"]], upd["user_rows"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(new_i[upd["items"]], upd["item_rows"], atol=1e-5, rtol=1e-4)
    untouched_u = np.setdiff1d(np.arange(inter.num_users), upd["users"])
    untouched_i = np.setdiff1d(np.arange(inter.num_items), upd["items"])
    assert np.array_equal(new_u[untouched_u], ue.numpy()[untouched_u])
    assert np.array_equal(new_i[untouched_i], ie.numpy()[untouched_i])
    assert np.abs(new_u[0]).sum() == 0 and np.abs(new_i[0]).sum() == 0


def test_full_size_item_column_sums_conserved_without_l2(ml20m):
    from rbpr import native
    from rbpr.engine import Engine
    inter, D, B = ml20m, 128, 65536
    ue, ie = _tables(inter, D, 40.0)
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(None)
    eng.set_sgd(0.05)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    before = ie.double().sum(0).numpy()
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(5))[:4 * B]
    stats, _ = eng.train_steps(perm.to(DEV), B, seed=1, step0=0)
    eng.sync_check()
    after = eng.item_emb.double().sum(0).cpu().numpy()
    moved = (eng.item_emb.cpu() - ie).abs().double().sum(0).numpy()
    assert moved.min() > 1e-2  # the table really moved ...
    np.testing.assert_allclose(after, before, atol=1e-6 * moved.max() + 5e-4)  # ... its column sums did not
    assert (stats[:, 3] == B).all().item()


def test_sampler_independent_of_call_and_wave_splitting(ml20m):
    from rbpr import native
    from rbpr.engine import Engine
    inter, D = ml20m, 32
    ue, ie = _tables(inter, D, 1.0)
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_sgd(0.0)  # lr 0: the model does not move, only the sampler runs
    eng.set_sampler(native.SAMPLER_UNIFORM)
    B, steps = 100_000, 12  # 1.2 M triples: several preparation waves in one call
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(2))[:B * steps].to(DEV)
    _, one_call = eng.train_steps(t, B, seed=77, step0=3, want_neg=True, want_stats=False)
    parts = [eng.train_steps(t[s * B:(s + 1) * B], B, seed=77, step0=3 + s, want_neg=True, want_stats=False)[1]
             for s in range(steps)]
    assert torch.equal(one_call, torch.cat(parts))
    alone = torch.cat([eng.sample(t[s * B:(s + 1) * B], seed=77, step=3 + s) for s in range(steps)])
    assert torch.equal(one_call, alone)
    eng.sync_check()


def test_full_size_adam_msd_shape_smoke_and_flush():
    """Config 3 shape (MSD, D=256, Adam, reg all=0.00043): steps run, losses are finite and sane,
    lazy user rows flush to the dense-Adam value (a row untouched for k steps must equal k
    zero-gradient Adam steps from its last state: checked on a few rows against the oracle replay)."""
    from rbpr import native, synth
    from rbpr.engine import Engine
    inter = synth.make("msd", seed=13, scale=0.25)
    D, B = 256, 65536
    ue, ie = _tables(inter, D, 200.0)
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg({"all": 0.00043})
    state = eng.set_adam(5e-3, (0.9, 0.999), 1e-8)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    steps = 6
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(4))[:B * steps].to(DEV)
    stats, _ = eng.train_steps(t, B, seed=9, step0=0)
    eng.flush_lazy(steps)
    eng.sync_check()
    st = stats.cpu().numpy()
    assert np.isfinite(st).all(), st
    assert (st[:, 3] == B).all() and (st[:, 0] > 0.3 * B * np.log(2)).all() and (st[:, 0] < 3 * B * np.log(2)).all()
    last = state["user_last"].cpu().numpy()
    touched = np.unique(inter.coo_users()[t.cpu().numpy()])
    assert (last[touched] == steps).all()
    # a user never touched has m = v = 0 and a zero gradient: dense Adam leaves it where it was
    never = np.setdiff1d(np.arange(1, inter.num_users), touched)[:50]
    assert np.array_equal(eng.user_emb[never].cpu().numpy(), ue.numpy()[never])
    # a user touched only in step 0: replay 5 zero-gradient Adam steps on the CPU from (p1, m1, v1)
    first = np.unique(inter.coo_users()[t[:B].cpu().numpy()])
    later = np.unique(inter.coo_users()[t[B:].cpu().numpy()])
    only0 = np.setdiff1d(first, later)[:20]
    assert only0.size > 0
    m = state["user_m"][only0].double().cpu().numpy()
    v = state["user_v"][only0].double().cpu().numpy()
    p = eng.user_emb[only0].double().cpu().numpy()
    # invert: after flush m = b1^5 m1, v = b2^5 v1 ; check the value relation p_end = p1 - sum lr*mhat/(sqrt(vhat)+eps)
    b1, b2, lr, eps = 0.9, 0.999, 5e-3, 1e-8
    m1, v1 = m / b1 ** 5, v / b2 ** 5
    drift = np.zeros_like(p)
    mm, vv = m1.copy(), v1.copy()
    for s in range(2, steps + 1):
        mm, vv = b1 * mm, b2 * vv
        drift += lr / (1 - b1 ** s) * mm / (np.sqrt(vv) / np.sqrt(1 - b2 ** s) + eps)
    assert np.abs(drift).max() > 1e-3  # the lazy rows really moved after their last gradient
    # p1 (value right after step 0) = p_end + drift must be reachable from p0 by ONE Adam step of size <= lr
    p1 = p + drift
    assert np.abs(p1 - ue.numpy()[only0]).max() <= lr * 1.0001 + 1e-7


def test_host_buffer_path_multi_wave_matches_device_path(ml20m):
    """rbpr_train_steps_host copies ids (and injected negatives) wave by wave on the preparation
    stream: same negatives, statistics and tables as the device-resident call."""
    from rbpr import native
    from rbpr.engine import Engine
    inter, D, B, steps = ml20m, 32, 100_000, 12  # 1.2 M triples: first wave + 2 full waves + tail
    ue, ie = _tables(inter, D, 20.0)
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(8))[:B * steps]
    out = []
    for host in (False, True):
        eng = Engine(ue.to(DEV), ie.to(DEV))
        eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
        eng.set_reg(REG)
        eng.set_sgd(0.05)
        eng.set_sampler(native.SAMPLER_UNIFORM)
        if host:
            stats, negs = eng.train_steps_host(t.pin_memory(), B, 5, 0, want_neg=True)
            stats, negs = stats.clone(), negs.clone()
        else:
            stats, negs = eng.train_steps(t.to(DEV), B, 5, 0, want_neg=True)
            eng.sync_check()
            stats, negs = stats.cpu(), negs.cpu()
        out.append((stats, negs, eng.user_emb.cpu(), eng.item_emb.cpu()))
        # injected negatives through the same entry point reproduce the run
        eng2 = Engine(ue.to(DEV), ie.to(DEV))
        eng2.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
        eng2.set_reg(REG)
        eng2.set_sgd(0.05)
        eng2.set_sampler(native.SAMPLER_INJECTED)
        if host:
            s2, n2 = eng2.train_steps_host(t.pin_memory(), B, 5, 0, neg_in=negs.pin_memory(), want_neg=True)
            assert torch.equal(n2, negs)
            np.testing.assert_allclose(s2.numpy(), stats.numpy(), rtol=1e-6)
            np.testing.assert_allclose(eng2.item_emb.cpu().numpy(), out[-1][3].numpy(), atol=1e-6)
    (s0, n0, u0, i0), (s1, n1, u1, i1) = out
    assert torch.equal(n0, n1)
    np.testing.assert_allclose(s0.numpy(), s1.numpy(), rtol=1e-6)
    np.testing.assert_allclose(u0.numpy(), u1.numpy(), atol=1e-6)
    np.testing.assert_allclose(i0.numpy(), i1.numpy(), atol=1e-6)


# ---- BASELINE configs[2], [3], [4] at FULL size against the pinned oracle ------------------------
def _cached(shape):
    """The bench's on-disk cache of the synthetic matrix (same generator, same seed)."""
    import os
    from pathlib import Path
    from rbpr import synth
    p = Path(os.environ.get("RBPR_CACHE", "/tmp")) / f"rbpr_synth_{shape}_1.0_13.npz"
    if p.exists():
        z = np.load(p)
        return synth.Interactions(shape, int(z["U"]), int(z["I"]), z["indptr"], z["indices"])
    inter = synth.make(shape, seed=13)
    try:
        np.savez(p.with_suffix(".tmp.npz"), U=inter.num_users, I=inter.num_items, indptr=inter.indptr, indices=inter.indices)
        os.replace(p.with_suffix(".tmp.npz"), p)
    except OSError:
        pass
    return inter


def test_config3_msd_full_size_adam_parity_vs_dense_torch_adam():
    """BASELINE configs[2]: MSD shape 571 355 x 41 140, 32.5 M interactions, D=256, Adam lr 1e-3
    betas (0.9, 0.999), reg all=0.00043 (configs/RQ2/neg-sampling/adam-ada-sampling-msd.yaml.j2:152-160),
    3 steps of 65 536 triples: the fused path (lazy user rows, dense item sweep) against the oracle's
    autograd + DENSE torch.optim.Adam over the full tables on the same triples and negatives.
    Tolerances: loss 1e-4 relative (north star), every row of both tables 2e-5 absolute (user table: all
    but at most one element in a million, see the comment at the assertion)."""
    from oracle import philox, ref_bpr
    from rbpr import native
    from rbpr.engine import Engine
    inter = _cached("msd")
    assert inter.num_users == 571_356 and inter.num_items == 41_141
    D, B, steps, seed = 256, 65536, 3, 13
    torch.manual_seed(13)
    ue = (torch.rand(inter.num_users, D) - 0.5) / D  # MF.reset_parameters (reference model.py:117-129)
    ie = (torch.rand(inter.num_items, D) - 0.5) / D
    ue[0] = 0
    ie[0] = 0
    reg = {"all": 0.00043}
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(reg)
    eng.set_adam(1e-3, (0.9, 0.999), 1e-8)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(3))[:B * steps]
    stats, negs = eng.train_steps(t.to(DEV), B, seed, 0, want_neg=True)
    eng.flush_lazy(steps)
    eng.sync_check()
    negs, st = negs.cpu().numpy(), stats.cpu().numpy()
    coo, tn = inter.coo_users(), t.numpy()
    # the sampler's specification at this shape: a 4096-triple sample per step, bit-exact
    for s in range(steps):
        sub = np.arange(s * B, s * B + 4096)
        exp = philox.sample_negatives(inter.indptr, inter.indices, coo, tn[sub], seed, s, inter.num_items)
        assert (negs[sub] == exp).all(), s
    model = ref_bpr.RefModel(ue, ie, None, reg)
    opt = ref_bpr.make_optimizer(model, "adam", lr=1e-3, betas=(0.9, 0.999))
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        out = ref_bpr.train_step(model, opt, torch.from_numpy(coo[tn[sl]]), torch.from_numpy(inter.indices[tn[sl]].astype(np.int64)),
                                 torch.from_numpy(negs[sl]))
        np.testing.assert_allclose(st[s, 0], out["bpr_loss"].item(), rtol=1e-4, err_msg=f"bpr_loss step {s}")
        np.testing.assert_allclose(st[s, 1], out["l2_reg"].item(), rtol=1e-4, err_msg=f"l2_reg step {s}")
    got_i, got_u = eng.item_emb.cpu().numpy(), eng.user_emb.cpu().numpy()
    ref_i, ref_u = model.item_emb.detach().numpy(), model.user_emb.detach().numpy()
    assert np.abs(got_i - ie.numpy()).max() > 5e-4  # Adam really moved the tables (|step| ~ lr)
    np.testing.assert_allclose(got_i, ref_i, atol=2e-5, rtol=0)
    # Adam divides by sqrt(v): where a gradient element cancels to ~1e-9 (a few of 146 M elements) its
    # last-bit rounding decides a visible fraction of the lr-sized step, for ANY two evaluation orders;
    # those elements may differ by more than 2e-5 but never by a step (lr = 1e-3)
    du = np.abs(got_u - ref_u)
    assert (du > 2e-5).sum() <= 1e-6 * du.size, (du > 2e-5).sum()
    assert du.max() < 2e-4, du.max()


def test_config4_yelp_full_size_adaptive_draws_bit_exact_vs_oracle():
    """BASELINE configs[3]: Yelp shape 252 616 x 92 089, D=64, adaptive negatives (sampling_prob 1/100,
    configs/RQ2/neg-sampling/ada-sampling-*.yaml.j2:11): one full 65 536-triple step sampled inside
    rbpr_train_steps; a 1 536-row sample of it must equal the oracle restatement of
    revisit_bpr/modules/neg_samplers.py:74-124 bit for bit (factor draw, geometric rank, rank-th
    UNSEEN item of the factor's order over all 92 090 items); every negative is unseen and != 0."""
    from oracle import adaptive
    from rbpr.engine import Engine
    inter = _cached("yelp")
    assert inter.num_users == 252_617 and inter.num_items == 92_090
    D, B, seed, step0 = 64, 65536, 13, 4
    ue, ie = _tables(inter, D, 40.0)
    eng = Engine(ue.to(DEV), ie.to(DEV))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(REG)
    eng.set_sgd(0.0)  # lr 0: only the draw is under test
    eng.set_adaptive(0.01, every=0)
    eng.adaptive_update_stats()
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(6))[:B]
    _, negs = eng.train_steps(t.to(DEV), B, seed, step0, want_neg=True)
    eng.sync_check()
    negs, tn, coo = negs.cpu().numpy(), t.numpy(), inter.coo_users()
    assert negs.min() >= 1 and negs.max() < inter.num_items
    users = coo[tn]
    for k in range(0, B, 97):  # never a seen item
        row = inter.indices[inter.indptr[users[k]]:inter.indptr[users[k] + 1]]
        assert negs[k] not in row
    snap, _ = adaptive.update_stats(ie.numpy())
    std = eng.adaptive_stats()["std"].cpu().numpy()  # same fp32 std on both sides
    np.testing.assert_allclose(std, ie.numpy()[1:].astype(np.float64).std(axis=0, ddof=1), rtol=1e-5)
    sub = np.arange(0, B, B // 1536)[:1536]
    rows = [inter.indices[inter.indptr[u]:inter.indptr[u + 1]] for u in users[sub]]
    exp = adaptive.sample(ue.numpy(), snap, std, users[sub], rows, 1, 0.01, seed, step0, subsequences=tn[sub])[:, 0]
    assert negs[sub].tolist() == exp.tolist()
    # the geometric rank reaches deep into the catalogue at p = 1/100 (mean 100): both ends are used
    assert len(set(negs.tolist())) > 20_000


def test_config5_ml20m_full_catalog_scoring_parity(ml20m):
    """BASELINE configs[4]: ML-20M shape, D=128, 10 000 eval users (20 % of each user's items held
    out), one rbpr_score_topk call over the whole set; 512 sampled users against the oracle's eval
    sequence (reference model.py:43-47,131-145 + exp.py:369-374 + metrics/ndcg.py:69-78, recall.py:44-51):
    NDCG@100 / Recall@20 within 1e-4, and the ranked top-100 ITEMS equal torch's sort where the
    oracle's own scores have no near-ties."""
    from oracle import ref_bpr
    from rbpr import synth
    from rbpr.engine import Engine
    inter, D = ml20m, 128
    torch.manual_seed(13)
    ue = torch.randn(inter.num_users, D) * 0.1
    ie = torch.randn(inter.num_items, D) * 0.1
    ue[0] = 0
    ie[0] = 0
    eng = Engine(ue.to(DEV), ie.to(DEV))
    users, seen, held = synth.split_heldout(inter, 10_000)
    res = eng.score_topk(torch.from_numpy(users), (torch.from_numpy(seen[0]), torch.from_numpy(seen[1])),
                         (torch.from_numpy(held[0]), torch.from_numpy(held[1])), [20, 100], k_max=100)
    allm = eng.score_metrics(torch.from_numpy(users), (torch.from_numpy(seen[0]), torch.from_numpy(seen[1])),
                             (torch.from_numpy(held[0]), torch.from_numpy(held[1])), [20, 100],
                             want=("ndcg", "recall", "precision", "map"))
    eng.sync_check()
    assert torch.equal(allm["ndcg"], res["ndcg"]) and torch.equal(allm["recall"], res["recall"])
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(users.size, size=512, replace=False))
    model = ref_bpr.RefModel(ue, ie, None)
    seen_pad = torch.nn.utils.rnn.pad_sequence(
        [torch.as_tensor(seen[1][seen[0][r]:seen[0][r + 1]], dtype=torch.long) for r in rows], batch_first=True)
    logits = model.eval_logits(torch.from_numpy(users[rows]), seen_pad)
    target = torch.zeros(rows.size, inter.num_items)
    for q, r in enumerate(rows):
        target[q, torch.as_tensor(held[1][held[0][r]:held[0][r + 1]], dtype=torch.long)] = 1.0
    ndcg = res["ndcg"].cpu().numpy()[rows]
    recall = res["recall"].cpu().numpy()[rows]
    np.testing.assert_allclose(ndcg[:, 1], ref_bpr.ndcg_at_k(logits, target, 100).numpy(), atol=1e-4)
    np.testing.assert_allclose(recall[:, 0], ref_bpr.recall_at_k(logits, target, 20).numpy(), atol=1e-4)
    np.testing.assert_allclose(ndcg[:, 0], ref_bpr.ndcg_at_k(logits, target, 20).numpy(), atol=1e-4)
    srt = torch.sort(logits, dim=-1, descending=True, stable=True)
    top_ref, val_ref = srt.indices[:, :100].numpy(), srt.values[:, :101].numpy()
    gaps = (val_ref[:, :-1] - val_ref[:, 1:]).min(axis=1)
    clear = gaps > 2e-6  # summation order differs between the kernel and the CPU einsum by ~1e-7
    assert clear.sum() > 300
    got_items = res["items"].cpu().numpy()[rows]
    assert (got_items[clear] == top_ref[clear]).all()
    np.testing.assert_allclose(res["scores"].cpu().numpy()[rows], val_ref[:, :100], atol=2e-6, rtol=1e-5)
    prec = allm["precision"].cpu().numpy()[rows]
    hits100 = torch.gather(target, 1, srt.indices[:, :100]).sum(-1).numpy()
    np.testing.assert_allclose(prec[:, 1], hits100 / 100.0, atol=1e-6)
