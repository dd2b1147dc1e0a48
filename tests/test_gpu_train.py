"""GPU parity tests of the fused training path, called through the C ABI (ctypes) and checked
against the oracle on the same inputs.  Tolerances: negatives bit-exact; loss 1e-4 relative
(BASELINE.json north_star), updated rows 1e-5 absolute."""
import numpy as np
import pytest
import torch

from helpers import TRAIN_CASES, load_train_case, oracle_run

REG = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}

pytestmark = pytest.mark.gpu


def _engine(case, dev):
    from rbpr import native
    from rbpr.engine import Engine
    ue = torch.as_tensor(case["init_user"]).to(dev).contiguous()
    ie = torch.as_tensor(case["init_item"]).to(dev).contiguous()
    ib = torch.as_tensor(case["init_item_bias"]).to(dev).contiguous() if case["bias"] else None
    eng = Engine(ue, ie, ib)
    eng.bind_csr(torch.as_tensor(case["indptr"]), torch.as_tensor(case["indices"]))
    eng.set_reg(case["reg"])
    kw = case["opt_kw"]
    if case["opt"] == "SGD" and kw.get("momentum", 0) == 0:
        eng.set_sgd(kw["lr"])
    elif case["opt"] == "SGD":
        eng.set_sgd_momentum(kw["lr"], kw["momentum"], kw.get("nesterov", False))
    elif case["opt"] == "RMSprop":
        eng.set_rmsprop(kw["lr"], kw.get("alpha", 0.99), kw.get("eps", 1e-8))
    else:
        eng.set_adam(kw["lr"], kw.get("betas", (0.9, 0.999)), kw.get("eps", 1e-8))
    eng.set_sampler(native.SAMPLER_INJECTED)
    return eng


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_golden_trajectory_injected_negatives(name):
    """Reference-minted trajectories: same triples + same negatives -> same losses and tables."""
    dev = torch.device("cuda:0")
    case = load_train_case(name)
    eng = _engine(case, dev)
    steps = case["triples"].shape[0]
    for s in range(steps):
        t = torch.as_tensor(case["triples"][s], dtype=torch.int64, device=dev)
        neg = torch.as_tensor(case["negs"][s], dtype=torch.int64, device=dev)
        stats, neg_out = eng.train_steps(t, case["B"], seed=1, step0=s, neg_in=neg, want_neg=True)
        eng.sync_check()
        st = stats.cpu().numpy()[0]
        assert neg_out.cpu().tolist() == case["negs"][s].tolist()
        assert st[3] == case["B"]
        np.testing.assert_allclose(st[0], case["bpr_loss"][s], rtol=1e-4)
        np.testing.assert_allclose(st[1], case["l2_reg"][s], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(st[2] / st[3], case["logits_abs_mean"][s], rtol=1e-4)
    eng.flush_lazy(steps)
    torch.cuda.synchronize()
    np.testing.assert_allclose(eng.user_emb.cpu().numpy(), case["final_user"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), case["final_item"], atol=1e-5, rtol=1e-4)
    if case["bias"]:
        np.testing.assert_allclose(eng.item_bias.cpu().numpy(), case["final_item_bias"], atol=1e-5, rtol=1e-4)
    assert eng.launch_count() >= 2 * steps


@pytest.mark.parametrize("name", ["sgd_reg3", "adam_all"])
def test_multi_step_single_call_matches_stepwise(name):
    """All steps in ONE rbpr_train_steps call (batch splitting inside the library)."""
    dev = torch.device("cuda:0")
    case = load_train_case(name)
    eng = _engine(case, dev)
    steps = case["triples"].shape[0]
    t = torch.as_tensor(case["triples"].reshape(-1), dtype=torch.int64, device=dev)
    neg = torch.as_tensor(case["negs"].reshape(-1), dtype=torch.int64, device=dev)
    stats, _ = eng.train_steps(t, case["B"], seed=1, step0=0, neg_in=neg)
    eng.flush_lazy(steps)
    eng.sync_check()
    st = stats.cpu().numpy()
    np.testing.assert_allclose(st[:, 0], case["bpr_loss"], rtol=1e-4)
    np.testing.assert_allclose(eng.user_emb.cpu().numpy(), case["final_user"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), case["final_item"], atol=1e-5, rtol=1e-4)


def test_host_buffer_entry_point():
    dev = torch.device("cuda:0")
    case = load_train_case("sgd_reg3")
    eng = _engine(case, dev)
    t = torch.as_tensor(case["triples"].reshape(-1), dtype=torch.int64).pin_memory()
    neg = torch.as_tensor(case["negs"].reshape(-1), dtype=torch.int64).pin_memory()
    stats, neg_out = eng.train_steps_host(t, case["B"], seed=1, step0=0, neg_in=neg, want_neg=True)
    np.testing.assert_allclose(stats.numpy()[:, 0], case["bpr_loss"], rtol=1e-4)
    assert neg_out.tolist() == neg.tolist()
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), case["final_item"], atol=1e-5, rtol=1e-4)


def _random_problem(U, I, D, nnz_per_user, seed, dev, bias=False):
    from rbpr import synth
    inter = synth.generate("t", U - 1, I - 1, (U - 1) * nnz_per_user, nnz_per_user, 2, 0.8, seed)
    g = torch.Generator().manual_seed(seed)
    ue = (torch.rand(U, D, generator=g) - 0.5) * 1.5
    ie = (torch.rand(I, D, generator=g) - 0.5) * 1.5
    ue[0] = 0
    ie[0] = 0
    ib = (torch.randn(I, generator=g) * 0.1) if bias else None
    if bias:
        ib[0] = 0
    return inter, ue, ie, ib


@pytest.mark.parametrize("D", [4, 8, 16, 20, 64, 96, 128, 256, 512, 1024])
def test_sampler_bit_exact_and_step_parity_on_device_sampling(D):
    """On-device sampling: negatives equal the CPU restatement of the spec bit for bit; feeding
    those negatives to the oracle reproduces loss and tables. Duplicate users and items in the
    batch are the norm here (exact minibatch semantics)."""
    from oracle import philox, ref_bpr
    from rbpr import native
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    U, I, B = 301, 203, 1024
    inter, ue, ie, ib = _random_problem(U, I, D, 12, 100 + D, dev, bias=(D == 64))
    reg = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}
    eng = Engine(ue.to(dev), ie.to(dev), None if ib is None else ib.to(dev))
    eng.bind_csr(torch.as_tensor(inter.indptr), torch.as_tensor(inter.indices))
    eng.set_reg(reg)
    eng.set_sgd(0.05)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    model = ref_bpr.RefModel(ue, ie, ib, reg)
    opt = ref_bpr.make_optimizer(model, "sgd", lr=0.05)
    coo = inter.coo_users()
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(13))
    seed = 0x1234567887654321
    for s in range(3):
        t = perm[s * B:(s + 1) * B]
        stats, negs = eng.train_steps(t.to(dev), B, seed=seed, step0=s, want_neg=True)
        eng.sync_check()
        exp = philox.sample_negatives(inter.indptr, inter.indices, coo, t.numpy(), seed, s, I)
        assert negs.cpu().numpy().tolist() == exp.tolist()
        # negatives are valid: not padding, not seen
        for k in range(0, t.numel(), 97):
            u = coo[t[k]]
            row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
            assert exp[k] > 0 and exp[k] not in row
        out = ref_bpr.train_step(model, opt, torch.as_tensor(coo[t.numpy()]),
                                 torch.as_tensor(inter.indices[t.numpy()], dtype=torch.long),
                                 torch.as_tensor(exp))
        st = stats.cpu().numpy()[0]
        np.testing.assert_allclose(st[0], out["bpr_loss"].item(), rtol=1e-4)
        np.testing.assert_allclose(st[1], out["l2_reg"].item(), rtol=1e-4)
    np.testing.assert_allclose(eng.user_emb.cpu().numpy(), model.user_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), model.item_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    if ib is not None:
        np.testing.assert_allclose(eng.item_bias.cpu().numpy(), model.item_bias.detach().numpy(), atol=1e-5, rtol=1e-4)


def test_adam_lazy_users_match_dense_adam():
    """Dense torch.optim.Adam moves every row every step; the lazy user catch-up must agree after
    flush, including users untouched for several steps."""
    from oracle import philox, ref_bpr
    from rbpr import native
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    U, I, D, B = 401, 157, 32, 256
    inter, ue, ie, ib = _random_problem(U, I, D, 8, 7, dev, bias=True)
    reg = {"all": 0.00043}
    eng = Engine(ue.to(dev), ie.to(dev), ib.to(dev))
    eng.bind_csr(torch.as_tensor(inter.indptr), torch.as_tensor(inter.indices))
    eng.set_reg(reg)
    eng.set_adam(1e-2, (0.9, 0.999), 1e-8)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    model = ref_bpr.RefModel(ue, ie, ib, reg)
    opt = ref_bpr.make_optimizer(model, "adam", lr=1e-2, betas=(0.9, 0.999))
    coo = inter.coo_users()
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(3))
    n_steps = 10
    for s in range(n_steps):
        t = perm[s * B:(s + 1) * B]
        stats, negs = eng.train_steps(t.to(dev), B, seed=5, step0=s, want_neg=True)
        exp = philox.sample_negatives(inter.indptr, inter.indices, coo, t.numpy(), 5, s, I)
        assert negs.cpu().numpy().tolist() == exp.tolist()
        out = ref_bpr.train_step(model, opt, torch.as_tensor(coo[t.numpy()]),
                                 torch.as_tensor(inter.indices[t.numpy()], dtype=torch.long),
                                 torch.as_tensor(exp))
        np.testing.assert_allclose(stats.cpu().numpy()[0, 0], out["bpr_loss"].item(), rtol=1e-4)
    eng.flush_lazy(n_steps)
    eng.sync_check()
    np.testing.assert_allclose(eng.user_emb.cpu().numpy(), model.user_emb.detach().numpy(), atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_emb.cpu().numpy(), model.item_emb.detach().numpy(), atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(eng.item_bias.cpu().numpy(), model.item_bias.detach().numpy(), atol=2e-5, rtol=1e-4)


def test_weighted_sampler_bit_exact_and_distribution():
    from oracle import philox
    from rbpr import native
    from rbpr.engine import Engine, build_alias
    dev = torch.device("cuda:0")
    U, I, D = 120, 64, 16
    inter, ue, ie, _ = _random_problem(U, I, D, 6, 21, dev)
    counts = np.bincount(inter.indices, minlength=I).astype(np.float64)
    w = counts ** 0.75
    w[0] = 0
    eng = Engine(ue.to(dev), ie.to(dev))
    eng.bind_csr(torch.as_tensor(inter.indptr), torch.as_tensor(inter.indices))
    eng.bind_item_weights(torch.as_tensor(w))
    prob, alias = build_alias(w.copy())
    coo = inter.coo_users()
    t = torch.arange(inter.nnz)
    got = eng.sample(t.to(dev), seed=77, step=3, sampler=native.SAMPLER_WEIGHTED).cpu().numpy()
    exp = philox.sample_negatives(inter.indptr, inter.indices, coo, t.numpy(), 77, 3, I, alias=(prob, alias))
    assert got.tolist() == exp.tolist()
    # distribution: one user, many steps -> frequencies ∝ w over unseen items
    u = int(np.argmax(np.diff(inter.indptr)))
    tt = torch.arange(int(inter.indptr[u]), int(inter.indptr[u + 1]), device=dev)
    counts = np.zeros(I)
    for s in range(600):
        d = eng.sample(tt, seed=1, step=s, sampler=native.SAMPLER_WEIGHTED).cpu().numpy()
        np.add.at(counts, d, 1)
    row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
    allowed = np.setdiff1d(np.nonzero(w > 0)[0], row)
    assert counts[row].sum() == 0 and counts[0] == 0 and counts[w <= 0].sum() == 0
    expc = counts.sum() * w[allowed] / w[allowed].sum()
    chi2 = ((counts[allowed] - expc) ** 2 / expc).sum()
    dof = allowed.size - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)


def test_uniform_sampler_distribution_chi2():
    """Distributional equivalence with the reference's UniformSampler: uniform over the unseen,
    non-padding items of the user (neg_samplers.py:135-141)."""
    from rbpr import native
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    U, I, D = 50, 41, 16
    inter, ue, ie, _ = _random_problem(U, I, D, 10, 5, dev)
    eng = Engine(ue.to(dev), ie.to(dev))
    eng.bind_csr(torch.as_tensor(inter.indptr), torch.as_tensor(inter.indices))
    u = int(np.argmax(np.diff(inter.indptr)))
    tt = torch.arange(int(inter.indptr[u]), int(inter.indptr[u + 1]), device=dev)
    counts = np.zeros(I)
    for s in range(600):  # every (step, triple) pair is an independent draw for the same user
        d = eng.sample(tt, seed=2024, step=s).cpu().numpy()
        np.add.at(counts, d, 1)
    row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
    allowed = np.setdiff1d(np.arange(1, I), row)
    assert counts[row].sum() == 0 and counts[0] == 0
    tot = counts.sum()
    expc = tot / allowed.size
    chi2 = ((counts[allowed] - expc) ** 2 / expc).sum()
    # dof = allowed.size-1 (~30): mean 30, sd ~7.7; 5 sd bound
    assert chi2 < (allowed.size - 1) + 5 * np.sqrt(2 * (allowed.size - 1)), chi2


def test_error_paths():
    from rbpr import native
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    ue = torch.zeros(5, 8, device=dev)
    ie = torch.zeros(4, 8, device=dev)
    eng = Engine(ue, ie)
    # a user who has seen every non-padding item: the reference's multinomial raises (all-zero row)
    indptr = torch.tensor([0, 0, 3, 3, 3, 3])
    indices = torch.tensor([1, 2, 3], dtype=torch.int32)
    with pytest.raises(native.NativeError, match="seen every"):
        eng.bind_csr(indptr, indices)
    # unsorted row
    with pytest.raises(native.NativeError, match="ascending"):
        eng.bind_csr(torch.tensor([0, 0, 2, 2, 2, 2]), torch.tensor([2, 1], dtype=torch.int32))
    eng.bind_csr(torch.tensor([0, 0, 2, 2, 2, 2]), torch.tensor([1, 2], dtype=torch.int32))
    eng.set_sgd(0.1)
    # triple index out of range
    eng.train_steps(torch.tensor([5], device=dev), 1, 0, 0)
    with pytest.raises(native.NativeError, match="out of range"):
        eng.sync_check()
    # empty batch is a no-op
    eng.train_steps(torch.zeros(0, dtype=torch.int64, device=dev), 4, 0, 0)
    eng.sync_check()
    # Adam without state
    eng.hp.optimizer = native.OPT_ADAM
    with pytest.raises(native.NativeError, match="Adam"):
        eng.train_steps(torch.tensor([0], device=dev), 1, 0, 0)
    with pytest.raises(native.NativeError):
        Engine(torch.zeros(5, 6, device=dev), torch.zeros(4, 6, device=dev))  # dim % 4 != 0


@pytest.mark.parametrize("D,B,bias", [(128, 256, False), (64, 256, True), (20, 100, False), (256, 512, False)])
def test_small_batch_cluster_kernel_matches_large_batch_path_and_oracle(D, B, bias, monkeypatch):
    """Small batches (the reference configs' train_batch_size 256) run whole waves of steps inside ONE
    launch of a thread-block cluster (csrc/train_small.cu).  Same negatives, statistics and tables as
    the two-launches-per-step path, and as the oracle's autograd + SGD, on a problem dense in repeated
    users and hot items (duplicates inside a step are the hard part of the exact minibatch)."""
    from oracle import ref_bpr
    from rbpr import native
    from rbpr.engine import Engine
    dev = torch.device("cuda:0")
    inter, ue, ie, ib = _random_problem(1500, 300, D, 12, 5 + D, dev, bias=bias)
    steps = min(37, inter.nnz // B)
    t = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(1))[:B * steps]
    assert t.numel() == B * steps
    runs = []
    for small in (True, False):
        if small:
            monkeypatch.delenv("RBPR_NO_SMALL_BATCH", raising=False)
        else:
            monkeypatch.setenv("RBPR_NO_SMALL_BATCH", "1")
        eng = Engine(ue.to(dev), ie.to(dev), None if ib is None else ib.to(dev))
        eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
        eng.set_reg(REG)
        eng.set_sgd(0.05)
        eng.set_sampler(native.SAMPLER_UNIFORM)
        l0 = eng.launch_count()
        stats, negs = eng.train_steps(t.to(dev), B, seed=3, step0=11, want_neg=True)
        eng.sync_check()
        runs.append((stats.cpu().numpy(), negs.cpu().numpy(), eng.user_emb.cpu().numpy(), eng.item_emb.cpu().numpy(),
                     None if ib is None else eng.item_bias.cpu().numpy(), eng.launch_count() - l0))
        # a second call on the same engine continues correctly (per-item epoch stamps keep advancing)
        if small:
            st2, _ = eng.train_steps(t[:B * 3].to(dev), B, seed=3, step0=11 + steps)
            eng.sync_check()
            assert np.isfinite(st2.cpu().numpy()).all() and (st2[:, 3] == B).all().item()
    (s0, n0, u0, i0, b0, l_small), (s1, n1, u1, i1, b1, l_big) = runs
    assert l_small < 8 and l_big >= 2 * steps  # one cluster launch per wave vs two launches per step
    assert (n0 == n1).all()
    np.testing.assert_allclose(s0, s1, rtol=2e-6)
    np.testing.assert_allclose(u0, u1, atol=2e-6)
    np.testing.assert_allclose(i0, i1, atol=2e-6)
    model = ref_bpr.RefModel(ue, ie, ib, REG)
    opt = ref_bpr.make_optimizer(model, "sgd", lr=0.05)
    coo, tn = inter.coo_users(), t.numpy()
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        out = ref_bpr.train_step(model, opt, torch.from_numpy(coo[tn[sl]]), torch.from_numpy(inter.indices[tn[sl]].astype(np.int64)),
                                 torch.from_numpy(n0[sl]))
        np.testing.assert_allclose(s0[s, 0], out["bpr_loss"].item(), rtol=1e-4)
        np.testing.assert_allclose(s0[s, 1], out["l2_reg"].item(), rtol=1e-4)
    np.testing.assert_allclose(u0, model.user_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(i0, model.item_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    if bias:
        np.testing.assert_allclose(b0, model.item_bias.detach().numpy(), atol=1e-5, rtol=1e-4)
        np.testing.assert_allclose(b0, b1, atol=2e-6)
