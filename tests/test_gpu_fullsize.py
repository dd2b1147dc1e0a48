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
