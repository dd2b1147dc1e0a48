"""N>1 host logic on CPU with gloo (world_size 2): user-owner sharding, one all-reduce of the dense
item gradient per step, identical item update on every rank — equal to the single-process oracle on
the concatenated batch (the loss is a sum, so W shards of B ≡ one batch of W·B; SURVEY §8 e).
The per-rank arithmetic is supplied by an oracle-backed stand-in with the Engine's step interface
(closed-form minibatch gradients); the CUDA engine itself is covered by the -m gpu tests
and by bench.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_train_case


class OracleStepEngine:
    """rbpr.engine.Engine's data-parallel interface, evaluated in numpy float64 (tests only)."""

    def __init__(self, case, lr):
        self.U, self.I, self.D = case["U"], case["I"], case["D"]
        self.user = case["init_user"].astype(np.float64).copy()
        self.item = case["init_item"].astype(np.float64).copy()
        self.coo, self.indices = case["coo_user"], case["indices"]
        from oracle.ref_bpr import resolve_reg
        self.reg = resolve_reg(case["reg"])
        self.lr = lr
        self._grad = torch.zeros(self.I * self.D, dtype=torch.float64)

    def item_grad_tensor(self):
        return self._grad

    def grad_step(self, triple_idx, seed, step, neg_in=None, want_neg=False):
        t = triple_idx.numpy()
        u, i, j = self.coo[t], self.indices[t], neg_in.numpy()
        ru, ri, rn = self.reg
        g = self._grad.view(self.I, self.D).numpy()
        U0, V0 = self.user.copy(), self.item
        x = (U0[u] * (V0[i] - V0[j])).sum(1)
        c = 1.0 / (1.0 + np.exp(x))  # sigmoid(-x)
        np.add.at(g, i, -c[:, None] * U0[u] + ri * V0[i])
        np.add.at(g, j, c[:, None] * U0[u] + rn * V0[j])
        gu = np.zeros_like(self.user)
        np.add.at(gu, u, -c[:, None] * (V0[i] - V0[j]) + ru * U0[u])
        self.user -= self.lr * gu  # users are owned by this rank: local update, no communication
        loss = np.logaddexp(0.0, -x).sum()
        return torch.tensor([[loss, 0.0, np.abs(x).sum(), float(len(t))]], dtype=torch.float64), None

    def apply_item_grads(self, step):
        self.item -= self.lr * self._grad.view(self.I, self.D).numpy()
        self._grad.zero_()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rbpr.parallel import DataParallelTrainer, owned_triples
        case = load_train_case("sgd_reg3")
        lr = case["opt_kw"]["lr"]
        eng = OracleStepEngine(case, lr)
        dp = DataParallelTrainer(eng)
        lo, hi = owned_triples(case["indptr"], world, rank)
        losses = []
        for s in range(case["triples"].shape[0]):
            t, neg = case["triples"][s], case["negs"][s]
            mine = (t >= lo) & (t < hi)  # this rank's share of the GLOBAL batch: triples of its users
            stats, _ = dp.step(torch.as_tensor(t[mine]), 1, s, neg_in=torch.as_tensor(neg[mine]))
            losses.append(dp.reduce_stats(stats)[0, 0].item())
        # user rows live on their owner; gather them for the check
        cuts = np.searchsorted(case["indptr"], [lo, hi], side="left")
        owned = torch.zeros(case["U"], case["D"], dtype=torch.float64)
        rows = np.unique(case["coo_user"][lo:hi])
        owned[rows] = torch.as_tensor(eng.user[rows])
        dist.all_reduce(owned)
        untouched = torch.as_tensor(case["init_user"], dtype=torch.float64)
        all_rows = np.unique(case["coo_user"])
        untouched[all_rows] = owned[all_rows]
        if rank == 0:
            ret["losses"] = losses
            ret["user"] = untouched.numpy()
            ret["item"] = eng.item.copy()
        ret[f"item_{rank}"] = eng.item.copy()
        del cuts
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_equal_the_single_process_reference_trajectory():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    case = load_train_case("sgd_reg3")
    np.testing.assert_allclose(ret["losses"], case["bpr_loss"], rtol=1e-5)
    np.testing.assert_allclose(ret["user"], case["final_user"], atol=2e-6)
    np.testing.assert_allclose(ret["item"], case["final_item"], atol=2e-6)
    np.testing.assert_array_equal(ret["item_0"], ret["item_1"])  # replicas stay identical


def test_owner_sharding_partitions_all_triples():
    from rbpr import synth
    from rbpr.parallel import owned_triples, shard_bounds
    inter = synth.make("c1-small", seed=13)
    for world in (1, 2, 3, 8):
        cuts = shard_bounds(inter.indptr, world)
        assert cuts[0] == 0 and cuts[-1] == inter.num_users and (np.diff(cuts) >= 0).all()
        spans = [owned_triples(inter.indptr, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == inter.nnz
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
        sizes = np.array([b - a for a, b in spans])
        assert sizes.max() - sizes.min() <= np.diff(inter.indptr).max()  # balanced up to one user row
        # a user's triples never straddle two ranks
        users = inter.coo_users()
        for a, b in spans[:-1]:
            if 0 < b < inter.nnz:
                assert users[b - 1] != users[b]
