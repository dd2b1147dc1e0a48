"""Correctness check of the in-library data-parallel step (one process per GPU, inside an initialised
torch.distributed NCCL group): W ranks x their share of a global batch == the oracle (autograd +
dense torch.optim on ONE process) on the whole batch, for SGD and for Adam; item replicas
bit-identical across ranks; user rows exact on their owner.  Used by tests/tools/check_multi_gpu.py
and by bench.py (N>1) as the `parity_check` that precedes the timed region.  CHECKER ONLY."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def run(dev: torch.device, rank: int, world: int, optimizer: str = "sgd", setup=None) -> tuple[bool, str]:
    from oracle import philox, ref_bpr
    from rbpr import native, synth
    from rbpr.engine import Engine
    from rbpr.parallel import owned_triples
    inter = synth.generate("t", 4000, 900, 60000, 12, 4, 0.9, 7)
    D, B, steps, seed = 32, 8192, 4, 321
    reg = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}
    torch.manual_seed(1)
    ue = (torch.rand(inter.num_users, D) - 0.5) * 1.2
    ie = (torch.rand(inter.num_items, D) - 0.5) * 1.2
    ue[0] = 0
    ie[0] = 0
    eng = Engine(ue.to(dev), ie.to(dev))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(reg)
    model = ref_bpr.RefModel(ue, ie, None, reg)
    if optimizer == "adam":
        eng.set_adam(5e-3, (0.9, 0.999), 1e-8)
        opt = ref_bpr.make_optimizer(model, "adam", lr=5e-3, betas=(0.9, 0.999))
    else:
        eng.set_sgd(0.05)
        opt = ref_bpr.make_optimizer(model, "sgd", lr=0.05)
    eng.set_sampler(native.SAMPLER_UNIFORM)
    eng.init_comm()
    if setup is not None:
        setup(eng)
    lo, hi = owned_triples(inter.indptr, world, rank)
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(3)).numpy()
    coo = inter.coo_users()
    ok, why = True, []
    for s in range(steps):
        t = perm[s * B:(s + 1) * B]
        own = (t >= lo) & (t < hi)
        mine = t[own]
        stats, negs = eng.train_steps(torch.as_tensor(mine).to(dev), max(len(mine), 1), seed, s, want_neg=True)
        eng.sync_check()
        tot = stats.clone()
        dist.all_reduce(tot)
        # negatives are a function of (seed, step, triple): rank-independent
        exp_neg = philox.sample_negatives(inter.indptr, inter.indices, coo, t, seed, s, inter.num_items)
        if not (negs.cpu().numpy() == exp_neg[own]).all():
            ok = False
            why.append(f"negatives step {s}")
        out = ref_bpr.train_step(model, opt, torch.from_numpy(coo[t]), torch.from_numpy(inter.indices[t].astype(np.int64)),
                                 torch.from_numpy(exp_neg))
        bpr = out["bpr_loss"].item()
        if not (abs(tot[0, 0].item() - bpr) <= 1e-4 * abs(bpr) and tot[0, 3].item() == B):
            ok = False
            why.append(f"loss step {s}: {tot[0, 0].item()} vs {bpr}")
    eng.flush_lazy(steps)
    eng.sync_check()
    item = eng.item_emb.clone()
    gathered = [torch.empty_like(item) for _ in range(world)]
    dist.all_gather(gathered, item)
    if not all(torch.equal(gathered[0], g) for g in gathered):
        ok = False
        why.append("item replicas differ across ranks")
    if not np.allclose(item.cpu().numpy(), model.item_emb.detach().numpy(), atol=1e-5, rtol=1e-4):
        ok = False
        why.append(f"item table vs oracle: {np.abs(item.cpu().numpy() - model.item_emb.detach().numpy()).max()}")
    cuts = np.searchsorted(inter.indptr, [lo, hi], side="left")
    rows = np.arange(cuts[0], cuts[1])
    if rows.size and not np.allclose(eng.user_emb.cpu().numpy()[rows], model.user_emb.detach().numpy()[rows],
                                     atol=1e-5, rtol=1e-4):
        ok = False
        why.append("owned user rows vs oracle")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    eng.close()
    del eng
    return flag.item() == 1, "; ".join(why)
