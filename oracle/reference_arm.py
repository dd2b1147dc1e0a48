"""The reference's OWN CPU implementation of the hot path, timed.  TEST / BASELINE INFRASTRUCTURE ONLY
(bench.py's `--impl reference` arm and its `cpu_baseline` leg; never imported by the product).

`__graft_entry__.build()` installs the reference's library package (`revisit_bpr`, unmodified) into
the git-ignored `oracle/_ref/` when /root/reference exists — the file copy `pip install --target`
would make; it travels to the GPU box with the snapshot, /root/reference does not.  When it is
there, the loops below drive the reference's own classes (`kind = "reference"`):
    UniformSampler.sample -> Model.forward -> loss.backward() -> torch.optim step -> zero_grad
(reference example.py:172-180 / experiments/trainer.py:64-83) and the eval sequence MF eval forward
-> scatter mask (experiments/bpr/exp.py:369-374) -> NDCG / Recall (example.py:209-221).  Otherwise the
same loops run on the oracle restatement oracle/ref_bpr.py (`kind = "port"`).

Must run in a process that has NOT imported this repo's drop-in `revisit_bpr` package (same name).
"""
from __future__ import annotations

import os
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

REF_DIR = Path(__file__).resolve().parent / "_ref"


def load_reference():
    """Import the installed reference package; None when oracle/_ref is absent."""
    if not (REF_DIR / "revisit_bpr" / "__init__.py").exists():
        return None
    if "revisit_bpr" in sys.modules and not str(getattr(sys.modules["revisit_bpr"], "__file__", "")).startswith(str(REF_DIR)):
        raise RuntimeError("the drop-in revisit_bpr is already imported in this process: run the reference arm "
                           "in its own process")
    if "accelerate" not in sys.modules:  # revisit_bpr/metrics/metric.py:5 imports it for a type annotation only
        try:
            import accelerate  # noqa: F401
        except ImportError:
            stub = types.ModuleType("accelerate")
            stub.Accelerator = type("Accelerator", (), {})
            sys.modules["accelerate"] = stub
    sys.path.insert(0, str(REF_DIR))
    import revisit_bpr  # noqa: F401
    from revisit_bpr.metrics import NDCG, Recall
    from revisit_bpr.models.bpr import MF, Model
    from revisit_bpr.modules import AdaptiveSampler, UniformSampler
    assert Path(revisit_bpr.__file__).resolve().is_relative_to(REF_DIR)
    return types.SimpleNamespace(Model=Model, MF=MF, UniformSampler=UniformSampler, AdaptiveSampler=AdaptiveSampler,
                                 NDCG=NDCG, Recall=Recall)


def _padded_seen(indptr: np.ndarray, indices: np.ndarray, users: np.ndarray) -> torch.Tensor:
    lens = indptr[users + 1] - indptr[users]
    width = max(1, int(lens.max()))
    out = np.zeros((users.size, width), dtype=np.int64)
    for r, u in enumerate(users):
        out[r, :lens[r]] = indices[indptr[u]:indptr[u + 1]]
    return torch.from_numpy(out)


def train_throughput(indptr, indices, coo_users, num_users: int, num_items: int, dim: int, batch: int, steps: int,
                     warmup: int, opt: str = "sgd", lr: float = 1e-3, reg: dict | None = None, seed: int = 13,
                     sampler: str = "uniform", adaptive_prob: float = 0.01) -> dict:
    """triples/s of the reference's training step on this host's cores.  One step = one batch of
    `batch` triples: sampler -> forward -> backward -> optimizer step (dense gradients, dense update)."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    torch.manual_seed(seed)
    nnz = int(indptr[-1])
    perm = torch.randperm(nnz, generator=torch.Generator().manual_seed(seed)).numpy()
    gen = torch.Generator().manual_seed(seed)
    if ref is not None:
        kind = "reference"
        model = ref.Model(ref.MF(torch.nn.Embedding(num_users, dim, padding_idx=0),
                                 torch.nn.Embedding(num_items, dim, padding_idx=0)), reg_alphas=reg, fuse_forward=True)
        model.train()
        params = list(model.parameters())
        smp = (ref.AdaptiveSampler(model, num_items, adaptive_prob, gen, every=10 ** 9) if sampler == "adaptive"
               else ref.UniformSampler(num_items, gen))
        if sampler == "adaptive":
            smp.update_stats()
    else:
        kind = "port"
        from oracle import ref_bpr
        ue = (torch.rand(num_users, dim) - 0.5) / dim
        ie = (torch.rand(num_items, dim) - 0.5) / dim
        ue[0] = 0
        ie[0] = 0
        model = ref_bpr.RefModel(ue, ie, None, reg)
        params = model.parameters()
        weights = torch.ones(num_items)
    optim = (torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999)) if opt == "adam" else torch.optim.SGD(params, lr=lr))
    times = []
    for s in range(warmup + steps):
        off = (s * batch) % max(1, nnz - batch)
        t = perm[off:off + batch]
        users_np = coo_users[t].astype(np.int64)
        users = torch.from_numpy(users_np)
        items = torch.from_numpy(indices[t].astype(np.int64))
        seen = _padded_seen(indptr, indices, users_np)  # a precomputed matrix in the reference (dataset.py:157-181): not timed
        t0 = time.perf_counter()
        if ref is not None:
            b = {"user": users, "item": items.unsqueeze(-1), "seen_items": seen}
            b["neg"] = smp.sample(b)
            out = model(b)
            out["loss"].backward()
            optim.step()
            optim.zero_grad()
        else:
            neg = ref_bpr.reference_style_negatives(weights, seen, gen)
            ref_bpr.train_step(model, optim, users, items, neg)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "kind": kind,
            "cores": torch.get_num_threads(), "steps": len(times), "batch": batch}


def eval_throughput(users: np.ndarray, seen: tuple, held: tuple, num_users: int, num_items: int, dim: int,
                    batch: int = 128, max_users: int = 512, seed: int = 13) -> dict:
    """users/s of the reference's eval sequence on this host's cores: all-item logits for `batch`
    users at a time, seen mask, NDCG@100 and Recall@20 (BASELINE configs[4])."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    torch.manual_seed(seed)
    n = min(max_users, users.size)
    items = torch.arange(num_items).unsqueeze(0)
    if ref is not None:
        kind = "reference"
        model = ref.Model(ref.MF(torch.nn.Embedding(num_users, dim, padding_idx=0),
                                 torch.nn.Embedding(num_items, dim, padding_idx=0)))
        model.eval()
        ndcg, recall = ref.NDCG(100), ref.Recall(20)
    else:
        kind = "port"
        from oracle import ref_bpr
        ue = (torch.rand(num_users, dim) - 0.5) / dim
        ie = (torch.rand(num_items, dim) - 0.5) / dim
        model = ref_bpr.RefModel(ue, ie, None, None)
    t_total = 0.0
    for a in range(0, n, batch):
        rows = np.arange(a, min(a + batch, n))
        u = torch.from_numpy(users[rows].astype(np.int64))
        seen_pad = _padded_seen(seen[0], seen[1], rows)
        target = torch.zeros(rows.size, num_items)
        for r, q in enumerate(rows):
            target[r, torch.from_numpy(held[1][held[0][q]:held[0][q + 1]].astype(np.int64))] = 1.0
        t0 = time.perf_counter()
        with torch.no_grad():
            if ref is not None:
                logits = model({"user": u, "item": items.repeat(rows.size, 1)})["logits"]
                logits.scatter_(-1, seen_pad, -1e13)  # exp.py:369-374
                logits[:, 0] = -1e13
                ndcg(logits, target)
                recall(logits, target)
            else:
                logits = model.eval_logits(u, seen_pad)
                ref_bpr.ndcg_at_k(logits, target, 100)
                ref_bpr.recall_at_k(logits, target, 20)
        t_total += time.perf_counter() - t0
    return {"value": n / t_total, "kind": kind, "cores": torch.get_num_threads(), "users": int(n), "batch": batch}
