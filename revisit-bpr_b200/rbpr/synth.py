"""Deterministic synthetic interaction matrices with the shapes BASELINE.json names.

Shapes follow the reference README table (README.md:51-57): user degrees are log-normal with
the published median, items are drawn from a Zipf-like popularity without repetition per
user, ids are shifted by +1 so that row 0 is the padding row of both tables
(bin/datasets/format-repro.sh:56-63).  Nothing here is on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

# name -> (users, items, nnz, median user degree, min degree, zipf exponent)
SHAPES = {
    "c1-small": (1_000, 500, 20_000, 18, 5, 0.8),
    "ml-20m": (136_677, 20_108, 9_700_000, 37, 5, 0.9),
    "msd": (571_355, 41_140, 32_500_000, 39, 20, 0.6),
    "yelp": (252_616, 92_089, 2_200_000, 5, 5, 0.9),
    "netflix-rq1": (9_949, 4_825, 563_577, 30, 5, 0.8),
}


@dataclass
class Interactions:
    name: str
    num_users: int  # table rows, including the padding row 0
    num_items: int
    indptr: np.ndarray  # (num_users+1,) int64; row 0 is empty
    indices: np.ndarray  # (nnz,) int32 ascending within a row, values in [1, num_items)

    @property
    def nnz(self) -> int:
        return int(self.indices.size)

    def coo_users(self) -> np.ndarray:
        return np.repeat(np.arange(self.num_users, dtype=np.int64), np.diff(self.indptr))


def make(name: str, seed: int = 13, scale: float = 1.0) -> Interactions:
    users0, items0, nnz0, median, min_deg, zipf = SHAPES[name]
    if scale != 1.0:
        users0 = max(8, int(users0 * scale))
        nnz0 = max(users0 * min_deg, int(nnz0 * scale))
    return generate(name, users0, items0, nnz0, median, min_deg, zipf, seed)


def generate(name: str, users0: int, items0: int, nnz0: int, median: float, min_deg: int,
             zipf: float, seed: int = 13) -> Interactions:
    rng = np.random.default_rng(seed)
    max_deg = min(2048, items0 // 2)
    # log-normal degrees: median fixed, sigma solved (bisection) so the clipped mean hits nnz0/users0
    target_mean = nnz0 / users0
    z = rng.standard_normal(users0)
    lo, hi = 0.01, 3.0
    for _ in range(40):
        sig = 0.5 * (lo + hi)
        deg = np.clip(np.exp(np.log(median) + sig * z), min_deg, max_deg)
        if deg.mean() < target_mean:
            lo = sig
        else:
            hi = sig
    deg = np.clip(np.rint(deg), min_deg, max_deg).astype(np.int64)
    # item popularity ~ rank^-zipf over a random relabelling of items
    pop = (np.arange(1, items0 + 1, dtype=np.float64)) ** (-zipf)
    pop /= pop.sum()
    cdf = np.cumsum(pop)
    relabel = rng.permutation(items0).astype(np.int64) + 1  # ids 1..items0
    # oversample with replacement, dedupe per user, keep up to deg[u] per user
    over = 1.6
    want = np.ceil(deg * over).astype(np.int64) + 4
    owner = np.repeat(np.arange(users0, dtype=np.int64), want)
    draws = np.searchsorted(cdf, rng.random(owner.size), side="right").clip(0, items0 - 1)
    key = owner * np.int64(items0 + 1) + relabel[draws]
    # stable unique keeping first-draw order inside a user is not needed: take unique pairs,
    # then a random subset of deg[u] per user (random priority), then sort by item.
    # (torch's multi-threaded sorts: same results as np.unique / np.lexsort / np.sort, 20x faster at
    # the MSD shape, where the bench generates 54 M candidate pairs)
    key = torch.unique(torch.from_numpy(key)).numpy()
    owner_u = key // np.int64(items0 + 1)
    prio = rng.random(key.size)
    order = torch.argsort(torch.from_numpy(owner_u.astype(np.float64) + prio), stable=True).numpy()
    owner_s = owner_u[order]
    key_s = key[order]
    start = np.searchsorted(owner_s, np.arange(users0), side="left")
    rank_in_user = np.arange(key_s.size) - start[owner_s]
    keep = rank_in_user < deg[owner_s]
    key_k = torch.sort(torch.from_numpy(key_s[keep])).values.numpy()  # sort by (user, item)
    users_k = key_k // np.int64(items0 + 1)
    items_k = (key_k % np.int64(items0 + 1)).astype(np.int32)
    counts = np.bincount(users_k, minlength=users0)
    indptr = np.zeros(users0 + 2, dtype=np.int64)  # +1 padding row, +1 for indptr
    indptr[2:] = np.cumsum(counts)
    return Interactions(name=name, num_users=users0 + 1, num_items=items0 + 1, indptr=indptr,
                        indices=items_k)


def split_heldout(inter: Interactions, n_eval_users: int, frac: float = 0.2, seed: int = 13):
    """Fold-in style split for the scoring config (experiments/datasets/revisit-ials/
    generate_data.py:65-104): choose n_eval_users users, hold out `frac` of each user's items.
    Returns (users int64, seen CSR (indptr,indices), held CSR (indptr,indices)) in local rows."""
    rng = np.random.default_rng(seed + 1)
    deg = np.diff(inter.indptr)
    cand = np.nonzero(deg >= 5)[0]
    users = np.sort(rng.choice(cand, size=min(n_eval_users, cand.size), replace=False)).astype(np.int64)
    seen_ptr, held_ptr = [0], [0]
    seen_idx, held_idx = [], []
    for u in users:
        row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
        n_held = max(1, int(round(frac * row.size)))
        mask = np.zeros(row.size, dtype=bool)
        mask[rng.choice(row.size, size=n_held, replace=False)] = True
        held_idx.append(row[mask])
        seen_idx.append(row[~mask])
        held_ptr.append(held_ptr[-1] + n_held)
        seen_ptr.append(seen_ptr[-1] + row.size - n_held)
    return (users,
            (np.asarray(seen_ptr, dtype=np.int64), np.concatenate(seen_idx).astype(np.int32)),
            (np.asarray(held_ptr, dtype=np.int64), np.concatenate(held_idx).astype(np.int32)))
