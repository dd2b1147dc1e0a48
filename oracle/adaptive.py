"""CPU restatement of the adaptive negative sampler.  TEST ORACLE ONLY.

Two layers:
  * `reference_pick` restates the DETERMINISTIC part of the reference exactly as it is written
    (revisit_bpr/modules/neg_samplers.py:90-121 / experiments/bpr/exp.py:311-342): given the factor
    and the geometric draw, clamp, choose top/bottom by the sign of u_f, mask seen items and item
    0 to -1e13, argsort descending, take the rank.  Pinned against the real reference by
    tests/golden/adaptive.npz (minted by tests/golden/make_golden.py).
  * `sample` restates OUR counter-based draws (factor by a blocked fp32 inverse CDF from Philox
    word 0, geometric rank from word 1 in fp64) and then calls `reference_pick`, so equality with
    the CUDA kernel checks both the stream and the kernel's rank-skip search.
"""
from __future__ import annotations

import math

import numpy as np

from oracle.philox import philox4x32_10


def update_stats(item_emb: np.ndarray):
    """neg_samplers.py:126-132: snapshot (D,I) = item table transposed; unbiased std over items[1:]."""
    snap = np.ascontiguousarray(item_emb.T).astype(np.float32)
    std = item_emb[1:].astype(np.float64).std(axis=0, ddof=1).astype(np.float32)
    return snap, std


def reference_pick(snap_row: np.ndarray, seen: np.ndarray, u_f: float, geom: int) -> int:
    num_items = snap_row.size
    banned = np.unique(np.concatenate([seen[seen > 0], [0]])).astype(np.int64)
    n_unseen = num_items - banned.size
    rank = min(int(geom), n_unseen)                      # .clamp_(max=num_notseen_items)
    rank = rank - 1 if u_f > 0 else n_unseen - rank      # torch.where(u_f.gt(0), rank-1, n-rank)
    vals = snap_row.astype(np.float32).copy()
    vals[banned] = -1e13                                 # scatter(seen ∪ {0}, -1e13)
    order = np.argsort(-vals, kind="stable")             # ties: lower item id first (our rule)
    return int(order[rank])


def sample(user_emb: np.ndarray, snap: np.ndarray, std: np.ndarray, users, seen_rows, num: int,
           p: float, seed: int, step: int, subsequences=None) -> np.ndarray:
    """seen_rows: list of 1-D int arrays (0 entries ignored).  Returns (B,num) int64."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    off = int(step) << 8
    log1m = math.log1p(-p)
    out = np.zeros((len(users), num), dtype=np.int64)
    for row, u in enumerate(users):
        urow = user_emb[u].astype(np.float32)
        for s in range(num):
            slot = row * num + s if subsequences is None else int(subsequences[row])
            w = [int(x) for x in philox4x32_10(np.uint32(off & 0xFFFFFFFF), np.uint32((off >> 32) & 0xFFFFFFFF),
                                               np.uint32(slot & 0xFFFFFFFF), np.uint32(slot >> 32), k0, k1)]
            weights = np.abs(urow) * std                  # float32 products
            # blocked fp32 inverse CDF (8 blocks of `blk` consecutive factors): block sums s_l in
            # sequential order, prefix P_l accumulated in block order, then the first factor whose
            # running sum fl(fl(P_l + w_a) + w_b ...) exceeds the target
            D = weights.size
            blk = ((((D + 7) // 8) + 3) // 4) * 4
            sums = []
            for l in range(8):
                acc = np.float32(0)
                for x in weights[l * blk:(l + 1) * blk]:
                    acc = np.float32(acc + x)
                sums.append(acc)
            prefix, total = [], np.float32(0)
            for l in range(8):
                prefix.append(total)
                total = np.float32(total + sums[l])
            if not total > 0:
                raise RuntimeError("invalid multinomial distribution (sum of probabilities <= 0)")
            target = np.float32(np.float32(w[0] >> 8) * np.float32(1.0 / 16777216.0)) * total
            factor, last_pos = -1, 0
            for l in range(8):
                cum = prefix[l]
                for f in range(l * blk, min(D, (l + 1) * blk)):
                    x = weights[f]
                    cum = np.float32(cum + x)
                    if x > 0:
                        last_pos = f
                    if factor < 0 and cum > target:
                        factor = f
            if factor < 0:
                factor = last_pos
            u2 = (float(w[1] >> 8) + 1.0) * (1.0 / 16777216.0)
            geom = max(1, math.ceil(math.log(u2) / log1m))
            out[row, s] = reference_pick(snap[factor], np.asarray(seen_rows[row]), float(urow[factor]), geom)
    return out
