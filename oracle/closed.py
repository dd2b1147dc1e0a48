"""numpy closed-form restatement of one synchronous BPR minibatch step (SGD).  TEST ORACLE ONLY.

Independent of autograd: with x = u·(i⁺ − i⁻) + b[i⁺] − b[i⁻] and c = σ(−x) the gradients of
Σ softplus(−x) + ½Σ(λ_u‖u‖² + λ_i‖i⁺‖² + λ_n‖i⁻‖²) are (verified there against the
reference's autograd):  ∂u = −c(i⁺−i⁻) + λ_u u,  ∂i⁺ = −c·u + λ_i i⁺,  ∂i⁻ = +c·u + λ_n i⁻,
∂b[i⁺] = −c, ∂b[i⁻] = +c, duplicates summed (reference model.py:38,65-66: the loss is a sum).
Vectorised, float64: fast enough for BASELINE-size steps (65 536+ triples of the ML-20M shape).
"""
from __future__ import annotations

import numpy as np


def sgd_step(user_emb, item_emb, u, i, j, lr, reg=(0.0, 0.0, 0.0), item_bias=None):
    """Returns (bpr_loss, l2_reg, touched rows -> new values) without copying the tables:
    {'users': ids, 'user_rows': (n,D), 'items': ids, 'item_rows': (m,D), 'bias': (m,) | None}."""
    ru, ri, rn = reg
    U = user_emb[u].astype(np.float64)
    P = item_emb[i].astype(np.float64)
    N = item_emb[j].astype(np.float64)
    x = (U * (P - N)).sum(1)
    if item_bias is not None:
        x = x + item_bias[i].astype(np.float64) - item_bias[j].astype(np.float64)
    c = 1.0 / (1.0 + np.exp(x))
    bpr = np.logaddexp(0.0, -x).sum()
    l2 = 0.5 * (ru * (U * U).sum() + ri * (P * P).sum() + rn * (N * N).sum())
    users, uinv = np.unique(u, return_inverse=True)
    gu = np.zeros((users.size, U.shape[1]))
    np.add.at(gu, uinv, -c[:, None] * (P - N) + ru * U)
    items, iinv = np.unique(np.concatenate([i, j]), return_inverse=True)
    gi = np.zeros((items.size, U.shape[1]))
    np.add.at(gi, iinv[:i.size], -c[:, None] * U + ri * P)
    np.add.at(gi, iinv[i.size:], c[:, None] * U + rn * N)
    out = {"users": users, "user_rows": user_emb[users].astype(np.float64) - lr * gu,
           "items": items, "item_rows": item_emb[items].astype(np.float64) - lr * gi, "bias": None}
    if item_bias is not None:
        gb = np.zeros(items.size)
        np.add.at(gb, iinv[:i.size], -c)
        np.add.at(gb, iinv[i.size:], c)
        out["bias"] = item_bias[items].astype(np.float64) - lr * gb
    return float(bpr), float(l2), out
