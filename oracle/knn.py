"""numpy restatement of the reference's item-neighbourhood logits models.  TEST ORACLE ONLY.

Follows revisit_bpr/models/bpr/model.py:176-196 (ItemKNN.forward) and :222-248
(FreeItemKNN.forward), float64, explicit loops over the batch.  The gradients are written out by
hand (the reference gets them from autograd) and pinned against the reference's autograd through
tests/golden/knn.npz (tests/test_oracle_knn.py):

  ItemKNN      x[b,i] = w[it[b,i]] . P_b + bias[it[b,i]],  P_b = SUM_{s kept} w[seen[b,s]]
               dw[it[b,i]] += g[b,i] P_b ;  dw[seen[b,s] kept] += SUM_i g[b,i] w[it[b,i]] ;  dbias[it[b,i]] += g[b,i]
  FreeItemKNN  x[b,i] = SUM_{s kept} W[it[b,i], seen[b,s]] + bias[it[b,i]]
               dW[it[b,i], seen[b,s] kept] += g[b,i] ;  dbias[it[b,i]] += g[b,i]

"kept" = the seen id does not occur among it[b,:] (model.py:184-190, 230-235).  Row 0 is an
ordinary row: the reference zeroes it at init only (model.py:173-174).
"""
from __future__ import annotations

import numpy as np


def keep_mask(item: np.ndarray, seen: np.ndarray) -> np.ndarray:
    """(B,S) bool: True where the seen entry takes part in the sum."""
    return ~(seen[:, None, :] == item[:, :, None]).any(axis=1)


def itemknn_forward(w, bias, item, seen):
    w = w.astype(np.float64)
    keep = keep_mask(item, seen)
    out = np.zeros(item.shape)
    for b in range(item.shape[0]):
        profile = w[seen[b][keep[b]]].sum(0)
        out[b] = w[item[b]] @ profile
    if bias is not None:
        out = out + bias.astype(np.float64)[item]
    return out


def itemknn_backward(w, item, seen, grad, with_bias):
    w = w.astype(np.float64)
    keep = keep_mask(item, seen)
    gw = np.zeros_like(w)
    gb = np.zeros(w.shape[0]) if with_bias else None
    for b in range(item.shape[0]):
        kept = seen[b][keep[b]]
        profile = w[kept].sum(0)
        np.add.at(gw, item[b], grad[b][:, None] * profile[None, :])
        np.add.at(gw, kept, (grad[b][:, None] * w[item[b]]).sum(0)[None, :].repeat(kept.size, 0))
        if with_bias:
            np.add.at(gb, item[b], grad[b])
    return gw, gb


def freeknn_forward(W, bias, item, seen):
    W = W.astype(np.float64)
    keep = keep_mask(item, seen)
    out = np.zeros(item.shape)
    for b in range(item.shape[0]):
        out[b] = W[item[b]][:, seen[b][keep[b]]].sum(1)
    if bias is not None:
        out = out + bias.astype(np.float64)[item]
    return out


def freeknn_backward(num_items, item, seen, grad, with_bias):
    keep = keep_mask(item, seen)
    gW = np.zeros((num_items, num_items))
    gb = np.zeros(num_items) if with_bias else None
    for b in range(item.shape[0]):
        kept = seen[b][keep[b]]
        for i, g in zip(item[b], grad[b]):
            np.add.at(gW[i], kept, g)
        if with_bias:
            np.add.at(gb, item[b], grad[b])
    return gW, gb


def bpr_step(kind, w, bias, item, neg, seen, reg=(0.0, 0.0), fuse=False):
    """Train-mode Model.forward (model.py:48-68) + backward for a KNN logits model.
    reg = (item, neg) L2 weights on the rows of features['item'] (model.py:86-90).
    Returns dict(logits_pos, logits_neg, bpr_loss, l2_reg, grad_w, grad_bias)."""
    fwd = itemknn_forward if kind == "itemknn" else freeknn_forward
    if fuse:
        both = fwd(w, bias, np.concatenate([item, neg], 1), seen)
        pos, ng = both[:, :item.shape[1]], both[:, item.shape[1]:]
    else:
        pos, ng = fwd(w, bias, item, seen), fwd(w, bias, neg, seen)
    x = pos - ng
    c = 1.0 / (1.0 + np.exp(x))  # -d softplus(-x)/dx
    w64 = w.astype(np.float64)
    ri, rn = reg
    l2 = 0.5 * (ri * (w64[item] ** 2).sum() + rn * (w64[neg] ** 2).sum())

    def back(ids, g):
        if kind == "itemknn":
            return itemknn_backward(w, ids, seen, g, bias is not None)
        return freeknn_backward(w.shape[0], ids, seen, g, bias is not None)

    if fuse:
        gw, gb = back(np.concatenate([item, neg], 1), np.concatenate([-c, c], 1))
    else:
        gw, gb = back(item, -c)
        gw2, gb2 = back(neg, c)
        gw = gw + gw2
        gb = None if gb is None else gb + gb2
    np.add.at(gw, item.reshape(-1), ri * w64[item.reshape(-1)])
    np.add.at(gw, neg.reshape(-1), rn * w64[neg.reshape(-1)])
    return {"logits_pos": pos, "logits_neg": ng, "bpr_loss": np.logaddexp(0.0, -x).sum(), "l2_reg": l2,
            "grad_w": gw, "grad_bias": gb}
