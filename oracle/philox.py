"""numpy restatement of the negative-sampler specification.  TEST ORACLE ONLY.

Specification.  For triple id t (index into the COO flattening of the CSR, reference
experiments/bpr/dataset.py:153-156), global step `step` and 64-bit `seed`:

  block b = 0,1,...,255:
      (w0,w1,w2,w3) = Philox4x32-10(counter = (lo32(o), hi32(o), lo32(t), hi32(t)),
                                    key = (lo32(seed), hi32(seed))),  o = (step << 8) | b
  UNIFORM  : each word w is one attempt:  m = w * (I-1);  if lo32(m) < (2^32 mod (I-1)): skip
             (Lemire's rejection, exactly uniform);  j = 1 + hi32(m);
             accept iff j not in seen(user(t))           -> uniform over {1..I-1} \\ seen
  WEIGHTED : each word pair (w0,w1),(w2,w3) is one attempt: m = w_a * I; Lemire skip as above;
             col = hi32(m); uf = float32(w_b >> 8) * 2^-24; j = col if uf < prob[col] else
             alias[col]; skip if j == 0; accept iff j not in seen(user(t))
                                                         -> ∝ weight over {1..I-1} \\ seen
  first accepted attempt wins; 1024 (512) failures = error.

The target distribution is the one the reference samples from
(revisit_bpr/modules/neg_samplers.py:135-141 + torch.multinomial at :31-37;
experiments/bpr/exp.py:282-293 for the popularity-weighted variant); the random stream is
ours and is what "bit-exact given the same seed" refers to.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _seen_mask(indptr, indices, users, cand):
    """cand[k] in row users[k] of the CSR?  Vectorised via a global sorted key."""
    num_items = int(indices.max()) + 2 if indices.size else 2
    row_of = np.repeat(np.arange(indptr.size - 1, dtype=np.int64), np.diff(indptr))
    keys = row_of * num_items + indices.astype(np.int64)  # ascending because rows are sorted
    q = users.astype(np.int64) * num_items + cand.astype(np.int64)
    pos = np.searchsorted(keys, q)
    pos = np.minimum(pos, keys.size - 1)
    return keys[pos] == q


def sample_negatives(indptr, indices, coo_user, triple_idx, seed, step, num_items,
                     alias=None):
    """Negatives for the given triple ids. alias=(prob float32, alias int32) selects WEIGHTED."""
    t = np.asarray(triple_idx, dtype=np.int64)
    users = np.asarray(coo_user)[t]
    out = np.full(t.size, -1, dtype=np.int64)
    pending = np.arange(t.size)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    n = np.uint64(num_items - 1 if alias is None else num_items)
    thresh = np.uint64((1 << 32) % int(n))
    for blk in range(256):
        if pending.size == 0:
            break
        off = (int(step) << 8) | blk
        tp = t[pending]
        words = philox4x32_10(np.uint32(off & 0xFFFFFFFF), np.uint32((off >> 32) & 0xFFFFFFFF),
                              (tp & 0xFFFFFFFF).astype(np.uint32), (tp >> 32).astype(np.uint32),
                              k0, k1)
        attempts = range(4) if alias is None else range(2)
        alive = np.ones(tp.size, dtype=bool)
        cand_final = np.full(tp.size, -1, dtype=np.int64)
        for a in attempts:
            if alias is None:
                w = words[a].astype(np.uint64)
                m = w * n
                ok = (m & MASK32) >= thresh
                j = (m >> np.uint64(32)).astype(np.int64) + 1
            else:
                w = words[2 * a].astype(np.uint64)
                m = w * n
                ok = (m & MASK32) >= thresh
                col = (m >> np.uint64(32)).astype(np.int64)
                uf = (words[2 * a + 1] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
                j = np.where(uf < alias[0][col], col, alias[1][col].astype(np.int64))
                ok &= j != 0
            ok &= alive
            if ok.any():
                seen = _seen_mask(indptr, indices, users[pending], j)
                acc = ok & ~seen
                cand_final[acc] = j[acc]
                alive &= ~acc
        done = cand_final >= 0
        out[pending[done]] = cand_final[done]
        pending = pending[~done]
    if pending.size:
        raise RuntimeError("negative sampler exhausted its attempts")
    return out


def sample_negatives_padded(seen, num_items, num, seed, step, alias=None):
    """Padded-seen variant (rbpr_sample_negatives_padded): seen (B,S) int64 0-padded, any order;
    slot = row*num + s is the Philox subsequence.  Returns (B,num) int64."""
    seen = np.asarray(seen, dtype=np.int64)
    B = seen.shape[0]
    out = np.zeros((B, num), dtype=np.int64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    n = num_items - 1 if alias is None else num_items
    thresh = (1 << 32) % n
    for row in range(B):
        banned = set(int(x) for x in seen[row])
        for s in range(num):
            slot = row * num + s
            res = -1
            for blk in range(256):
                off = (int(step) << 8) | blk
                w = [int(x) for x in philox4x32_10(np.uint32(off & 0xFFFFFFFF), np.uint32((off >> 32) & 0xFFFFFFFF),
                                                   np.uint32(slot & 0xFFFFFFFF), np.uint32(slot >> 32), k0, k1)]
                if alias is None:
                    for a in range(4):
                        m = w[a] * n
                        if (m & 0xFFFFFFFF) < thresh:
                            continue
                        j = 1 + (m >> 32)
                        if j not in banned:
                            res = j
                            break
                else:
                    for a in range(2):
                        m = w[2 * a] * n
                        if (m & 0xFFFFFFFF) < thresh:
                            continue
                        col = m >> 32
                        uf = np.float32(w[2 * a + 1] >> 8) * np.float32(1.0 / 16777216.0)
                        j = col if uf < alias[0][col] else int(alias[1][col])
                        if j == 0 or j in banned:
                            continue
                        res = j
                        break
                if res >= 0:
                    break
            if res < 0:
                raise RuntimeError("negative sampler exhausted its attempts")
            out[row, s] = res
    return out
