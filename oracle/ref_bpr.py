"""torch-CPU restatement of the reference's BPR training/eval op sequence.  TEST ORACLE ONLY.

Each function names the reference lines it restates (paths relative to the reference repo).
Pinned against the reference itself by tests/golden/make_golden.py -> tests/golden/*.npz
(tests/test_oracle_golden.py); tests/golden/make_golden.py re-runs against /root/reference in the
build container and reproduces every committed fixture bit for bit.

This is also the "port" CPU baseline bench.py times: it performs the same work as the
reference's own CPU path (materialised (B,I) sampling weights + torch.multinomial, dense
autograd gradients, dense torch.optim step), with all host threads torch can use.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class RefModel:
    """Tables + the train/eval forward of revisit_bpr/models/bpr/model.py:40-68 (Model) and
    :131-145 (MF), with regularisation :70-93 and the loss of loss.py:19-21 (sum, not mean)."""

    def __init__(self, user_emb: torch.Tensor, item_emb: torch.Tensor,
                 item_bias: torch.Tensor | None = None, reg_alphas: dict | None = None) -> None:
        self.user_emb = user_emb.detach().clone().requires_grad_(True)
        self.item_emb = item_emb.detach().clone().requires_grad_(True)
        self.item_bias = None if item_bias is None else item_bias.detach().clone().requires_grad_(True)
        self.reg = resolve_reg(reg_alphas)

    def parameters(self) -> list[torch.Tensor]:
        ps = [self.user_emb, self.item_emb]
        if self.item_bias is not None:
            ps.append(self.item_bias)
        return ps

    def logits(self, user: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        # model.py:134-138: u (B,D), v (B,...,D) -> sum over D (+ item bias)
        # padding_idx=0 of nn.Embedding only blocks the gradient of row 0 (never indexed here)
        u = F.embedding(user, self.user_emb, padding_idx=0)
        v = F.embedding(items, self.item_emb, padding_idx=0)
        out = torch.einsum("bh,b...h->b...", u, v)
        if self.item_bias is not None:
            out = out + self.item_bias[items]
        return out

    def train_forward(self, user: torch.Tensor, item: torch.Tensor, neg: torch.Tensor) -> dict:
        # model.py:48-68 with fuse_forward=True: one logits call over hstack((item, neg))
        if item.dim() < 2:
            item = item.unsqueeze(-1)  # exp.py:359-360
        if neg.dim() < 2:
            neg = neg.unsqueeze(-1)
        both = self.logits(user, torch.hstack((item, neg)))
        n_pos = item.size(-1)
        pos, ng = both[:, :n_pos], both[:, n_pos:]
        x = pos - ng
        bpr = (-F.logsigmoid(x)).sum()  # loss.py:19-21, size_average=False (model.py:38)
        ru, ri, rn = self.reg
        # model.py:87-93: squared norms of the gathered rows, halved
        l2 = (ri * self.item_emb[item].pow(2).flatten(1).sum(1)
              + rn * self.item_emb[neg].pow(2).flatten(1).sum(1)
              + ru * self.user_emb[user].pow(2).flatten(1).sum(1)) / 2
        l2 = l2.sum()
        return {"logits_pos": pos, "logits_neg": ng, "logits": x, "bpr_loss": bpr, "l2_reg": l2, "loss": bpr + l2}

    @torch.no_grad()
    def eval_logits(self, users: torch.Tensor, seen_padded: torch.Tensor | None) -> torch.Tensor:
        # AllItemsCollator (experiments/bpr/dataset.py:279-296): item = arange(I) per user;
        # MF.forward gathers (B,I,D) and contracts (model.py:134-138); then
        # _remove_seen_items (experiments/bpr/exp.py:369-374)
        num_items = self.item_emb.size(0)
        items = torch.arange(num_items).unsqueeze(0).expand(users.numel(), -1)
        out = self.logits(users, items).clone()
        if seen_padded is not None:
            out.scatter_(dim=-1, index=seen_padded, value=-1e13)
            out[:, 0] = -1e13
        return out


def resolve_reg(reg_alphas: dict | None) -> tuple[float, float, float]:
    """model.py:74-86: `all` overrides the others; missing -> 0; `neg` falls back to `item`."""
    r = reg_alphas or {}
    a, u, i, n = r.get("all"), r.get("user"), r.get("item"), r.get("neg")
    if a is None and u is None and i is None and n is None:
        return 0.0, 0.0, 0.0
    if a is not None:
        u = i = n = a
    u = u or 0
    i = i or 0
    n = n or i
    return float(u), float(i), float(n)


def make_optimizer(model: RefModel, kind: str, **kw) -> torch.optim.Optimizer:
    """The dense torch optimizers the configs instantiate (exp.py:103-105)."""
    if kind == "sgd":
        return torch.optim.SGD(model.parameters(), **kw)
    if kind == "adam":
        return torch.optim.Adam(model.parameters(), **kw)
    if kind == "rmsprop":
        return torch.optim.RMSprop(model.parameters(), **kw)
    raise ValueError(kind)


def train_step(model: RefModel, opt: torch.optim.Optimizer, user, item, neg) -> dict:
    """experiments/trainer.py:64-83: forward, backward, step, zero_grad."""
    out = model.train_forward(user, item, neg)
    out["loss"].backward()
    opt.step()
    opt.zero_grad()
    return {k: v.detach() for k, v in out.items()}


def sampling_weights(item_weights: torch.Tensor, seen_padded: torch.Tensor) -> torch.Tensor:
    """revisit_bpr/modules/neg_samplers.py:135-141 (== experiments/bpr/exp.py:282-288)."""
    w = item_weights.unsqueeze(0).repeat(seen_padded.size(0), 1)
    w.scatter_(dim=-1, index=seen_padded, value=0.0)
    w[:, 0] = 0.0
    w *= w.sum(dim=-1, keepdim=True).reciprocal()
    return w


def reference_style_negatives(item_weights, seen_padded, gen: torch.Generator, num: int = 1):
    """UniformSampler.sample (neg_samplers.py:31-37) / _static_sampling (exp.py:290-293)."""
    return torch.multinomial(sampling_weights(item_weights, seen_padded), num_samples=num,
                             generator=gen)


def padded_seen(indptr, indices, users: torch.Tensor) -> torch.Tensor:
    """Rows of the dense 0-padded seen matrix of experiments/bpr/dataset.py:157-163,175-181."""
    rows = [torch.as_tensor(indices[indptr[u]:indptr[u + 1]], dtype=torch.long) for u in users.tolist()]
    rows = [r if r.numel() else torch.zeros(1, dtype=torch.long) for r in rows]
    return torch.nn.utils.rnn.pad_sequence(rows, batch_first=True, padding_value=0)


def _sorted_target(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    # revisit_bpr/metrics/metric.py:110-113
    return torch.gather(target, dim=-1, index=torch.argsort(-output, dim=-1))


def ndcg_at_k(output: torch.Tensor, target: torch.Tensor, k: int) -> torch.Tensor:
    """revisit_bpr/metrics/ndcg.py:8-13,69-78 (exponential gain)."""
    k = min(output.size(-1), k)
    disc = torch.log2(torch.arange(k, dtype=torch.float) + 2.0)

    def dcg(t):
        return ((2 ** t - 1) / disc).sum(-1)
    got = dcg(_sorted_target(output, target)[:, :k])
    ideal = dcg(_sorted_target(target, target)[:, :k])
    return torch.nan_to_num(got / ideal)


def recall_at_k(output: torch.Tensor, target: torch.Tensor, k: int) -> torch.Tensor:
    """revisit_bpr/metrics/recall.py:44-51."""
    k = min(output.size(-1), k)
    hits = _sorted_target(output, target)[:, :k].sum(-1)
    return torch.nan_to_num(hits / target.sum(-1))


def multi_hot(held_indptr, held_indices, num_items: int) -> torch.Tensor:
    """target of AllItemsCollator (experiments/bpr/dataset.py:283-285)."""
    n = len(held_indptr) - 1
    t = torch.zeros(n, num_items)
    for r in range(n):
        t[r, torch.as_tensor(held_indices[held_indptr[r]:held_indptr[r + 1]], dtype=torch.long)] = 1.0
    return t
