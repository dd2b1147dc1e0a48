"""Mint tests/golden/knn.npz by running the UNMODIFIED reference ItemKNN / FreeItemKNN
(/root/reference/revisit_bpr/models/bpr/model.py:156-251) inside the reference's Model on fixed
inputs: train-mode outputs, autograd gradients, the weights after three SGD steps, eval logits
over a wider item list (with ids that collide with the seen list), fused and unfused forward.

Run in the build container only:  python tests/golden/make_golden_knn.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, import_reference  # noqa: E402


def batch(rng, I, B, S, n_eval):
    seen = np.zeros((B, S), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(1, S + 1))
        seen[b, :k] = np.sort(rng.choice(np.arange(1, I), size=k, replace=False))
    item = np.array([[rng.choice(seen[b][seen[b] > 0])] for b in range(B)])  # a positive is a seen item
    neg = rng.integers(1, I, size=(B, 1))
    wide = rng.integers(0, I, size=(B, n_eval))
    wide[:, 0] = seen[:, 0]  # guaranteed collision with the seen list
    return seen, item, neg, wide


def reg_pair(reg):
    """(item, neg) weights as model.py:79-85 resolves them."""
    if not reg:
        return np.zeros(2)
    if "all" in reg:
        return np.array([reg["all"], reg["all"]])
    ri = reg.get("item") or 0.0
    return np.array([ri, reg.get("neg") or ri])


def run(kind, Model, cls, rng, *, I, H, B, S, bias, fuse, reg, lr=0.05, steps=3, n_eval=6):
    torch.manual_seed(3)
    lm = cls(I, H, bias=bias) if kind == "itemknn" else cls(I, bias=bias)
    with torch.no_grad():  # the reference zeroes row 0 at init only; move it so that it is pinned as ordinary
        lm._weights[0].uniform_(-0.2, 0.2)
        if bias:
            lm._bias.uniform_(-0.3, 0.3)
        lm._weights.mul_(0.3)
    model = Model(lm, reg_alphas=reg, fuse_forward=fuse)
    seen, item, neg, wide = batch(rng, I, B, S, n_eval)
    t = lambda a: torch.as_tensor(a)  # noqa: E731
    inputs = {"user": torch.zeros(B, dtype=torch.long), "item": t(item), "neg": t(neg), "seen_items": t(seen)}
    out = {"w0": lm._weights.detach().numpy().copy(), "seen": seen, "item": item, "neg": neg, "wide": wide,
           "reg": reg_pair(reg),
           "lr": np.array(lr)}
    if bias:
        out["b0"] = lm._bias.detach().numpy().copy()
    model.eval()
    with torch.no_grad():
        out["eval_logits"] = model({"user": inputs["user"], "item": t(wide), "seen_items": t(seen)})["logits"].numpy()
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    losses = []
    for s in range(steps):
        o = model(inputs)
        opt.zero_grad()
        o["loss"].backward()
        if s == 0:
            out.update(logits_pos=o["logits_pos"].detach().numpy(), logits_neg=o["logits_neg"].detach().numpy(),
                       bpr_loss=o["bpr_loss"].detach().numpy(), l2_reg=np.asarray(float(o["l2_reg"])),
                       grad_w=lm._weights.grad.numpy().copy())
            if bias:
                out["grad_b"] = lm._bias.grad.numpy().copy()
        losses.append(float(o["loss"]))
        opt.step()
    out["losses"] = np.array(losses)
    out["w_end"] = lm._weights.detach().numpy().copy()
    if bias:
        out["b_end"] = lm._bias.detach().numpy().copy()
    return out


def main():
    import_reference()
    from revisit_bpr.models.bpr import FreeItemKNN, ItemKNN, Model
    rng = np.random.default_rng(21)
    cases = {
        "itemknn_bias": run("itemknn", Model, ItemKNN, rng, I=23, H=8, B=12, S=7, bias=True, fuse=False,
                            reg={"item": 0.01, "neg": 0.02}),
        "itemknn_fused": run("itemknn", Model, ItemKNN, rng, I=40, H=33, B=9, S=11, bias=False, fuse=True,
                             reg={"all": 0.005}),
        "itemknn_noreg": run("itemknn", Model, ItemKNN, rng, I=31, H=300, B=5, S=4, bias=False, fuse=False, reg=None),
        "freeknn_bias": run("freeknn", Model, FreeItemKNN, rng, I=19, H=0, B=10, S=6, bias=True, fuse=False,
                            reg={"item": 0.01}),
        "freeknn_fused": run("freeknn", Model, FreeItemKNN, rng, I=27, H=0, B=8, S=9, bias=False, fuse=True,
                             reg=None),
    }
    flat = {f"{name}/{k}": v for name, c in cases.items() for k, v in c.items()}
    np.savez_compressed(OUT / "knn.npz", **flat)
    print("wrote", OUT / "knn.npz", {k: float(c["losses"][0]) for k, c in cases.items()})


if __name__ == "__main__":
    main()
