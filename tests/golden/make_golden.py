"""Mint golden vectors by running the UNMODIFIED reference (/root/reference) on fixed inputs.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/train_*.npz, metrics.npz, sampler.npz.  `accelerate` is stubbed because
revisit_bpr/metrics/metric.py:5 imports it for a type annotation only.
The reference publishes no fixtures of its own, so these files are the pin.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def import_reference():
    stub = types.ModuleType("accelerate")
    stub.Accelerator = type("Accelerator", (), {})
    sys.modules.setdefault("accelerate", stub)
    # make sure `revisit_bpr` resolves to the reference, not to our drop-in package
    for k in [k for k in sys.modules if k == "revisit_bpr" or k.startswith("revisit_bpr.")]:
        del sys.modules[k]
    sys.path.insert(0, str(REF))
    import revisit_bpr  # noqa: F401
    from revisit_bpr.metrics import NDCG, Recall
    from revisit_bpr.models.bpr import MF, Model
    from revisit_bpr.modules import UniformSampler
    assert Path(revisit_bpr.__file__).resolve().is_relative_to(REF)
    return Model, MF, UniformSampler, NDCG, Recall


def adaptive_case(Model, MF):
    """The reference AdaptiveSampler on fixed inputs.  Its two random draws (factor, geometric) are
    recovered by replaying the same generator on the same calls, so the golden file pins the
    DETERMINISTIC part: (factor, geometric draw, user row, seen row, snapshot) -> item."""
    from revisit_bpr.modules import AdaptiveSampler
    rng = np.random.default_rng(11)
    U, I, D, B = 60, 45, 12, 96
    indptr, indices = tiny_csr(U, I, 2, 25, rng)
    torch.manual_seed(17)
    model = Model(MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0)))
    with torch.no_grad():
        model.logits_model._user_emb.weight.mul_(D * 2.0)
        model.logits_model._item_emb.weight.mul_(D * 2.0)
    users = torch.as_tensor(rng.integers(1, U, size=B))
    seen = padded_seen(indptr, indices, users.tolist())
    p = 0.2
    sampler = AdaptiveSampler(model, I, p, torch.Generator().manual_seed(5), every=1000)
    sampler.update_stats()
    batch = {"user": users, "item": torch.zeros(B, 1, dtype=torch.long), "seen_items": seen}
    negs = sampler.sample(batch)
    # replay the generator: same calls in the same order (neg_samplers.py:84-94)
    g = torch.Generator().manual_seed(5)
    feats = model.logits_model.get_features()
    factor = torch.multinomial(feats["user"].abs()[users] * sampler._factor_std, num_samples=1, generator=g)
    geom = torch.empty_like(factor).geometric_(p, generator=g)
    np.savez_compressed(OUT / "adaptive.npz", indptr=indptr, indices=indices, users=users.numpy(),
                        seen=seen.numpy(), user_emb=feats["user"].detach().numpy(),
                        item_emb=feats["item"].detach().numpy(), snapshot=sampler._factor_to_items.numpy(),
                        factor_std=sampler._factor_std.numpy().ravel(), factor=factor.numpy().ravel(),
                        geom=geom.numpy().ravel(), negs=negs.numpy().ravel(), p=p)
    print("adaptive negs", negs.ravel()[:8].tolist(), "factors", factor.ravel()[:8].tolist())


def tiny_csr(num_users, num_items, deg_lo, deg_hi, rng):
    indptr = [0, 0]
    idx = []
    for _ in range(1, num_users):
        d = int(rng.integers(deg_lo, deg_hi + 1))
        row = np.sort(rng.choice(np.arange(1, num_items), size=d, replace=False))
        idx.append(row)
        indptr.append(indptr[-1] + d)
    return np.asarray(indptr, dtype=np.int64), np.concatenate(idx).astype(np.int32)


def padded_seen(indptr, indices, users):
    rows = [torch.as_tensor(indices[indptr[u]:indptr[u + 1]], dtype=torch.long) for u in users]
    return torch.nn.utils.rnn.pad_sequence(rows, batch_first=True, padding_value=0)


def train_case(name, Model, MF, UniformSampler, *, U, I, D, B, steps, opt, opt_kw, reg, bias, seed):
    rng = np.random.default_rng(seed)
    indptr, indices = tiny_csr(U, I, 3, min(12, I - 3), rng)
    coo_user = np.repeat(np.arange(U), np.diff(indptr))
    nnz = indices.size
    torch.manual_seed(seed)
    model = Model(MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0),
                     item_bias=bias), reg_alphas=reg, fuse_forward=True)
    if bias:  # biases start at zero in the reference; perturb so the bias path is exercised
        with torch.no_grad():
            model.logits_model._item_bias.copy_(torch.randn(I) * 0.1)
            model.logits_model._item_bias[0] = 0
    # scale up the init so the loss is not ~log 2 everywhere
    with torch.no_grad():
        model.logits_model._user_emb.weight.mul_(D * 1.5)
        model.logits_model._item_emb.weight.mul_(D * 1.5)
    init = {k: v.detach().clone().numpy() for k, v in model.logits_model.get_features().items() if v is not None}
    optimizer = getattr(torch.optim, opt)(model.parameters(), **opt_kw)
    gen = torch.Generator().manual_seed(seed)
    sampler = UniformSampler(I, gen)
    perm = torch.randperm(nnz, generator=torch.Generator().manual_seed(seed)).numpy()
    rec = {"triples": [], "negs": [], "bpr_loss": [], "l2_reg": [], "logits_abs_mean": []}
    model.train()
    for s in range(steps):
        t = perm[np.arange(s * B, (s + 1) * B) % nnz]
        users = torch.as_tensor(coo_user[t], dtype=torch.long)
        items = torch.as_tensor(indices[t], dtype=torch.long).unsqueeze(-1)
        batch = {"user": users, "item": items, "seen_items": padded_seen(indptr, indices, users.tolist())}
        batch["neg"] = sampler.sample(batch)
        out = model(batch)
        out["loss"].backward()
        optimizer.step()
        optimizer.zero_grad()
        rec["triples"].append(t)
        rec["negs"].append(batch["neg"].squeeze(-1).numpy())
        rec["bpr_loss"].append(out["bpr_loss"].item())
        rec["l2_reg"].append(float(out["l2_reg"].detach()))
        rec["logits_abs_mean"].append(out["logits"].abs().mean().item())
    final = {k: v.detach().numpy() for k, v in model.logits_model.get_features().items() if v is not None}
    np.savez_compressed(
        OUT / f"train_{name}.npz", indptr=indptr, indices=indices, U=U, I=I, D=D, B=B,
        opt=opt, opt_kw=np.asarray(repr(opt_kw)), reg=np.asarray(repr(reg)), bias=bias,
        triples=np.stack(rec["triples"]), negs=np.stack(rec["negs"]),
        bpr_loss=np.asarray(rec["bpr_loss"]), l2_reg=np.asarray(rec["l2_reg"]),
        logits_abs_mean=np.asarray(rec["logits_abs_mean"]),
        **{f"init_{k}": v for k, v in init.items()}, **{f"final_{k}": v for k, v in final.items()})
    print(name, "bpr", rec["bpr_loss"][0], "->", rec["bpr_loss"][-1])


def metrics_case(NDCG, Recall):
    rng = np.random.default_rng(7)
    n, I = 24, 97
    output = torch.as_tensor(rng.standard_normal((n, I)).astype(np.float32))
    target = torch.as_tensor((rng.random((n, I)) < 0.06).astype(np.float32))
    target[3] = 0  # a user without positives still counts (ndcg.py:65-67,77)
    output[:, 0] = -1e13
    res = {"output": output.numpy(), "target": target.numpy()}
    for k in (1, 5, 20, 100):
        res[f"ndcg@{k}"] = NDCG(k).compute(output, target).numpy()
        res[f"recall@{k}"] = Recall(k).compute(output, target).numpy()
    m = NDCG(20)
    m(output[:10], target[:10])
    m(output[10:], target[10:])
    res["ndcg@20_stream"] = m.get_metric().numpy()
    # the analytic KATs listed in SURVEY.md §4
    o = torch.tensor([[.1, .9, .8, .7, .2], [.5, .4, .3, .2, .1], [.5, .4, .3, .2, .1]])
    t = torch.tensor([[0., 0, 1, 0, 1], [0, 0, 0, 0, 0], [1, 1, 0, 0, 0]])
    res["kat_output"], res["kat_target"] = o.numpy(), t.numpy()
    res["kat_ndcg@3"] = NDCG(3).compute(o, t).numpy()
    res["kat_recall@3"] = Recall(3).compute(o, t).numpy()
    np.savez_compressed(OUT / "metrics.npz", **res)
    print("metrics kat ndcg@3", res["kat_ndcg@3"], "recall@3", res["kat_recall@3"])


def sampler_case(UniformSampler):
    rng = np.random.default_rng(5)
    U, I = 40, 30
    indptr, indices = tiny_csr(U, I, 2, 20, rng)
    users = rng.integers(1, U, size=64)
    seen = padded_seen(indptr, indices, users.tolist())
    gen = torch.Generator().manual_seed(99)
    s = UniformSampler(I, gen)
    negs = s.sample({"item": torch.zeros(64, 1, dtype=torch.long), "seen_items": seen})
    np.savez_compressed(OUT / "sampler.npz", indptr=indptr, indices=indices, users=users, I=I,
                        seed=99, negs=negs.squeeze(-1).numpy(), torch_version=np.asarray(torch.__version__))
    print("sampler negs", negs.squeeze(-1)[:8].tolist())


def main():
    Model, MF, UniformSampler, NDCG, Recall = import_reference()
    common = dict(U=50, I=37, D=16, B=64, steps=6, seed=13)
    train_case("sgd_reg3", Model, MF, UniformSampler, opt="SGD", opt_kw={"lr": 0.05},
               reg={"user": 0.0016, "item": 0.0001, "neg": 0.00375}, bias=False, **common)
    train_case("sgd_bias_all", Model, MF, UniformSampler, opt="SGD", opt_kw={"lr": 0.05},
               reg={"all": 0.01}, bias=True, **common)
    train_case("sgd_noreg", Model, MF, UniformSampler, opt="SGD", opt_kw={"lr": 0.1},
               reg=None, bias=False, **{**common, "D": 12})
    train_case("adam_all", Model, MF, UniformSampler, opt="Adam",
               opt_kw={"lr": 1e-2, "betas": (0.9, 0.999)}, reg={"all": 0.00043}, bias=False, **common)
    train_case("adam_bias_ui", Model, MF, UniformSampler, opt="Adam",
               opt_kw={"lr": 5e-3, "betas": (0.8, 0.99)}, reg={"user": 0.01, "item": 0.02}, bias=True,
               **{**common, "steps": 9})
    train_case("sgdm_nesterov", Model, MF, UniformSampler, opt="SGD",
               opt_kw={"lr": 0.02, "momentum": 0.9, "nesterov": True},
               reg={"user": 0.0016, "item": 0.0001, "neg": 0.00375}, bias=True, **{**common, "steps": 8})
    train_case("rmsprop", Model, MF, UniformSampler, opt="RMSprop",
               opt_kw={"lr": 0.005, "alpha": 0.9, "momentum": 0.0}, reg={"all": 0.001}, bias=True,
               **{**common, "steps": 8})
    metrics_case(NDCG, Recall)
    sampler_case(UniformSampler)
    adaptive_case(Model, MF)


if __name__ == "__main__":
    main()
