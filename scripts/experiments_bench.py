"""Experiment-level measurement for bench.py (`configs.e2e_experiment`): what
`experiments.bpr.Experiment.run()` achieves when it is driven by a config in the reference's own
schema (the keys of configs/RQ2/neg-sampling/*.yaml.j2: `experiment` with the 14 stock metrics,
`datasets` of torch DataLoaders over the jsonl dataset classes, `model`, `optimizer`) —
  * stock:      train_batch_size 256, every batch through DataLoader -> collate_fn -> sampler hook ->
                Model.forward (one fused step per call), exactly the reference's control flow;
  * fast_train: the same config plus `fast_train: true` (whole chunks of steps inside the library).
Data: ML-20M shape at 1/10 of the users (13.6 k users x 20 108 items, ~0.97 M interactions), written
to disk in the reference's jsonl format (bin/datasets/format-repro.sh:56-81); eval on 2 048 users.
Times are wall clock around the train / eval engines with a device synchronise on both sides."""
from __future__ import annotations

import json
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

CONFIG = """
num_users: &num_users {{ (num_users | int) + 1 }}
num_items: &num_items {{ (num_items | int) + 1 }}
epochs: 1
experiment:
  _target_: experiments.bpr.Experiment
  early_stopping_metric: ndcg@100
  early_stopping_patience: 13
  metrics:
    ndcg@100: {_target_: revisit_bpr.metrics.NDCG, topk: 100}
    recall@100: {_target_: revisit_bpr.metrics.Recall, topk: 100}
    ndcg@10: {_target_: revisit_bpr.metrics.NDCG, topk: 10}
    recall@10: {_target_: revisit_bpr.metrics.Recall, topk: 10}
    auc: {_target_: revisit_bpr.metrics.RocAucManySlow}
    ndcg@5: {_target_: revisit_bpr.metrics.NDCG, topk: 5}
    recall@5: {_target_: revisit_bpr.metrics.Recall, topk: 5}
    recall@20: {_target_: revisit_bpr.metrics.Recall, topk: 20}
    ndcg@50: {_target_: revisit_bpr.metrics.NDCG, topk: 50}
    recall@50: {_target_: revisit_bpr.metrics.Recall, topk: 50}
    precision@5: {_target_: revisit_bpr.metrics.Precision, topk: 5}
    precision@10: {_target_: revisit_bpr.metrics.Precision, topk: 10}
    precision@50: {_target_: revisit_bpr.metrics.Precision, topk: 50}
    precision@100: {_target_: revisit_bpr.metrics.Precision, topk: 100}
datasets:
  train:
    _target_: torch.utils.data.DataLoader
    dataset:
      _target_: experiments.bpr.dataset.SparseSamplingInMemoryWithCollator
      path: {{ dataset }}/full-train-with-fold-in.jsonl
      seen_items_path: {{ dataset }}/full-train-with-fold-in-user-seen-items.jsonl
      num_users: *num_users
      num_items: *num_items
      put_on_cuda: true
    batch_size: {{ train_batch_size | int }}
    shuffle: true
    {% if max_iters %}max_iters: {{ max_iters }}{% endif %}
  eval:
    _target_: torch.utils.data.DataLoader
    dataset:
      _target_: experiments.bpr.dataset.InMemory
      path: {{ dataset }}/test-grouped.jsonl
      seen_items_path: {{ dataset }}/full-train-with-fold-in-user-seen-items.jsonl
    collate_fn:
      _target_: experiments.bpr.dataset.AllItemsCollator
      num_items: *num_items
    batch_size: 128
    shuffle: false
model:
  _target_: revisit_bpr.models.bpr.Model
  fuse_forward: true
  logits_model:
    _target_: revisit_bpr.models.bpr.MF
    item_bias: false
    user_bias: false
    user_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_users, embedding_dim: {{ embedding_dim | int }}, padding_idx: 0}
    item_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_items, embedding_dim: {{ embedding_dim | int }}, padding_idx: 0}
  reg_alphas: {user: 0.0016, item: 0.0001, neg: 0.00375}
optimizer:
  _partial_: true
  _target_: torch.optim.SGD
  lr: 0.05
"""


def write_dataset(root: Path, scale: float = 0.1, n_eval: int = 2048, seed: int = 13):
    from rbpr import synth
    inter = synth.make("ml-20m", seed=seed, scale=scale)
    users, seen, held = synth.split_heldout(inter, n_eval)
    held_of = {int(u): held[1][held[0][r]:held[0][r + 1]] for r, u in enumerate(users)}
    with open(root / "full-train-with-fold-in.jsonl", "w") as ft, \
            open(root / "full-train-with-fold-in-user-seen-items.jsonl", "w") as fs, \
            open(root / "test-grouped.jsonl", "w") as fe:
        n_train = 0
        for u in range(1, inter.num_users):
            row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]]
            if u in held_of:
                row = np.setdiff1d(row, held_of[u])
                fe.write(json.dumps({"user": u, "item": held_of[u].tolist()}) + "\n")
            items = row.tolist()
            if not items:
                continue
            n_train += len(items)
            ft.write("".join(f'{{"user": {u}, "item": {i}}}\n' for i in items))
            fs.write(json.dumps({"user": u, "seen_items": items}) + "\n")
    return inter, n_train, len(users)


def _run(cfg: dict, dev: torch.device) -> dict:
    from experiments._instantiate import instantiate
    from experiments.bpr.exp import BPRExperiment
    from experiments.trainer import Events
    exp_cfg = cfg.pop("experiment")
    exp = instantiate(exp_cfg, exp_config=lambda: cfg, dir=None, debug=False, seed=13, trackers_params={})
    t = {"train": 0.0, "eval": [], "train_iters": 0}
    orig = BPRExperiment._get_trainer

    def spy(self, *a, **k):
        tr = orig(self, *a, **k)

        def tic(engine):
            torch.cuda.synchronize()
            engine.state._t0 = time.perf_counter()

        def toc_train(engine):
            torch.cuda.synchronize()
            t["train"] += time.perf_counter() - engine.state._t0
            t["train_iters"] = engine.state.iteration

        def toc_eval(engine):
            torch.cuda.synchronize()
            t["eval"].append(time.perf_counter() - engine.state._t0)

        tr.add_event("train", Events.EPOCH_STARTED, tic)
        tr.add_event("train", Events.EPOCH_COMPLETED, toc_train)
        tr.add_event("eval", Events.EPOCH_STARTED, tic)
        tr.add_event("eval", Events.EPOCH_COMPLETED, toc_eval)
        return tr

    BPRExperiment._get_trainer = spy
    try:
        exp.run()
    finally:
        BPRExperiment._get_trainer = orig
    t["opt_steps"] = int(exp._model._opt_step)
    t["ndcg@100"] = float(exp.metrics["ndcg@100"])
    t["launches"] = int(exp._model.logits_model.engine().launch_count())
    return t


def run_experiment_bench(dev: torch.device, batch: int = 256, stock_iters: int = 600) -> dict:
    import jinja2
    import yaml
    with tempfile.TemporaryDirectory() as d:
        root = Path(d)
        inter, n_train, n_eval = write_dataset(root)
        out: dict = {"workload": f"experiments.bpr.Experiment.run() from a reference-schema config: ML-20M shape at 1/10 of the users "
                                 f"({inter.num_users - 1} x {inter.num_items - 1}, {n_train} train interactions), dim=128, SGD, "
                                 f"train_batch_size={batch}, 14 stock metrics, eval batch 128 over {n_eval} users"}
        for mode in ("stock", "fast_train"):
            text = jinja2.Template(CONFIG).render(dataset=str(root), num_users=inter.num_users - 1, num_items=inter.num_items - 1,
                                                  train_batch_size=batch, embedding_dim=128,
                                                  max_iters=stock_iters if mode == "stock" else None)
            cfg = yaml.safe_load(text)
            if mode == "fast_train":
                cfg["fast_train"] = True
            r = _run(cfg, dev)
            steps = r["opt_steps"]
            ev = min(r["eval"]) if r["eval"] else None
            out[mode] = {"triples_per_s": steps * batch / r["train"], "us_per_step": 1e6 * r["train"] / max(steps, 1),
                         "steps": steps, "engine_iterations": r["train_iters"], "train_s": r["train"],
                         "eval_users_per_s": (n_eval / ev) if ev else None, "eval_s": ev, "ndcg@100": r["ndcg@100"]}
        return out
