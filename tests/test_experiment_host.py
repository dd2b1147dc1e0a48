"""CPU checks of the experiment-side host code: the instantiate stand-in, the jsonl dataset classes
(reference on-disk format -> COO/CSR, batch dictionaries) and the Experiment constructor surface."""
import ast
import inspect
import json
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference")


def test_instantiate_subset():
    from experiments._instantiate import instantiate
    cfg = {"_target_": "torch.nn.Embedding", "num_embeddings": 5, "embedding_dim": 4, "padding_idx": 0}
    emb = instantiate(cfg)
    assert isinstance(emb, torch.nn.Embedding) and emb.padding_idx == 0
    part = instantiate({"_partial_": True, "_target_": "torch.optim.SGD", "lr": 0.5})
    opt = part(emb.parameters())
    assert isinstance(opt, torch.optim.SGD) and opt.param_groups[0]["lr"] == 0.5
    nested = instantiate({"_target_": "revisit_bpr.models.bpr.MF", "item_bias": True,
                          "user_emb": cfg, "item_emb": dict(cfg, num_embeddings=7)})
    assert nested.get_features()["item"].shape == (7, 4) and nested.get_features()["item_bias"] is not None
    assert instantiate({"a": [1, {"_target_": "builtins.int", "_args_": ["7"]}]}) == {"a": [1, 7]}
    assert instantiate(cfg, embedding_dim=8).embedding_dim == 8  # call-site override


def _write(tmp_path):
    rows = {1: [2, 5, 3], 2: [1], 4: [6, 2]}
    with open(tmp_path / "train.jsonl", "w") as f:
        for u, items in rows.items():
            for i in items:
                f.write(json.dumps({"user": u, "item": i}) + "\n")
        f.write(json.dumps({"user": 1, "item": 5}) + "\n")  # duplicate interaction: the matrix is binary
    with open(tmp_path / "seen.jsonl", "w") as f:
        for u, items in rows.items():
            f.write(json.dumps({"user": u, "seen_items": items}) + "\n")
    with open(tmp_path / "test-grouped.jsonl", "w") as f:
        f.write(json.dumps({"user": 1, "item": [4, 6]}) + "\n")
        f.write(json.dumps({"user": 4, "item": [1]}) + "\n")
    return rows


def test_sparse_sampling_dataset_contract(tmp_path):
    from experiments.bpr.dataset import SparseSamplingInMemoryWithCollator
    rows = _write(tmp_path)
    ds = SparseSamplingInMemoryWithCollator(tmp_path / "train.jsonl", tmp_path / "seen.jsonl", num_users=6, num_items=8)
    assert len(ds) == 6 and ds[3] == 3
    indptr, indices = ds.csr()
    assert indptr.tolist() == [0, 0, 3, 4, 4, 6, 6] and indices.tolist() == [2, 3, 5, 1, 2, 6]
    batch = ds.collate_fn([0, 3, 5])
    assert batch["user"].tolist() == [1, 2, 4] and batch["item"].tolist() == [2, 1, 6]
    assert batch["seen_items"].shape == (3, 3)
    assert batch["seen_items"][0].tolist() == rows[1] and batch["seen_items"][1].tolist() == [1, 0, 0]
    with pytest.raises(IndexError):
        SparseSamplingInMemoryWithCollator(tmp_path / "train.jsonl", tmp_path / "seen.jsonl", num_users=3, num_items=8)


def test_eval_dataset_and_all_items_collator(tmp_path):
    from experiments.bpr.dataset import AllItemsCollator, InMemory, Iter
    _write(tmp_path)
    ds = InMemory(tmp_path / "test-grouped.jsonl", tmp_path / "seen.jsonl")
    assert len(ds) == 2 and ds[0] == {"user": 1, "item": [4, 6], "seen_items": [2, 5, 3]}
    assert list(Iter(tmp_path / "test-grouped.jsonl", tmp_path / "seen.jsonl")) == [ds[0], ds[1]]
    batch = AllItemsCollator(num_items=8)([ds[0], ds[1]])
    assert batch["user"].tolist() == [1, 4]
    assert batch["item"].shape == (2, 8) and batch["item"][1].tolist() == list(range(8))
    assert batch["target"].tolist() == [[0, 0, 0, 0, 1, 0, 1, 0], [0, 1, 0, 0, 0, 0, 0, 0]]
    assert batch["seen_items"].tolist() == [[2, 5, 3], [6, 2, 0]]


def test_experiment_constructor_surface():
    from experiments.bpr import Experiment
    ours = inspect.signature(Experiment.__init__).parameters
    assert list(ours)[:4] == ["self", "exp_config", "dir", "n_checkpoints"]
    assert ours["seed"].default == 13 and ours["skip_seen"].default is True
    assert ours["adaptive_sampling_prob"].default is None and ours["neg_sampling_alpha"].default == 0.0
    if REF.exists():  # same parameter names, order and literal defaults as the reference class
        tree = ast.parse((REF / "experiments/bpr/exp.py").read_text())
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BPRExperiment")
        init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
        ref_names = [a.arg for a in init.args.args]
        assert list(ours) == ref_names
        ref_defaults = [ast.literal_eval(d) for d in init.args.defaults]
        ours_defaults = [p.default for p in ours.values() if p.default is not inspect.Parameter.empty]
        assert ours_defaults == ref_defaults
    with pytest.raises(NotImplementedError):
        Experiment({"num_items": 4, "datasets": {}}, mixed_precision="fp16")


def test_attach_metrics_and_early_stopping_host_logic():
    from experiments._accel import Accelerator
    from experiments.options import attach_early_stopping, attach_metrics
    from experiments.trainer import Trainer
    from revisit_bpr.metrics import Metric

    class CountingMetric(Metric):  # host-only stand-in: sums the logits it is fed
        def __init__(self): self.total = torch.tensor(0.0)
        def state_dict(self): return {"total": self.total}
        def load_state_dict(self, d): self.total = d["total"]
        def __call__(self, output, target): self.total = self.total + self.compute(output, target).sum()
        def compute(self, output, target): return (output * target).sum(-1)
        def get_metric(self, reset=False): return self.total
        def reset(self): self.total = torch.tensor(0.0)

    class EvalOnly(torch.nn.Module):
        def forward(self, batch): return {"logits": batch["x"]}

    model = EvalOnly()
    trainer = Trainer(model, torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1), Accelerator("cpu"))
    m = CountingMetric()
    attach_metrics(trainer, Accelerator("cpu"), {"dot": m})
    es = attach_early_stopping(trainer, "dot", patience=2)
    batches = [{"x": torch.ones(2, 3), "target": torch.eye(3)[:2]}, {"x": 2 * torch.ones(1, 3), "target": torch.eye(3)[:1]}]
    for _ in range(4):  # the same score every time: improvement only on the first call
        trainer.engines["eval"].run(batches)
        assert trainer.engines["eval"].state.metrics["dot"].item() == 4.0  # reset at every epoch start
    assert es.counter == 3 and trainer.engines["train"].should_terminate


def test_library_jsonl_datasets_and_collator(tmp_path):
    from revisit_bpr.datasets.jsonl import Collator, InMemory, Iter
    rows = [{"user": 1, "item": 3, "seen_items": [3, 5]}, {"user": 2, "item": 0, "seen_items": [4]},
            {"user": 3, "item": 1, "seen_items": [9, 8, 7]}]
    p = tmp_path / "rq1.jsonl"
    p.write_text("\n".join(json.dumps(r) for r in rows) + "\n")
    ds = InMemory(p)
    assert len(ds) == 3 and ds[2] == rows[2] and list(Iter(p)) == rows
    batch = Collator(pad=["seen_items"])(rows)
    assert batch["user"].tolist() == [1, 2, 3] and batch["item"].tolist() == [3, 0, 1]
    assert batch["seen_items"].tolist() == [[3, 5, 0], [4, 0, 0], [9, 8, 7]]
    assert batch["seen_items_mask"].tolist() == [[1, 1, 0], [1, 0, 0], [1, 1, 1]]
    loader = torch.utils.data.DataLoader(Iter(p), batch_size=2, collate_fn=Collator(pad=["seen_items"]))
    assert [b["user"].tolist() for b in loader] == [[1, 2], [3]]
