"""End-to-end GPU test of `experiments.bpr.Experiment` driven by a config in the reference's own
schema (jinja2-templated YAML with `_target_` blocks, same keys as
configs/RQ2/neg-sampling/*.yaml.j2) on a small jsonl dataset in the reference's on-disk format:
train with the hooks (negatives, fused step), evaluate (all-item logits, seen masking, metrics),
and check the reported metrics against the oracle evaluated on the final tables."""
import json

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

CONFIG = """
num_users: &num_users {{ (num_users | int) + 1 }}
num_items: &num_items {{ (num_items | int) + 1 }}
epochs: {{ epochs | default(2, true) | int }}
experiment:
  _target_: experiments.bpr.Experiment
  early_stopping_metric: ndcg@10
  early_stopping_patience: 13
  {% if adaptive %}adaptive_sampling_prob: !!float {{ 1 / 20 | float }}{% endif %}
  metrics:
    ndcg@10: {_target_: revisit_bpr.metrics.NDCG, topk: 10}
    recall@20: {_target_: revisit_bpr.metrics.Recall, topk: 20}
    precision@5: {_target_: revisit_bpr.metrics.Precision, topk: 5}
    auc: {_target_: revisit_bpr.metrics.RocAucManySlow}
datasets:
  train:
    _target_: torch.utils.data.DataLoader
    dataset:
      _target_: experiments.bpr.dataset.SparseSamplingInMemoryWithCollator
      path: {{ dataset }}/train.jsonl
      seen_items_path: {{ dataset }}/train-user-seen-items.jsonl
      num_users: *num_users
      num_items: *num_items
      put_on_cuda: true
    batch_size: {{ train_batch_size | int }}
    shuffle: true
  eval:
    _target_: torch.utils.data.DataLoader
    dataset:
      _target_: experiments.bpr.dataset.InMemory
      path: {{ dataset }}/test-grouped.jsonl
      seen_items_path: {{ dataset }}/train-user-seen-items.jsonl
    collate_fn:
      _target_: experiments.bpr.dataset.AllItemsCollator
      num_items: *num_items
    batch_size: 32
    shuffle: false
model:
  _target_: revisit_bpr.models.bpr.Model
  fuse_forward: true
  logits_model:
    _target_: revisit_bpr.models.bpr.MF
    item_bias: {{ item_bias | default(true, true) }}
    user_bias: false
    user_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_users, embedding_dim: {{ embedding_dim | int }}, padding_idx: 0}
    item_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_items, embedding_dim: {{ embedding_dim | int }}, padding_idx: 0}
  reg_alphas: {user: 0.0016, item: 0.0001, neg: 0.00375}
optimizer:
  _partial_: true
  _target_: {{ optimizer }}
  lr: {{ lr }}
"""


def _write_dataset(tmp_path, n_users=80, n_items=60, seed=3):
    from rbpr import synth
    inter = synth.generate("t", n_users, n_items, n_users * 10, 9, 4, 0.8, seed)
    rng = np.random.default_rng(seed)
    train_rows, test_rows = {}, {}
    for u in range(1, inter.num_users):
        row = inter.indices[inter.indptr[u]:inter.indptr[u + 1]].tolist()
        rng.shuffle(row)
        k = max(1, len(row) // 5)
        test_rows[u], train_rows[u] = sorted(row[:k]), sorted(row[k:])
    with open(tmp_path / "train.jsonl", "w") as f:
        for u, items in train_rows.items():
            for i in items:
                f.write(json.dumps({"user": u, "item": i}) + "\n")
    with open(tmp_path / "train-user-seen-items.jsonl", "w") as f:
        for u, items in train_rows.items():
            f.write(json.dumps({"user": u, "seen_items": items}) + "\n")
    with open(tmp_path / "test-grouped.jsonl", "w") as f:
        for u, items in test_rows.items():
            f.write(json.dumps({"user": u, "item": items}) + "\n")
    return inter, train_rows, test_rows


def _render(tmp_path, **kw):
    import jinja2
    text = jinja2.Template(CONFIG, undefined=jinja2.StrictUndefined).render(dataset=str(tmp_path), **kw)
    return yaml.safe_load(text)


@pytest.mark.parametrize("adaptive,optimizer,lr", [(False, "torch.optim.SGD", 0.05), (True, "torch.optim.Adam", 0.01)])
def test_experiment_from_reference_style_config(tmp_path, adaptive, optimizer, lr):
    from experiments._instantiate import instantiate
    from oracle import ref_bpr
    inter, train_rows, test_rows = _write_dataset(tmp_path)
    cfg = _render(tmp_path, num_users=inter.num_users - 1, num_items=inter.num_items - 1, epochs=3, adaptive=adaptive,
                  train_batch_size=64, embedding_dim=16, optimizer=optimizer, lr=lr, item_bias="true")
    exp_cfg = cfg.pop("experiment")
    exp = instantiate(exp_cfg, exp_config=lambda: cfg, dir=None, debug=False, seed=13, trackers_params={})
    state = exp.run()
    assert state is exp.trainer.engines["eval"].state
    tr = exp.trainer.engines["train"].state
    n_train = sum(len(v) for v in train_rows.values())
    assert tr.epoch == 3 and tr.iteration == 3 * ((n_train + 63) // 64)
    for k in ("loss", "bpr_loss", "l2_reg", "logits_diff"):
        assert np.isfinite(tr.metrics[k].item()), k
    np.testing.assert_allclose(tr.metrics["loss"].item(), tr.metrics["bpr_loss"].item() + tr.metrics["l2_reg"].item(), rtol=1e-4)
    # eval metrics of the last evaluation == oracle metrics of the final tables
    sd = exp._model.state_dict()
    ref = ref_bpr.RefModel(sd["logits_model._user_emb.weight"].cpu(), sd["logits_model._item_emb.weight"].cpu(),
                           sd["logits_model._item_bias"].cpu())
    users = sorted(test_rows)
    seen_pad = torch.nn.utils.rnn.pad_sequence([torch.as_tensor(train_rows[u]) for u in users], batch_first=True)
    logits = ref.eval_logits(torch.as_tensor(users), seen_pad)
    target = torch.zeros(len(users), inter.num_items)
    for r, u in enumerate(users):
        target[r, torch.as_tensor(test_rows[u])] = 1.0
    np.testing.assert_allclose(exp.metrics["ndcg@10"].item(), ref_bpr.ndcg_at_k(logits, target, 10).mean().item(), atol=1e-4)
    np.testing.assert_allclose(exp.metrics["recall@20"].item(), ref_bpr.recall_at_k(logits, target, 20).mean().item(), atol=1e-4)
    srt = torch.gather(target, 1, torch.argsort(-logits, dim=-1))[:, :5]
    np.testing.assert_allclose(exp.metrics["precision@5"].item(), (srt.sum(-1) / 5).mean().item(), atol=1e-4)
    auc = []
    for r in range(len(users)):  # RocAucManySlow restated (auc.py:149-166)
        pos, neg = logits[r][target[r] != 0], logits[r][target[r] == 0]
        auc.append((pos[:, None] > neg[None, :]).float().mean().item())
    np.testing.assert_allclose(exp.metrics["auc"].item(), np.mean(auc), atol=1e-4)
    assert 0.3 < exp.metrics["auc"].item() <= 1.0


def test_auc_and_seen_mask_kernels():
    from rbpr.engine import Context
    from revisit_bpr.metrics import RocAucManySlow
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    out = torch.randn(7, 3000, generator=g)
    tgt = (torch.rand(7, 3000, generator=g) < 0.4).float()  # > 1024 positives per row: several chunks
    tgt[2] = 0  # no positives -> NaN
    mask = (torch.rand(7, 3000, generator=g) < 0.8).float()
    for mk in (None, mask):
        got = RocAucManySlow().compute(out.to(dev), tgt.to(dev), None if mk is None else mk.to(dev)).cpu()
        for r in range(7):
            pos = out[r][tgt[r] != 0]
            neg = out[r][(tgt[r] == 0) & ((mk[r] != 0) if mk is not None else torch.ones(3000, dtype=torch.bool))]
            if pos.numel() == 0:
                assert got[r].isnan()
            else:
                np.testing.assert_allclose(got[r].item(), (pos[:, None] > neg[None, :]).float().mean().item(), rtol=1e-5)
    ctx = Context(dev)
    logits = torch.randn(5, 40, generator=g).to(dev)
    seen = torch.tensor([[3, 7, 0, 0], [1, 2, 3, 39], [0, 0, 0, 0], [5, 0, 0, 0], [9, 8, 7, 6]])
    exp = logits.clone().cpu().scatter_(-1, seen, -1e13)
    exp[:, 0] = -1e13
    ctx.mask_seen_padded(logits, seen)
    assert torch.equal(logits.cpu(), exp)


RQ1_CONFIG = """
num_users: &num_users {{ (num_users | int) + 1 }}
num_items: &num_items {{ (num_items | int) + 1 }}
epochs: 2
experiment:
  _target_: experiments.bpr.Experiment
  skip_seen: false
  metrics:
    auc: {_target_: revisit_bpr.metrics.RocAucOne}
datasets:
  train:
    _target_: torch.utils.data.DataLoader
    dataset: {_target_: revisit_bpr.datasets.jsonl.Iter, path: {{ dataset }}/rq1-train.jsonl}
    collate_fn: {_target_: revisit_bpr.datasets.jsonl.Collator, pad: [seen_items]}
    batch_size: 16
  eval:
    _target_: torch.utils.data.DataLoader
    dataset: {_target_: revisit_bpr.datasets.jsonl.Iter, path: {{ dataset }}/rq1-eval.jsonl}
    collate_fn: {_target_: experiments.bpr.dataset.OnePosCollator, num_items: *num_items}
    batch_size: 1
model:
  _target_: revisit_bpr.models.bpr.Model
  fuse_forward: true
  logits_model:
    _target_: revisit_bpr.models.bpr.MF
    user_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_users, embedding_dim: 16, padding_idx: 0}
    item_emb: {_target_: torch.nn.Embedding, num_embeddings: *num_items, embedding_dim: 16, padding_idx: 0}
optimizer: {_partial_: true, _target_: torch.optim.SGD, lr: 0.05}
"""


def test_rq1_protocol_config(tmp_path):
    """configs/RQ1/ours.yaml.j2 shape: jsonl.Iter + Collator(pad) train lines carrying their own
    seen_items, OnePosCollator eval (one positive vs every unseen item), RocAucOne, skip_seen false."""
    import jinja2
    from experiments._instantiate import instantiate
    inter, train_rows, test_rows = _write_dataset(tmp_path, n_users=40, n_items=50, seed=5)
    with open(tmp_path / "rq1-train.jsonl", "w") as f:
        for u, items in train_rows.items():
            for i in items:
                f.write(json.dumps({"user": u, "item": i, "seen_items": items}) + "\n")
    with open(tmp_path / "rq1-eval.jsonl", "w") as f:  # positive = index into the user's seen list
        for u, items in train_rows.items():
            f.write(json.dumps({"user": u, "item": len(items) - 1, "seen_items": items}) + "\n")
    cfg = yaml.safe_load(jinja2.Template(RQ1_CONFIG, undefined=jinja2.StrictUndefined).render(
        dataset=str(tmp_path), num_users=inter.num_users - 1, num_items=inter.num_items - 1))
    exp = instantiate(cfg.pop("experiment"), exp_config=lambda: cfg, dir=None, seed=13)
    exp.run()
    n_train = sum(len(v) for v in train_rows.values())
    tr = exp.trainer.engines["train"].state
    assert tr.epoch == 2 and tr.iteration == 2 * ((n_train + 15) // 16)
    auc = exp.metrics["auc"].item()
    # after training, the held positive (an item the user interacted with) outranks unseen items
    assert 0.55 < auc <= 1.0, auc
    # and equals the oracle on the final tables
    sd = exp._model.state_dict()
    ue, ie = sd["logits_model._user_emb.weight"].cpu(), sd["logits_model._item_emb.weight"].cpu()
    vals = []
    for u, items in train_rows.items():
        pos = items[-1]
        unseen = [i for i in range(1, inter.num_items) if i not in set(items)]
        s = ie @ ue[u]
        vals.append((s[pos] > s[unseen]).float().mean().item())
    np.testing.assert_allclose(auc, np.mean(vals), atol=1e-4)


@pytest.mark.parametrize("adaptive", [False, True])
def test_fast_train_extension_runs_whole_chunks_inside_the_library(tmp_path, adaptive):
    """`fast_train: true` (our extension key): the same reference-schema config, but the train loader
    hands over whole chunks of steps as triple ids and the library samples and trains them in one
    call per chunk.  Same bookkeeping surface (Trainer events, running losses, eval metrics)."""
    from experiments._instantiate import instantiate
    from oracle import ref_bpr
    inter, train_rows, test_rows = _write_dataset(tmp_path)
    cfg = _render(tmp_path, num_users=inter.num_users - 1, num_items=inter.num_items - 1, epochs=4, adaptive=adaptive,
                  train_batch_size=64, embedding_dim=16, optimizer="torch.optim.SGD", lr=0.05, item_bias="true")
    cfg["fast_train"] = True
    cfg["fast_steps_per_chunk"] = 4
    exp_cfg = cfg.pop("experiment")
    exp = instantiate(exp_cfg, exp_config=lambda: cfg, dir=None, debug=False, seed=13, trackers_params={})
    exp.run()
    n_train = sum(len(v) for v in train_rows.values())
    steps_per_epoch = (n_train + 63) // 64
    tr = exp.trainer.engines["train"].state
    assert tr.epoch == 4 and tr.iteration == 4 * ((steps_per_epoch + 3) // 4)  # iterations are chunks
    assert exp._model._opt_step == 4 * steps_per_epoch                        # ... steps are steps
    for k in ("loss", "bpr_loss", "l2_reg", "logits_diff"):
        assert np.isfinite(tr.metrics[k].item()), k
    assert tr.metrics["bpr_loss"].item() < 64 * np.log(2)  # per-step scale, and it learned something
    sd = exp._model.state_dict()
    ref = ref_bpr.RefModel(sd["logits_model._user_emb.weight"].cpu(), sd["logits_model._item_emb.weight"].cpu(),
                           sd["logits_model._item_bias"].cpu())
    users = sorted(test_rows)
    seen_pad = torch.nn.utils.rnn.pad_sequence([torch.as_tensor(train_rows[u]) for u in users], batch_first=True)
    logits = ref.eval_logits(torch.as_tensor(users), seen_pad)
    target = torch.zeros(len(users), inter.num_items)
    for r, u in enumerate(users):
        target[r, torch.as_tensor(test_rows[u])] = 1.0
    np.testing.assert_allclose(exp.metrics["ndcg@10"].item(), ref_bpr.ndcg_at_k(logits, target, 10).mean().item(), atol=1e-4)
    np.testing.assert_allclose(exp.metrics["recall@20"].item(), ref_bpr.recall_at_k(logits, target, 20).mean().item(), atol=1e-4)
    assert exp.metrics["auc"].item() > 0.5


@pytest.mark.parametrize("optimizer,lr,fast", [("torch.optim.SGD", 0.05, False), ("torch.optim.Adam", 0.01, False),
                                               ("torch.optim.Adam", 0.01, True)])
def test_experiment_checkpoints_and_resumes_from_dir(tmp_path, optimizer, lr, fast):
    """`dir` set: a checkpoint after every eval pass under <dir>/checkpoints (n_checkpoints kept,
    best copied to <dir>/best_iteration), and a second run on the same dir continues where the first
    stopped (reference exp.py:249-272, options.py:88-146): 2 epochs + resume to 4 == 4 epochs in one go
    (tables, optimizer step, sampler stream, shuffle order, early-stopping counters)."""
    from experiments._instantiate import instantiate
    data = tmp_path / "data"
    data.mkdir()
    inter, train_rows, _ = _write_dataset(data)

    def run(where, epochs):
        cfg = _render(data, num_users=inter.num_users - 1, num_items=inter.num_items - 1, epochs=epochs, adaptive=False,
                      train_batch_size=64, embedding_dim=16, optimizer=optimizer, lr=lr, item_bias="true")
        if fast:
            cfg["fast_train"], cfg["fast_steps_per_chunk"] = True, 3
        exp_cfg = cfg.pop("experiment")
        exp = instantiate(exp_cfg, exp_config=lambda: cfg, dir=where, n_checkpoints=2, debug=False, seed=13,
                          trackers_params={})
        exp.run()
        return exp

    full = run(tmp_path / "a", 4)
    ckpts = sorted(p.name for p in (tmp_path / "a" / "checkpoints").iterdir())
    assert ckpts == ["checkpoint_3", "checkpoint_4"]  # 5 eval passes (4 epoch starts + completion), 2 kept
    assert sorted(p.name for p in (tmp_path / "a" / "checkpoints" / "checkpoint_4").iterdir())[:3] == [
        "custom_checkpoint_0.pkl", "custom_checkpoint_1.pkl", "custom_checkpoint_2.pkl"]
    assert (tmp_path / "a" / "best_iteration" / "pytorch_model.bin").exists()

    first = run(tmp_path / "b", 2)
    it_per_epoch = first.trainer.engines["train"].state.epoch_length
    assert first.trainer.engines["train"].state.iteration == 2 * it_per_epoch
    rest = run(tmp_path / "b", 4)
    tr = rest.trainer.engines["train"].state
    assert tr.epoch == 4 and tr.iteration == 4 * it_per_epoch
    assert rest._model._opt_step == full._model._opt_step > 0
    assert rest._neg_calls == full._neg_calls
    a, b = full._model.state_dict(), rest._model.state_dict()
    for k in a:
        np.testing.assert_allclose(b[k].cpu().numpy(), a[k].cpu().numpy(), rtol=1e-4, atol=2e-5, err_msg=k)
    for k in ("ndcg@10", "recall@20", "precision@5", "auc"):
        np.testing.assert_allclose(float(rest.metrics[k]), float(full.metrics[k]), atol=2e-3, err_msg=k)
    np.testing.assert_allclose(tr.metrics["loss"].item(), full.trainer.engines["train"].state.metrics["loss"].item(),
                               rtol=1e-4)


STOCK_METRICS = """
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
"""


def test_eval_ranks_once_per_batch_for_the_stock_metric_set(tmp_path):
    """The 14 metrics every RQ2 config attaches (configs/RQ2/neg-sampling/ada-sampling-ml-20m.yaml.j2:13-55):
    the 13 ranking metrics come from ONE scoring + ranking call per eval batch (rbpr_score_metrics),
    only the AUC reads dense logits; values equal those of the metric classes fed dense tensors."""
    import jinja2
    from experiments._instantiate import instantiate
    from revisit_bpr import metrics as M
    inter, train_rows, test_rows = _write_dataset(tmp_path, n_users=150, n_items=260, seed=9)
    text = CONFIG[:CONFIG.index("  metrics:")] + "  metrics:" + STOCK_METRICS + CONFIG[CONFIG.index("datasets:"):]
    cfg = yaml.safe_load(jinja2.Template(text, undefined=jinja2.StrictUndefined).render(
        dataset=str(tmp_path), num_users=inter.num_users - 1, num_items=inter.num_items - 1, epochs=1, adaptive=False,
        train_batch_size=64, embedding_dim=16, optimizer="torch.optim.SGD", lr=0.05, item_bias="true"))
    # extra families in the same pass: MAP (both normalisations -> two passes), FBeta, linear-gain NDCG
    exp_cfg = cfg.pop("experiment")
    exp = instantiate(exp_cfg, exp_config=lambda: cfg, dir=None, debug=False, seed=13, trackers_params={})
    counts = {}
    orig_run = type(exp)._get_trainer

    def spy(self, *a, **k):
        tr = orig_run(self, *a, **k)
        eng = self._model.logits_model.engine()
        from experiments.trainer import Events
        tr.add_event("eval", Events.STARTED, lambda: counts.__setitem__("t0", eng.topk_launch_count()))
        tr.add_event("eval", Events.COMPLETED, lambda e: counts.__setitem__("last", (eng.topk_launch_count() - counts["t0"],
                                                                                    e.state.iteration)))
        return tr

    type(exp)._get_trainer = spy
    try:
        exp.run()
    finally:
        type(exp)._get_trainer = orig_run
    launches, batches = counts["last"]
    n_eval_users = len(test_rows)
    assert batches % ((n_eval_users + 31) // 32) == 0  # iteration counts accumulate over eval passes
    assert launches == (n_eval_users + 31) // 32, (launches, batches)  # ONE ranking launch per eval batch
    # same numbers as the metric classes on dense tensors (the reference's calling convention)
    dev = torch.device("cuda:0")
    model = exp._model
    model.eval()
    users = sorted(test_rows)
    eng = model.logits_model.engine()
    seen = torch.nn.utils.rnn.pad_sequence([torch.as_tensor(train_rows[u]) for u in users], batch_first=True)
    items = torch.arange(inter.num_items).unsqueeze(0).repeat(len(users), 1)
    with torch.no_grad():
        logits = model({"user": torch.as_tensor(users).to(dev), "item": items.to(dev)})["logits"]
    exp._sampler_ctx.mask_seen_padded(logits, seen)
    target = torch.zeros(len(users), inter.num_items)
    for r, u in enumerate(users):
        target[r, torch.as_tensor(test_rows[u])] = 1.0
    target = target.to(dev)
    for key, m in (("ndcg@100", M.NDCG(100)), ("recall@20", M.Recall(20)), ("precision@50", M.Precision(50)),
                   ("ndcg@5", M.NDCG(5)), ("recall@100", M.Recall(100))):
        np.testing.assert_allclose(exp.metrics[key].item(), m.compute(logits, target).mean().item(), atol=1e-6, err_msg=key)
    del eng


def test_fused_eval_output_families_match_dense_metric_classes():
    """rbpr_score_metrics vs the metric classes on the dense matrix, all families and both MAP
    normalisations, linear-gain NDCG and FBeta, through the update handler's grouping."""
    from types import SimpleNamespace
    from experiments.bpr.dataset import AllItemsCollator
    from experiments.options import _fused_update
    from revisit_bpr import metrics as M
    from revisit_bpr.models.bpr import MF, Model
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    U, I, D = 90, 333, 24
    model = Model(MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0), item_bias=True)).to(dev)
    with torch.no_grad():
        model.logits_model._item_bias.normal_(0, 0.01)
    model.eval()
    rng = np.random.default_rng(2)
    insts = []
    for u in range(1, 60):
        row = rng.permutation(np.arange(1, I))[: rng.integers(3, 40)]
        k = max(1, row.size // 4)
        insts.append({"user": u, "item": sorted(row[:k].tolist()), "seen_items": sorted(row[k:].tolist())})
    insts[5]["item"] = []  # a user without positives scores 0 and still counts
    batch = AllItemsCollator(I)(insts)
    for k, v in list(batch.items()):
        if torch.is_tensor(v):
            batch[k] = v.to(dev)
    out = model(batch)
    assert out.fused
    out.mask_seen(batch["seen_csr"])
    mk = lambda: {"ndcg@10": M.NDCG(10), "ndcgl@20": M.NDCG(20, "linear"), "recall@20": M.Recall(20),  # noqa: E731
                  "precision@7": M.Precision(7), "map@30": M.MAP(30), "mapu@30": M.MAP(30, normalized=False),
                  "f@15": M.FBeta(15, beta=0.5), "ndcg@500": M.NDCG(500), "auc": M.RocAucManySlow()}
    fused, dense = mk(), mk()
    eng = model.logits_model.engine()
    t0 = eng.topk_launch_count()
    state = SimpleNamespace(output=out, batch=batch, metrics={})
    rest = _fused_update(state, fused)
    assert set(rest) == {"ndcg@500", "auc"}          # k > 128 and AUC need the dense matrix
    assert eng.topk_launch_count() - t0 == 2          # two MAP normalisations -> two passes, nothing more
    assert out.fused                                   # ... and no (B, I) matrix was built for them
    logits, target = out["logits"], batch["target"]
    assert not out.fused and logits.shape == (len(insts), I) and target.shape == logits.shape
    for key, m in dense.items():
        if key in rest:
            continue
        m(logits, target)
        np.testing.assert_allclose(fused[key].get_metric().item(), m.get_metric().item(), atol=1e-6, err_msg=key)
        assert float(fused[key]._total_count) == len(insts)
