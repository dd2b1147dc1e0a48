"""Checkpoint / resume glue on CPU (SURVEY.md §8 f rank 3): the accelerate stand-in's save_state /
load_state layout and rotation, the engine stand-in's mid-epoch resume, and
attach_checkpointer / attach_checkpoint_loader / attach_preemptible around the Trainer
(reference experiments/options.py:88-146,188-219).  The model here is a pure-torch logits model
(test fixture) going through `Model`'s unfused path, so no GPU is involved; the same flow through
`experiments.bpr.Experiment` on the CUDA path is in tests/test_gpu_experiment.py."""
import copy
from pathlib import Path

import numpy as np
import pytest
import torch

from experiments._accel import Accelerator, ProjectConfiguration
from experiments._engine import Engine, Events
from experiments.options import (attach_checkpoint_loader, attach_checkpointer, attach_early_stopping,
                                 attach_metrics, attach_preemptible)
from experiments.trainer import Trainer
from revisit_bpr.models import BPR
from revisit_bpr.models.bpr import BaseLogitModel


class Dot(BaseLogitModel):
    """Test fixture: MF logits in plain torch (autograd), the shape the unfused Model path expects."""

    def __init__(self, users: int, items: int, dim: int) -> None:
        super().__init__()
        self.u = torch.nn.Embedding(users, dim)
        self.v = torch.nn.Embedding(items, dim)

    def forward(self, user, item, _other=None):
        return torch.einsum("bd,bnd->bn", self.u(user), self.v(item))

    def get_features(self):
        return {"user": self.u.weight, "item": self.v.weight}


def make_batches(n, B=8, users=20, items=15, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"user": torch.randint(1, users, (B,), generator=g), "item": torch.randint(1, items, (B, 1), generator=g),
             "neg": torch.randint(1, items, (B, 1), generator=g)} for _ in range(n)]


def build(tmp, total_limit=3, seed=0):
    torch.manual_seed(seed)
    model = BPR(Dot(20, 15, 6), reg_alphas={"all": 0.01})
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    acc = Accelerator("cpu", project_config=ProjectConfiguration(project_dir=str(tmp), automatic_checkpoint_naming=True,
                                                                  total_limit=total_limit))
    model, opt = acc.prepare(model, opt)
    return model, opt, acc


def test_save_state_layout_rotation_and_load(tmp_path):
    model, opt, acc = build(tmp_path, total_limit=2)

    class Counter:
        def __init__(self):
            self.n = 0

        def state_dict(self):
            return {"n": self.n}

        def load_state_dict(self, d):
            self.n = d["n"]

    c = Counter()
    acc.register_for_checkpointing(c)
    with pytest.raises(ValueError):
        acc.register_for_checkpointing(object())
    batch = make_batches(1)[0]
    saved = []
    for k in range(4):
        model.train()
        out = model(batch)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        c.n = k
        saved.append((acc.save_state(), copy.deepcopy(model.state_dict())))
    root = tmp_path / "checkpoints"
    assert sorted(p.name for p in root.iterdir()) == ["checkpoint_2", "checkpoint_3"]  # total_limit = 2
    assert acc.save_iteration == 4 and saved[-1][0] == str(root / "checkpoint_3")
    assert sorted(p.name for p in (root / "checkpoint_3").iterdir()) == [
        "custom_checkpoint_0.pkl", "optimizer.bin", "pytorch_model.bin", "random_states_0.pkl"]
    model2, opt2, acc2 = build(tmp_path, seed=5)
    c2 = Counter()
    acc2.register_for_checkpointing(c2)
    acc2.load_state()  # newest
    assert c2.n == 3
    for k, v in saved[-1][1].items():
        assert torch.equal(model2.state_dict()[k], v)
    assert opt2.state_dict()["state"][0]["step"] == 4
    acc2.load_state(str(root / "checkpoint_2"))
    assert c2.n == 2
    with pytest.raises(ValueError):
        acc2.load_state(str(root / "checkpoint_0"))
    acc2.project_configuration.iteration = 3
    with pytest.raises(ValueError, match="already exists"):
        acc2.save_state()


def test_engine_resumes_inside_an_epoch(tmp_path):
    data = list(range(10, 15))
    seen = []

    def step(engine, batch):
        seen.append((engine.state.epoch, batch))
        if engine.state.iteration == 7:
            engine.interrupt()

    e1 = Engine(step)
    e1.run(data, max_epochs=3)
    assert e1.state.iteration == 7 and len(seen) == 7
    e2 = Engine(lambda engine, batch: seen.append((engine.state.epoch, batch)))
    e2.load_state_dict(e1.state_dict())
    assert e2.state.epoch == 1
    acc = Accelerator("cpu")
    e2.add_event_handler(Events.STARTED, lambda eng: setattr(eng.state, "dataloader", acc.skip_first_batches(data, 2)))
    e2.add_event_handler(Events.EPOCH_COMPLETED, lambda eng: eng.set_data(data))
    e2.run(data, max_epochs=3)
    assert seen == [(ep, b) for ep in (1, 2, 3) for b in data]
    assert e2.state.iteration == 15 and e2.state.epoch == 3


def run_trainer(tmp, batches, epochs, stop_at=None, early=False):
    model, opt, acc = build(tmp)
    trainer = Trainer(model, opt, acc)
    es = attach_early_stopping(trainer, "loss", patience=50, direction="min") if early else None
    attach_preemptible(trainer, acc, min_seconds_between_saves=0)
    attach_metrics(trainer, acc, {})
    attach_checkpointer(trainer, acc, early_stopping=es)
    loaders = {"train": batches, "eval": batches[:2]}
    attach_checkpoint_loader(trainer, acc, loaders)
    if stop_at is not None:
        trainer.add_event("train", Events.ITERATION_COMPLETED,
                          lambda e: e.interrupt() if e.state.iteration == stop_at else None)
    root = Path(tmp) / "checkpoints"
    if root.is_dir() and any(root.iterdir()):
        acc.load_state()
        acc.project_configuration.iteration = max(int(p.name.rsplit("_", 1)[1]) for p in root.iterdir()) + 1
    losses = []
    trainer.add_event("train", Events.ITERATION_COMPLETED, lambda e: losses.append(e.state.output["loss"].item()))
    trainer.run(loaders, epochs=epochs)
    return model, trainer, losses


def test_interrupted_run_resumes_to_the_same_weights(tmp_path):
    batches = make_batches(6)
    ref_model, ref_trainer, ref_losses = run_trainer(tmp_path / "a", batches, epochs=3)
    assert len(ref_losses) == 18
    # 4 eval passes (3 epoch starts + completion) -> 4 checkpoints, 3 kept, best_iteration present
    assert sorted(p.name for p in (tmp_path / "a" / "checkpoints").iterdir()) == [
        "checkpoint_1", "checkpoint_2", "checkpoint_3"]
    assert (tmp_path / "a" / "best_iteration" / "pytorch_model.bin").exists()

    _, t1, first = run_trainer(tmp_path / "b", batches, epochs=3, stop_at=8)  # inside epoch 2
    assert len(first) == 8 and t1.engines["train"].state.was_interrupted
    model2, t2, rest = run_trainer(tmp_path / "b", batches, epochs=3)
    np.testing.assert_allclose(first + rest, ref_losses, rtol=1e-6)
    for k, v in ref_model.state_dict().items():
        np.testing.assert_allclose(model2.state_dict()[k].numpy(), v.numpy(), rtol=1e-6, atol=1e-7)
    st = t2.engines["train"].state
    assert st.iteration == 18 and st.epoch == 3 and not st.was_interrupted
    np.testing.assert_allclose(st.metrics["loss"].item(), np.mean(ref_losses[12:]), rtol=1e-6)


def test_epoch_boundary_checkpoint_resume_and_early_stopping_state(tmp_path):
    batches = make_batches(5, seed=3)
    ref_model, _, ref_losses = run_trainer(tmp_path / "a", batches, epochs=4, early=True)
    # a run killed after its second epoch: emulate by training 2 epochs, then starting again for 4
    _, t1, first = run_trainer(tmp_path / "b", batches, epochs=2, early=True)
    assert t1.engines["train"].state.iteration == 10
    model2, t2, rest = run_trainer(tmp_path / "b", batches, epochs=4, early=True)
    np.testing.assert_allclose(first + rest, ref_losses, rtol=1e-6)
    for k, v in ref_model.state_dict().items():
        np.testing.assert_allclose(model2.state_dict()[k].numpy(), v.numpy(), rtol=1e-6, atol=1e-7)
