"""Handlers of the reference's experiments/options.py that sit on the BPR scoring path:
`attach_metrics` (options.py:31-85: reset / update / reduce handlers feeding every configured
metric with `(output["logits"], batch["target"])`) and `attach_early_stopping` (options.py:166-185).
Checkpointing, trackers, progress bars and output savers are control-plane glue and out of scope
(SURVEY.md §2 #8)."""
from __future__ import annotations

from typing import Any, Callable

import torch

from experiments.trainer import Events, Trainer
from revisit_bpr.metrics import MaskedMetric, Metric


def attach_metrics(trainer: Trainer, accelerator: Any, metrics: dict[str, Metric] | None = None) -> None:
    if metrics is None:
        return

    def reset_handler(engine: Any) -> None:
        if engine.state.was_interrupted:
            return
        for m in metrics.values():
            m.reset()

    @torch.no_grad()
    def update_handler(engine: Any) -> None:
        state = engine.state
        if state.skip_metrics or "target" not in state.batch:
            return
        for key, m in metrics.items():
            kwargs = {"mask": state.batch.get("mask")} if isinstance(m, MaskedMetric) else {}
            m(state.output["logits"], state.batch["target"], **kwargs)
            state.metrics[key] = m.get_metric()

    def make_reducer() -> tuple[Callable, Callable]:
        done = {"v": False}

        def reduce_handler(engine: Any) -> None:
            if not done["v"]:
                engine.state.metrics = {k: accelerator.reduce(v, reduction="mean") if torch.is_tensor(v) else v
                                        for k, v in engine.state.metrics.items()}
            done["v"] = True

        def rearm() -> None:
            done["v"] = False

        return reduce_handler, rearm

    for name, eng in trainer.engines.items():
        eng.state.skip_metrics = False
        eng.state_dict_user_keys.append("metrics")
        reduce_handler, rearm = make_reducer()
        trainer.add_event(name, Events.EPOCH_STARTED, rearm)
        trainer.add_event(name, Events.EPOCH_STARTED, reset_handler)
        trainer.add_event(name, Events.ITERATION_COMPLETED, update_handler)
        trainer.add_event(name, Events.EPOCH_COMPLETED | Events.INTERRUPT, reduce_handler)


class EarlyStopping:
    """Stop the train engine when the eval metric has not improved for `patience` evaluations."""

    def __init__(self, patience: int, score_function: Callable[[Any], float], trainer_engine: Any) -> None:
        self.patience, self.score_function, self.trainer = patience, score_function, trainer_engine
        self.best_score: float | None = None
        self.counter = 0

    def __call__(self, engine: Any) -> None:
        score = self.score_function(engine)
        if self.best_score is None or score > self.best_score:
            self.best_score, self.counter = score, 0
            return
        self.counter += 1
        if self.counter >= self.patience:
            self.trainer.terminate()

    def state_dict(self) -> dict[str, Any]:
        return {"counter": self.counter, "best_score": self.best_score}

    def load_state_dict(self, d: dict[str, Any]) -> None:
        self.counter, self.best_score = d["counter"], d["best_score"]


def attach_early_stopping(trainer: Trainer, metric_name: str, patience: int, direction: str = "max") -> EarlyStopping:
    sign = 1.0 if direction == "max" else -1.0

    def score(engine: Any) -> float:
        v = engine.state.metrics[metric_name]
        return sign * (v.item() if torch.is_tensor(v) else float(v))

    handler = EarlyStopping(patience, score, trainer.engines["train"])
    trainer.add_event("eval", Events.COMPLETED, handler)
    return handler
