"""`Trainer(model, optimizer, accelerator, custom_engines=None)` with the reference's surface
(experiments/trainer.py:19-143): `.engines{"train","eval"}`, `.add_event`, `.run(loaders, max_iters,
epochs) -> State`, `ModelEvents`, the same event order inside a step, eval on EPOCH_STARTED and
COMPLETED of the train engine, running-mean loss in `state.metrics["loss"]`.

The step itself is the CUDA path: `model(batch)` in train mode performs forward, backward and the
optimizer update in the fused kernels (revisit_bpr.models.bpr.Model), so the `accelerator.backward`,
`optimizer.step` and `optimizer.zero_grad` calls below do no work — they stay so that handlers
attached to OPTIMIZER_STARTED / OPTIMIZER_COMPLETED and non-fused models keep their meaning."""
from __future__ import annotations

import copy
from typing import Any, Callable

import torch

try:  # the real packages when present
    from ignite.engine import Engine, EventEnum, Events, State
except ImportError:  # this image: stand-ins with the API subset used here
    from experiments._engine import Engine, EventEnum, Events, State


class ModelEvents(EventEnum):
    FORWARD_STARTED = "forward_started"
    FORWARD_COMPLETED = "forward_completed"
    OPTIMIZER_STARTED = "optimizer_started"
    OPTIMIZER_COMPLETED = "optimizer_completed"


# counter attribute of engine.state that each custom event advances / filters on
_EVENT_COUNTER = {ModelEvents.FORWARD_STARTED: "forward_iteration", ModelEvents.FORWARD_COMPLETED: "forward_iteration",
                  ModelEvents.OPTIMIZER_STARTED: "optimizer_iteration",
                  ModelEvents.OPTIMIZER_COMPLETED: "optimizer_iteration"}
# engine.state attributes that belong to a checkpoint of the engine
_PERSISTED = ("name", "forward_iteration", "optimizer_iteration", "epoch_iteration", "was_interrupted")


class Trainer:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, accelerator: Any,
                 custom_engines: dict[str, str] | None = None) -> None:
        self.model, self.optimizer, self._accelerator = model, optimizer, accelerator
        fused = getattr(getattr(model, "module", model), "bind_optimizer", None)
        if fused is not None:  # the fused CUDA step stands in for this optimizer
            fused(optimizer)
        self.engines = {"train": Engine(self._train_step), "eval": Engine(self._eval_step)}
        for alias, source in (custom_engines or {}).items():
            self.engines[alias] = copy.deepcopy(self.engines[source])
        # handler order as in the reference: the eval pass first, then the per-engine bookkeeping
        self.add_event("train", Events.EPOCH_STARTED | Events.COMPLETED, self._run_eval)
        for label, engine in self.engines.items():
            self._prepare_engine(label, engine)

    def _prepare_engine(self, label: str, engine: Engine) -> None:
        engine.register_events(*ModelEvents, event_to_attr=_EVENT_COUNTER)
        st = engine.state
        st.name, st.was_interrupted, st.epoch_iteration = label, False, 0
        engine.state_dict_user_keys.extend(_PERSISTED)
        engine.add_event_handler(Events.EPOCH_STARTED, self._on_epoch_started)
        engine.add_event_handler(Events.ITERATION_COMPLETED, self._on_iteration_completed)

    def add_event(self, engine: str, event_name: Any, handler: Callable, *args: Any, **kwargs: Any) -> None:
        self.engines[engine].add_event_handler(event_name, handler, *args, **kwargs)

    def run(self, loaders: dict[str, Any], max_iters: dict[str, int] | None = None,
            epochs: int | None = None) -> State:
        self._loaders, self._max_iters = loaders, dict(max_iters or {})
        self.engines["train"].run(loaders["train"], max_epochs=epochs, epoch_length=self._max_iters.get("train"))
        return self.engines["eval" if "eval" in loaders else "train"].state

    # ---- one step --------------------------------------------------------------------------------
    def _forward(self, engine: Engine, batch: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        engine.state.forward_iteration += 1
        engine.fire_event(ModelEvents.FORWARD_STARTED)
        engine.state.output = self.model(batch)
        engine.fire_event(ModelEvents.FORWARD_COMPLETED)
        return engine.state.output

    def _train_step(self, engine: Engine, batch: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        self.model.train()
        with self._accelerator.accumulate(self.model):
            out = self._forward(engine, batch)  # fused model: loss, gradients AND the update
            if "loss" in out:
                self._accelerator.backward(out["loss"])
                engine.state.optimizer_iteration += 1
                engine.fire_event(ModelEvents.OPTIMIZER_STARTED)
                self.optimizer.step()
                engine.fire_event(ModelEvents.OPTIMIZER_COMPLETED)
                self.optimizer.zero_grad()
                engine.state.metrics["_loss"] += out["loss"].detach()
        return out

    @torch.no_grad()
    def _eval_step(self, engine: Engine, batch: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        self.model.eval()
        out = self._forward(engine, batch)
        if "loss" in out:
            engine.state.metrics["_loss"] += out["loss"].detach()
        return out

    # ---- bookkeeping -----------------------------------------------------------------------------
    def _run_eval(self) -> None:
        # an eval pass runs before every train epoch and once after training; when a resumed run was
        # interrupted inside the TRAIN engine its eval pass had already finished
        if self.engines["train"].state.was_interrupted and not self.engines["eval"].state.was_interrupted:
            return
        if (loader := self._loaders.get("eval")) is not None:
            self.engines["eval"].run(loader, epoch_length=self._max_iters.get("eval"))

    def _on_epoch_started(self, engine: Engine) -> None:
        if not engine.state.was_interrupted:
            engine.state.epoch_iteration = 0
            engine.state.metrics["_loss"] = torch.tensor(0.0, device=self._accelerator.device)

    def _on_iteration_completed(self, engine: Engine) -> None:
        st = engine.state
        st.epoch_iteration += 1
        st.metrics["loss"] = st.metrics["_loss"] / st.epoch_iteration
