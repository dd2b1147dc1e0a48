"""`Trainer(model, optimizer, accelerator, custom_engines=None)` with the reference's surface
(experiments/trainer.py:19-143): `.engines{"train","eval"}`, `.add_event`, `.run(loaders, max_iters,
epochs) -> State`, `ModelEvents`, the same event order inside a step, eval on EPOCH_STARTED and
COMPLETED of the train engine, running-mean loss in `state.metrics["loss"]`.

The step itself is the CUDA path: `model(batch)` in train mode performs forward, backward and the
optimizer update in the fused kernels (revisit_bpr.models.bpr.Model), so `accelerator.backward`,
`optimizer.step` and `optimizer.zero_grad` below do no work — they stay so that handlers attached
to OPTIMIZER_STARTED / OPTIMIZER_COMPLETED and non-fused models keep their meaning."""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Callable

import torch

try:  # the real packages when present
    from ignite.engine import Engine, EventEnum, Events, State
except ImportError:  # this image: stand-ins with the API subset used here
    from experiments._engine import Engine, EventEnum, Events, State

_COUNTERS = ("name", "forward_iteration", "optimizer_iteration", "epoch_iteration", "was_interrupted")


class ModelEvents(EventEnum):
    FORWARD_STARTED = "forward_started"
    FORWARD_COMPLETED = "forward_completed"
    OPTIMIZER_STARTED = "optimizer_started"
    OPTIMIZER_COMPLETED = "optimizer_completed"


class Trainer:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, accelerator: Any,
                 custom_engines: dict[str, str] | None = None) -> None:
        self.model = model
        self.optimizer = optimizer
        self._accelerator = accelerator
        bind = getattr(getattr(model, "module", model), "bind_optimizer", None)
        if bind is not None:  # the fused CUDA step stands in for this optimizer
            bind(optimizer)
        self.engines = {"train": Engine(self._train_step), "eval": Engine(self._eval_step)}
        for name, base in (custom_engines or {}).items():
            self.engines[name] = deepcopy(self.engines[base])
        attr = {ModelEvents.FORWARD_STARTED: "forward_iteration", ModelEvents.FORWARD_COMPLETED: "forward_iteration",
                ModelEvents.OPTIMIZER_STARTED: "optimizer_iteration",
                ModelEvents.OPTIMIZER_COMPLETED: "optimizer_iteration"}
        for name, eng in self.engines.items():
            eng.register_events(*ModelEvents, event_to_attr=attr)
            eng.state.name = name
            eng.state.was_interrupted = False
            eng.state.epoch_iteration = 0
            eng.state_dict_user_keys.extend(_COUNTERS)
        self.add_event("train", Events.EPOCH_STARTED | Events.COMPLETED, self._run_eval)
        for name in self.engines:
            self.add_event(name, Events.EPOCH_STARTED, self._reset_epoch)
            self.add_event(name, Events.ITERATION_COMPLETED, self._count_iteration)
            self.add_event(name, Events.ITERATION_COMPLETED, self._mean_loss)

    def add_event(self, engine: str, event_name: Any, handler: Callable, *args: Any, **kwargs: Any) -> None:
        self.engines[engine].add_event_handler(event_name, handler, *args, **kwargs)

    def run(self, loaders: dict[str, Any], max_iters: dict[str, int] | None = None,
            epochs: int | None = None) -> State:
        self._loaders = loaders
        self._max_iters = max_iters or {}
        self.engines["train"].run(loaders["train"], epoch_length=self._max_iters.get("train"), max_epochs=epochs)
        return self.engines["eval" if "eval" in loaders else "train"].state

    # ---- steps ---------------------------------------------------------------------------------
    def _train_step(self, engine: Engine, batch: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        self.model.train()
        state = engine.state
        with self._accelerator.accumulate(self.model):
            state.forward_iteration += 1
            engine.fire_event(ModelEvents.FORWARD_STARTED)
            output = state.output = self.model(batch)  # fused: loss, gradients AND update
            engine.fire_event(ModelEvents.FORWARD_COMPLETED)
            if "loss" not in output:
                return output
            self._accelerator.backward(output["loss"])
            state.optimizer_iteration += 1
            engine.fire_event(ModelEvents.OPTIMIZER_STARTED)
            self.optimizer.step()
            engine.fire_event(ModelEvents.OPTIMIZER_COMPLETED)
            self.optimizer.zero_grad()
            state.metrics["_loss"] += output["loss"].detach()
        return output

    @torch.no_grad()
    def _eval_step(self, engine: Engine, batch: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        self.model.eval()
        state = engine.state
        state.forward_iteration += 1
        engine.fire_event(ModelEvents.FORWARD_STARTED)
        output = state.output = self.model(batch)
        engine.fire_event(ModelEvents.FORWARD_COMPLETED)
        if "loss" in output:
            state.metrics["_loss"] += output["loss"].detach()
        return output

    # ---- bookkeeping handlers --------------------------------------------------------------------
    def _run_eval(self) -> None:
        train, ev = self.engines["train"].state, self.engines["eval"].state
        if train.was_interrupted and not ev.was_interrupted:
            return  # resumed after an interruption that hit the train engine: eval already ran
        loader = self._loaders.get("eval")
        if loader is not None:
            self.engines["eval"].run(loader, epoch_length=self._max_iters.get("eval"))

    def _reset_epoch(self, engine: Engine) -> None:
        if engine.state.was_interrupted:
            return
        engine.state.metrics["_loss"] = torch.tensor(0.0, device=self._accelerator.device)
        engine.state.epoch_iteration = 0

    def _count_iteration(self, engine: Engine) -> None:
        engine.state.epoch_iteration += 1

    def _mean_loss(self, engine: Engine) -> None:
        m = engine.state.metrics
        m["loss"] = m["_loss"] / engine.state.epoch_iteration
