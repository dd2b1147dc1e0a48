"""Minimal event-driven loop with the subset of pytorch-ignite's `Engine` API that
experiments.trainer.Trainer and the BPR experiment hooks use (ignite itself is preferred when
importable): Events with `(every=N)` filters and `|`, custom EventEnum registration with
event_to_attr counters, State, add_event_handler / fire_event / run / interrupt / terminate."""
from __future__ import annotations

import inspect
from enum import Enum
from typing import Any, Callable, Iterable


class _Filtered:
    def __init__(self, event: Any, every: int) -> None:
        self.event, self.every = event, int(every)

    def __or__(self, other: Any) -> "EventsList":
        return EventsList([self]) | other


class EventsList:
    def __init__(self, events: list) -> None:
        self.events = list(events)

    def __or__(self, other: Any) -> "EventsList":
        more = other.events if isinstance(other, EventsList) else [other]
        return EventsList(self.events + more)

    def __iter__(self):
        return iter(self.events)


class EventEnum(Enum):
    def __call__(self, every: int | None = None) -> Any:
        return self if every is None else _Filtered(self, every)

    def __or__(self, other: Any) -> EventsList:
        return EventsList([self]) | other


class Events(EventEnum):
    STARTED = "started"
    EPOCH_STARTED = "epoch_started"
    GET_BATCH_STARTED = "get_batch_started"
    GET_BATCH_COMPLETED = "get_batch_completed"
    ITERATION_STARTED = "iteration_started"
    ITERATION_COMPLETED = "iteration_completed"
    EPOCH_COMPLETED = "epoch_completed"
    COMPLETED = "completed"
    INTERRUPT = "interrupt"
    TERMINATE = "terminate"
    EXCEPTION_RAISED = "exception_raised"


_DEFAULT_ATTR = {
    Events.STARTED: "epoch", Events.EPOCH_STARTED: "epoch", Events.EPOCH_COMPLETED: "epoch",
    Events.COMPLETED: "epoch", Events.GET_BATCH_STARTED: "iteration", Events.GET_BATCH_COMPLETED: "iteration",
    Events.ITERATION_STARTED: "iteration", Events.ITERATION_COMPLETED: "iteration",
}


class State:
    def __init__(self) -> None:
        self.iteration = 0
        self.epoch = 0
        self.epoch_length: int | None = None
        self.max_epochs: int | None = None
        self.output: Any = None
        self.batch: Any = None
        self.metrics: dict[str, Any] = {}
        self.dataloader: Any = None
        self.event_to_attr: dict[Any, str] = dict(_DEFAULT_ATTR)

    def get_event_attrib_value(self, event: Any) -> int:
        return getattr(self, self.event_to_attr[event])


class Engine:
    def __init__(self, process_function: Callable[["Engine", Any], Any]) -> None:
        self._process = process_function
        self._handlers: dict[Any, list[tuple[Callable, tuple, dict, int | None]]] = {}
        self.state = State()
        self.state_dict_user_keys: list[str] = []
        self.should_terminate = False
        self.should_interrupt = False
        self.should_terminate_epoch = False
        self._iter: Any = None

    # ---- events ----
    def register_events(self, *events: Any, event_to_attr: dict[Any, str] | None = None) -> None:
        for e in events:
            self._handlers.setdefault(e, [])
            if event_to_attr and e in event_to_attr:
                self.state.event_to_attr[e] = event_to_attr[e]
                if not hasattr(self.state, event_to_attr[e]):
                    setattr(self.state, event_to_attr[e], 0)

    def add_event_handler(self, event_name: Any, handler: Callable, *args: Any, **kwargs: Any) -> None:
        events: Iterable = event_name if isinstance(event_name, EventsList) else [event_name]
        for e in events:
            every = None
            if isinstance(e, _Filtered):
                e, every = e.event, e.every
            self._handlers.setdefault(e, []).append((handler, args, kwargs, every))

    def on(self, event_name: Any, *args: Any, **kwargs: Any) -> Callable:
        def deco(fn: Callable) -> Callable:
            self.add_event_handler(event_name, fn, *args, **kwargs)
            return fn
        return deco

    def fire_event(self, event: Any, *event_args: Any) -> None:
        """Handlers are called as handler(engine, *event_args, *args, **kwargs), like ignite's."""
        for handler, args, kwargs, every in list(self._handlers.get(event, [])):
            args = tuple(event_args) + tuple(args)
            if every is not None:
                count = self.state.get_event_attrib_value(event) if event in self.state.event_to_attr else 0
                if count % every != 0:
                    continue
            try:  # ignite lets a handler omit the leading `engine` argument
                inspect.signature(handler).bind(self, *args, **kwargs)
                takes_engine = True
            except TypeError:
                takes_engine = False
            if takes_engine:
                handler(self, *args, **kwargs)
            else:
                handler(*args, **kwargs)

    # ---- control ----
    def terminate(self) -> None:
        self.should_terminate = True

    def interrupt(self) -> None:
        self.should_interrupt = True

    def terminate_epoch(self) -> None:
        self.should_terminate_epoch = True

    def set_data(self, data: Iterable) -> None:
        """Replace the data of a running engine; the next batch comes from the new iterable."""
        self.state.dataloader = data
        self._iter = iter(data)

    def state_dict(self) -> dict[str, Any]:
        d = {"epoch_length": self.state.epoch_length, "max_epochs": self.state.max_epochs,
             "iteration": self.state.iteration}
        d.update({k: getattr(self.state, k, None) for k in self.state_dict_user_keys})
        return d

    def load_state_dict(self, d: dict[str, Any]) -> None:
        self.state.iteration = d.get("iteration", 0)
        self.state.epoch_length = d.get("epoch_length")
        self.state.max_epochs = d.get("max_epochs")
        if self.state.epoch_length:
            self.state.epoch = self.state.iteration // self.state.epoch_length
        for k in self.state_dict_user_keys:
            if k in d:
                setattr(self.state, k, d[k])

    def run(self, data: Iterable, max_epochs: int | None = None, epoch_length: int | None = None) -> State:
        st = self.state
        st.dataloader = data
        st.max_epochs = max_epochs if max_epochs is not None else (st.max_epochs or 1)
        if epoch_length is None:
            try:  # DataLoader over an IterableDataset has __len__ but raises TypeError
                epoch_length = len(data)
            except TypeError:
                epoch_length = None
        st.epoch_length = epoch_length
        if st.epoch >= st.max_epochs:  # a finished engine restarts from scratch (ignite semantics)
            st.epoch, st.iteration = 0, 0
        self.should_terminate = self.should_interrupt = False
        # a state restored in the middle of an epoch (load_state_dict) resumes inside that epoch:
        # only the remaining iterations run (handlers on STARTED may swap in a loader that skips
        # the batches already consumed, ignite semantics)
        resume_at = st.iteration % st.epoch_length if (st.epoch_length and st.iteration > 0) else 0
        try:
            self.fire_event(Events.STARTED)
            while st.epoch < st.max_epochs and not self.should_terminate:
                st.epoch += 1
                self.fire_event(Events.EPOCH_STARTED)
                self._iter, n_in_epoch, resume_at = iter(st.dataloader), resume_at, 0
                self.should_terminate_epoch = False
                while ((st.epoch_length is None or n_in_epoch < st.epoch_length) and not self.should_terminate
                       and not self.should_terminate_epoch):
                    self.fire_event(Events.GET_BATCH_STARTED)
                    try:
                        st.batch = next(self._iter)
                    except StopIteration:
                        if st.epoch_length is None or n_in_epoch == 0:
                            break
                        self._iter = iter(st.dataloader)
                        st.batch = next(self._iter)
                    st.iteration += 1
                    n_in_epoch += 1
                    self.fire_event(Events.GET_BATCH_COMPLETED)
                    self.fire_event(Events.ITERATION_STARTED)
                    st.output = self._process(self, st.batch)
                    self.fire_event(Events.ITERATION_COMPLETED)
                    if self.should_interrupt:
                        self.fire_event(Events.INTERRUPT)
                        return st
                self.fire_event(Events.EPOCH_COMPLETED)
            self.fire_event(Events.COMPLETED)
        except BaseException as exc:
            if self._handlers.get(Events.EXCEPTION_RAISED):
                self.fire_event(Events.EXCEPTION_RAISED, exc)
            raise
        return st
