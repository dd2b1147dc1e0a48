"""Handlers of the reference's experiments/options.py around the BPR path:
`attach_metrics` (options.py:31-85: reset / update / reduce handlers feeding every configured
metric with `(output["logits"], batch["target"])`), `attach_early_stopping` (options.py:166-185)
and the checkpoint / resume glue — `attach_checkpointer` (options.py:88-113: a checkpoint after
every eval pass, copied to `best_iteration/` when the early-stopping counter is 0),
`attach_checkpoint_loader` (options.py:116-146: skip the batches a resumed engine already
consumed) and `attach_preemptible` (options.py:188-219: mark the engines interrupted and save on
INTERRUPT / an exception).  Trackers, progress bars and output savers are control plane and out of
scope."""
from __future__ import annotations

import shutil
import time
from pathlib import Path
from typing import Any, Callable, Iterable

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
        rest = _fused_update(state, metrics)
        for key, m in rest.items():
            kwargs = {"mask": state.batch.get("mask")} if isinstance(m, MaskedMetric) else {}
            m(state.output["logits"], state.batch["target"], **kwargs)
            state.metrics[key] = m.get_metric()

    def make_reducer() -> tuple[Callable, Callable]:
        done = {"v": False}

        def reduce_handler(engine: Any) -> None:
            if not done["v"]:
                engine.state.metrics = _reduce_metrics(engine.state.metrics, metrics, accelerator)
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


def _reduce_metrics(values: dict[str, Any], metrics: dict[str, Metric], accelerator: Any) -> dict[str, Any]:
    """Cross-rank reduction of an engine's metric dict (reference options.py:53-59 takes the mean of
    the per-rank means).  Ranking metrics are reduced exactly instead — every rank contributes its
    (sum of per-user values, number of users) and ONE all-reduce carries all of them; other tensor
    entries (running losses) ride along as (value, 1), i.e. the mean over ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return {k: accelerator.reduce(v, reduction="mean") if torch.is_tensor(v) else v for k, v in values.items()}
    from rbpr.parallel import reduce_sum_count
    keys = [k for k, v in values.items() if torch.is_tensor(v)]
    if not keys:
        return values
    dev = accelerator.device
    pairs = []
    for k in keys:
        m = metrics.get(k)
        total = getattr(m, "_total", getattr(m, "_total_auc", None)) if m is not None else None
        count = getattr(m, "_total_count", None) if m is not None else None
        if torch.is_tensor(total) and torch.is_tensor(count):
            pairs.append(torch.stack([total.to(dev, torch.float64), count.to(dev, torch.float64)]))
        else:
            pairs.append(torch.stack([values[k].detach().to(dev, torch.float64).reshape(()),
                                      torch.ones((), dtype=torch.float64, device=dev)]))
    means = reduce_sum_count(torch.stack(pairs))
    out = dict(values)
    for k, v in zip(keys, means):
        out[k] = v.to(values[k].dtype)
    return out


def _fused_update(state: Any, metrics: dict[str, Metric]) -> dict[str, Metric]:
    """Feed every metric that can be derived from ONE scoring + ranking pass (rbpr_score_metrics:
    NDCG, Recall, Precision, MAP, FBeta at all their cut-offs) when the eval output is the lazy
    all-items object; returns the metrics still to be updated the reference's way (from dense logits)."""
    out, batch = state.output, state.batch
    fused = getattr(out, "fused", False) and getattr(out, "masked", False) and dict.__contains__(batch, "target_csr")
    if not fused:
        return metrics
    eng = out._model.logits_model.engine()
    rest: dict[str, Metric] = {}
    groups: dict[bool, list[tuple[str, Metric, tuple[str, ...]]]] = {}
    for key, m in metrics.items():
        req = m.fused_request() if hasattr(m, "fused_request") else None
        if req is None:
            rest[key] = m
        else:  # one MAP normalisation per call: metrics asking for the other one form a second group
            groups.setdefault(bool(getattr(m, "_normalized", True)), []).append((key, m, req))
    if len(groups) == 2:  # fold the group without a MAP into the other one
        for flag in (True, False):
            if not any("map" in req for _, _, req in groups[flag]):
                groups[not flag].extend(groups.pop(flag))
                break
    for normalized, members in groups.items():
        ks = sorted({min(m._topk, eng.I) for _, m, _ in members})
        want = tuple(sorted({name for _, _, req in members for name in req}))
        for a in range(0, len(ks), 16):  # the kernel takes 16 cut-offs per pass
            part = ks[a:a + 16]
            res = out.ranking_metrics(batch["target_csr"], part, want, map_normalized=normalized)
            for key, m, _ in members:
                k = min(m._topk, eng.I)
                if k in part:
                    m.accumulate(m.fused_value(res, part.index(k)))
                    state.metrics[key] = m.get_metric()
    return rest


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


# ---- checkpoint / resume -------------------------------------------------------------------------
CHECKPOINTS_DIR = "checkpoints"      # reference experiments/settings.py:2-3
BEST_ITERATION_PATH = "best_iteration"


def _save(accelerator: Any) -> str:
    """accelerator.save_state() under automatic naming; a folder that already carries the next
    number (an aborted save) is stepped over instead of failing the run."""
    folder = Path(accelerator.project_dir) / CHECKPOINTS_DIR / f"checkpoint_{accelerator.save_iteration}"
    if folder.exists():
        accelerator.project_configuration.iteration += 1
    return accelerator.save_state()


def attach_checkpointer(trainer: Trainer, accelerator: Any, early_stopping: Any = None,
                        checkpoint_objects: Iterable[Any] | None = None) -> None:
    """Everything a resumed run needs goes through `accelerator.save_state`: model and optimizer
    (prepared), the engines' states, the early-stopping counters and `checkpoint_objects` (metrics,
    the hot path's step / sampler counters)."""
    for obj in ([] if early_stopping is None else [early_stopping]) + list(trainer.engines.values()) \
            + list(checkpoint_objects or []):
        accelerator.register_for_checkpointing(obj)

    def after_eval(engine: Any) -> None:
        engine.state.save_location = _save(accelerator)
        improved = early_stopping is None or early_stopping.counter == 0
        if improved and accelerator.is_local_main_process:
            shutil.copytree(engine.state.save_location, Path(accelerator.project_dir) / BEST_ITERATION_PATH,
                            dirs_exist_ok=True)

    trainer.add_event("eval", Events.COMPLETED, after_eval)


def attach_checkpoint_loader(trainer: Trainer, accelerator: Any, datasets: dict[str, Any]) -> None:
    def on_started(engine: Any) -> None:
        st = engine.state
        for k, v in st.metrics.items():  # restored from a checkpoint written on another device
            if torch.is_tensor(v):
                st.metrics[k] = v.to(accelerator.device)
        if st.was_interrupted:
            done = st.iteration % st.epoch_length if st.epoch_length else st.iteration
            st.dataloader = accelerator.skip_first_batches(datasets[st.name], done)

    def on_epoch_completed(engine: Any) -> None:
        if engine.state.was_interrupted:  # the shortened first epoch is over: back to the full loader
            engine.set_data(datasets[engine.state.name])
            engine.state.was_interrupted = False

    for name in trainer.engines:
        trainer.add_event(name, Events.STARTED, on_started)
        trainer.add_event(name, Events.EPOCH_COMPLETED, on_epoch_completed)


def attach_preemptible(trainer: Trainer, accelerator: Any, min_seconds_between_saves: int = 10) -> None:
    def on_stop(engine: Any) -> None:
        engine.state.was_interrupted = True
        if accelerator.project_dir is None:
            return
        last = Path(accelerator.project_dir) / CHECKPOINTS_DIR / f"checkpoint_{accelerator.save_iteration - 1}"
        if last.exists() and time.time() - last.stat().st_mtime < min_seconds_between_saves:
            return  # the regular checkpoint was written a moment ago
        _save(accelerator)

    def on_exception(engine: Any, exc: BaseException | None = None) -> None:
        on_stop(engine)
        if exc is not None:
            raise exc

    for name in trainer.engines:
        trainer.add_event(name, Events.INTERRUPT, on_stop)
        trainer.add_event(name, Events.EXCEPTION_RAISED, on_exception)
