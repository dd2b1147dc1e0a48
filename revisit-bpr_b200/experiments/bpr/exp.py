"""`experiments.bpr.Experiment` (reference experiments/bpr/exp.py:44-405) — the hook wiring that
puts the BPR hot path inside the Trainer, with the reference's constructor signature so the jinja
configs' `experiment:` block instantiates unchanged:

  * GET_BATCH_COMPLETED on the train engine: `_train_batch` writes `batch["neg"]` (exp.py:356-367)
    from the device samplers — static (uniform, or popularity-weighted when the config supplies
    `item_counts` and `neg_sampling_alpha`, exp.py:85-91,282-293) or adaptive (exp.py:295-342) with
    its statistics refreshed every int(I·ln I / batch) iterations (exp.py:194-207);
  * FORWARD_COMPLETED on the eval engine: `_remove_seen_items` (exp.py:369-374);
  * ITERATION_COMPLETED: running means of bpr_loss / l2_reg / logits_diff (exp.py:383-405) and
    the configured metrics (`attach_metrics`).

Model, optimizer and loaders are built from the config with `instantiate` (hydra's when installed).
With `dir` set, a checkpoint is written after every eval pass under `<dir>/checkpoints/`
(`n_checkpoints` kept, the best one copied to `<dir>/best_iteration/`) and a run started on a `dir`
that already holds checkpoints resumes from the newest loadable one (exp.py:249-272,
options.py:88-146); the hot path's own counters (optimizer step, sampler stream position) travel
with it.  Trackers, progress bars, output savers and S3 sync of the reference are control plane and
are not reproduced; the corresponding constructor arguments are accepted and ignored.
"""
from __future__ import annotations

import json
import math
import random
import re
import shutil
import warnings
from pathlib import Path
from typing import Any, Callable, Literal

import numpy as np
import torch

from experiments.options import (CHECKPOINTS_DIR, attach_checkpoint_loader, attach_checkpointer,
                                 attach_early_stopping, attach_metrics, attach_preemptible)
from experiments.trainer import Events, ModelEvents, Trainer
from rbpr import native
from rbpr.engine import Context
from revisit_bpr.metrics import Metric

try:
    from hydra.utils import instantiate
except ImportError:
    from experiments._instantiate import instantiate

try:
    from accelerate import Accelerator
    from accelerate.utils import ProjectConfiguration
except ImportError:
    from experiments._accel import Accelerator, ProjectConfiguration


class _HotPathCounters:
    """What a resumed run needs beyond model / optimizer / engine state: the fused step's optimizer
    step number (plain SGD keeps no state that carries it) and the position of the counter-based
    negative sampler's stream.  Registered with `accelerator.register_for_checkpointing`."""

    def __init__(self, exp: "BPRExperiment") -> None:
        self._exp = exp

    def _generators(self) -> dict[str, torch.Generator]:
        """The loaders' own shuffle generators (exp.py:111-115 hands each DataLoader one): they are not
        part of the global RNG state, and the next epoch's permutation comes from them."""
        out = {}
        for key, loader in self._exp._datasets.items():
            if isinstance(g := getattr(loader, "generator", None), torch.Generator):
                out[key] = g
        return out

    def state_dict(self) -> dict[str, Any]:
        e = self._exp
        return {"opt_step": int(getattr(e._model, "_opt_step", 0)), "neg_seed": int(e._neg_seed),
                "neg_calls": int(e._neg_calls),
                "loader_rng": {k: g.get_state() for k, g in self._generators().items()}}

    def load_state_dict(self, d: dict[str, Any]) -> None:
        e = self._exp
        e._neg_seed, e._neg_calls = int(d["neg_seed"]), int(d["neg_calls"])
        if hasattr(e._model, "restore_step"):
            e._model.restore_step(int(d["opt_step"]))
        for k, g in self._generators().items():
            if k in d.get("loader_rng", {}):
                g.set_state(d["loader_rng"][k].cpu())


class BPRExperiment:
    def __init__(
        self,
        exp_config: dict[str, Any] | Callable[[], dict[str, Any]],
        dir: Path | None = None,  # noqa: A002
        n_checkpoints: int = 2,
        mixed_precision: str | None = None,
        datasets_key: str = "datasets",
        metrics: dict[str, Metric] | None = None,
        trackers_params: dict[str, Any] | None = None,
        events: dict[str, list[tuple[Any, Callable]]] | None = None,
        seed: int = 13,
        debug: bool = False,
        skip_seen: bool = True,
        save_logits: bool = False,
        save_user_metrics: bool = False,
        log_momentum: bool = False,
        early_stopping_metric: str | None = None,
        early_stopping_patience: int = 200,
        early_stopping_direction: Literal["min", "max"] = "max",
        neg_sampling_alpha: float = 0.0,
        adaptive_sampling_prob: float | None = None,
    ) -> None:
        self._config = exp_config if isinstance(exp_config, dict) else exp_config()
        self._dir = Path(dir) if dir is not None else None
        self._n_checkpoints = n_checkpoints
        self._seed = seed
        self._debug = debug
        self._skip_seen = skip_seen
        self._early_stopping_metric = early_stopping_metric
        self._early_stopping_patience = early_stopping_patience
        self._early_stopping_direction = early_stopping_direction
        self._datasets_key = datasets_key
        self._metrics = metrics or {}
        self._events = events or {}
        self._adaptive_sampling_prob = adaptive_sampling_prob
        if mixed_precision not in (None, "no"):
            raise NotImplementedError("the CUDA BPR path is fp32 only (the reference configs use no mixed precision)")
        del trackers_params, save_logits, save_user_metrics, log_momentum  # control plane
        # popularity weights count^alpha (exp.py:85-91); all ones = uniform
        self._item_counts = torch.ones(self._config["num_items"], dtype=torch.float32)
        self._weighted = False
        if (path := self._config[datasets_key].pop("item_counts", None)) is not None:
            with open(path, "r", encoding="utf-8") as fh:
                for rec in map(json.loads, fh):
                    self._item_counts[rec["item"]] = float(rec["count"]) ** neg_sampling_alpha
            self._weighted = bool((self._item_counts != 1).any())

    @property
    def metrics(self) -> dict[str, Any]:
        return self._state.metrics

    @property
    def _adaptive(self) -> bool:
        return self._adaptive_sampling_prob is not None and isinstance(self._adaptive_sampling_prob, float)

    # ---- run -------------------------------------------------------------------------------------
    def run(self) -> Any:
        self._accelerator = self._get_accelerator()
        dev = self._accelerator.device
        self._model = instantiate(self._config["model"])
        self._optimizer = instantiate(self._config["optimizer"])(self._model.parameters())
        self._model, self._optimizer = self._accelerator.prepare(self._model, self._optimizer)
        loaders_cfg = self._config[self._datasets_key]
        max_iters = {k: d.pop("max_iters", None) for k, d in loaders_cfg.items()}
        self._datasets = {key: instantiate(cfg, generator=torch.Generator().manual_seed(self._seed))
                          for key, cfg in loaders_cfg.items()}
        for loader in self._datasets.values():
            if hasattr(loader.dataset, "collate_fn"):
                loader.collate_fn = loader.dataset.collate_fn
        # Our extension: `fast_train: true` at the top level of the config replaces the B-sized train
        # batches by whole chunks of steps run inside the library (device sampling included).
        self._fast = bool(self._config.get("fast_train", False))
        self._world, self._rank = self._process_group()
        if self._world > 1:
            self._enable_data_parallel()
        if self._fast:
            self._enable_fast_train(dev)
        self.trainer = self._get_trainer(self._model, self._optimizer, self._datasets)
        # counter-based device sampler: seed + iteration, like the reference's reseeded generator
        # (ranks draw from disjoint streams)
        self._neg_seed = self._seed + self.trainer.engines["train"].state.iteration + 1000003 * self._rank
        self._neg_calls = 0
        self._sampler_ctx = Context(dev)
        if self._weighted:
            self._sampler_ctx.bind_item_weights(self._item_counts)
        self._load_checkpoint_if_needed()
        if self._adaptive and not self._fast:
            self._update_adaptive_stats()
        self._state = self.trainer.run(self._datasets, max_iters=max_iters, epochs=self._config["epochs"])
        self._accelerator.wait_for_everyone()
        self._accelerator.end_training()
        return self._state

    def _get_accelerator(self) -> Any:
        accelerator = Accelerator()
        if self._dir is not None:
            accelerator.project_configuration = ProjectConfiguration(
                project_dir=str(self._dir), automatic_checkpoint_naming=True, total_limit=self._n_checkpoints)
        self._seed_everything()
        for m in self._metrics.values():
            m.set_accelerator(accelerator)
        return accelerator

    def _load_checkpoint_if_needed(self) -> None:
        """Resume from the newest checkpoint under `<dir>/checkpoints` that loads; one that does not
        (a save cut short) is deleted and the one before it is tried (exp.py:249-272)."""
        if self._dir is None or not (root := self._dir / CHECKPOINTS_DIR).is_dir():
            return
        numbered = sorted((int(m.group(1)), d) for d in root.iterdir()
                          if d.is_dir() and (m := re.search(r"(\d+)$", d.name)))
        while numbered:
            num, folder = numbered.pop()
            try:
                self._accelerator.load_state()
            except Exception as exc:  # noqa: BLE001
                warnings.warn(f"checkpoint {folder} does not load ({exc!r}): removed, trying the one before",
                              stacklevel=2)
                shutil.rmtree(folder, ignore_errors=True)
                continue
            self._accelerator.project_configuration.iteration = num + 1  # accelerate does not restore it
            return

    def clean(self) -> None:
        # native contexts first, at the same point of the program on every rank
        eng = getattr(getattr(getattr(self, "_model", None), "logits_model", None), "_engine", None)
        if eng is not None:
            torch.cuda.synchronize()
            eng.close()
            self._model.logits_model._engine = None
        self._accelerator.free_memory()
        del self._accelerator, self.trainer

    @staticmethod
    def _process_group() -> tuple[int, int]:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def _enable_data_parallel(self) -> None:
        """One process per GPU (torchrun, or the reference's Distributed(...) + launcher.DDP,
        experiments/decorator.py:30-54, launcher.py:35-73): users are sharded by owner, so this rank's
        train loader only yields triples of ITS users (same number of batches on every rank), the
        model joins the library's communicator (one all-reduce of the dense item gradient per step),
        eval loaders hand every world-th batch to this rank, and the owner-sharded user rows are
        gathered before every eval pass / checkpoint."""
        from experiments.bpr.dataset import SparseSamplingInMemoryWithCollator
        from rbpr.parallel import OwnerBatchSampler, RoundRobinBatches, shard_bounds
        loader = self._datasets["train"]
        ds = loader.dataset
        if not isinstance(ds, SparseSamplingInMemoryWithCollator):
            raise NotImplementedError("data-parallel BPR needs the SparseSamplingInMemoryWithCollator train dataset "
                                      "(its CSR defines the owner blocks)")
        indptr, _ = ds.csr()
        self._user_cuts = shard_bounds(indptr, self._world)
        self._model.enable_data_parallel(self._user_cuts)
        if not self._fast:
            sampler = OwnerBatchSampler(indptr, self._world, self._rank, loader.batch_size,
                                        generator=torch.Generator().manual_seed(self._seed + self._rank))
            sharded = torch.utils.data.DataLoader(ds, batch_sampler=sampler, collate_fn=ds.collate_fn,
                                                  num_workers=loader.num_workers)
            sharded.total_batch_size = loader.batch_size
            self._datasets["train"] = sharded
        for key in [k for k in self._datasets if k != "train"]:
            self._datasets[key] = RoundRobinBatches(self._datasets[key], self._world, self._rank)

    def _enable_fast_train(self, dev: torch.device) -> None:
        from experiments.bpr.dataset import EpochChunks, SparseSamplingInMemoryWithCollator
        loader = self._datasets["train"]
        ds = loader.dataset
        if not isinstance(ds, SparseSamplingInMemoryWithCollator):
            raise NotImplementedError("fast_train needs the SparseSamplingInMemoryWithCollator train dataset")
        bs = loader.batch_size
        indptr, indices = ds.csr()
        kind, every = native.SAMPLER_UNIFORM, 0
        if self._adaptive:
            kind = native.SAMPLER_ADAPTIVE
            every = max(1, int(self._config["num_items"] * math.log(self._config["num_items"]) / bs))
        elif self._weighted:
            kind = native.SAMPLER_WEIGHTED
        self._model.bind_interactions(torch.from_numpy(indptr), torch.from_numpy(indices), sampler=kind,
                                      seed=self._seed, item_weights=self._item_counts if self._weighted else None,
                                      adaptive_prob=self._adaptive_sampling_prob, adaptive_every=every)
        owned = None
        if self._world > 1:
            from rbpr.parallel import owned_triples
            owned = owned_triples(indptr, self._world, self._rank)
        self._datasets["train"] = EpochChunks(ds, bs, steps_per_chunk=int(self._config.get("fast_steps_per_chunk", 64)),
                                              generator=torch.Generator().manual_seed(self._seed + self._rank),
                                              device=dev, owned=owned, world=self._world)

    def interrupt(self) -> None:
        for e in self.trainer.engines.values():
            e.interrupt()

    def _seed_everything(self) -> None:
        random.seed(self._seed)
        np.random.seed(self._seed)
        torch.manual_seed(self._seed)

    def _get_trainer(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, datasets: dict[str, Any]) -> Trainer:
        trainer = Trainer(model, optimizer=optimizer, accelerator=self._accelerator,
                          custom_engines=self._config.get("custom_engines", {}))
        if self._adaptive and not getattr(self, "_fast", False):
            bs = getattr(datasets["train"], "total_batch_size", None) or datasets["train"].batch_size
            every = max(1, int(self._config["num_items"] * math.log(self._config["num_items"]) / bs))
            trainer.add_event("train", Events.GET_BATCH_COMPLETED(every=every), self._update_adaptive_stats)
        if not getattr(self, "_fast", False):
            trainer.add_event("train", Events.GET_BATCH_COMPLETED, self._train_batch)
        if self._skip_seen:
            trainer.add_event("eval", ModelEvents.FORWARD_COMPLETED, self._remove_seen_items)
        for name in trainer.engines:
            if name != "train" and getattr(self, "_world", 1) > 1:
                trainer.add_event(name, Events.STARTED, self._sync_user_shards)
            # kernels neutralise bad ids / an exhausted sampler and raise a device flag instead of
            # failing in place: poll it after the first batch, at every epoch end and before eval
            trainer.add_event(name, Events.ITERATION_COMPLETED, self._poll_after_first_batch)
            trainer.add_event(name, Events.EPOCH_COMPLETED, self._poll_device_errors)
        trainer.add_event("eval", Events.STARTED, self._poll_device_errors)
        early_stopping = None
        if self._early_stopping_metric is not None:
            early_stopping = attach_early_stopping(trainer, metric_name=self._early_stopping_metric,
                                                   patience=self._early_stopping_patience,
                                                   direction=self._early_stopping_direction)
        attach_preemptible(trainer, self._accelerator)
        if self._debug:
            trainer.add_event("train", Events.ITERATION_COMPLETED(every=2000), lambda e: e.terminate())
        attach_metrics(trainer, self._accelerator, self._metrics)
        if self._dir is not None:
            attach_checkpointer(trainer, self._accelerator, early_stopping=early_stopping,
                                checkpoint_objects=[*self._metrics.values(), _HotPathCounters(self)])
            attach_checkpoint_loader(trainer, self._accelerator, datasets)
        for key, handlers in self._events.items():
            for event, handler in handlers:
                trainer.add_event(key, event, handler, accelerator=self._accelerator)
        trainer.add_event("train", Events.EPOCH_STARTED, self._reset_metrics)
        trainer.add_event("train", Events.ITERATION_COMPLETED, self._update_metrics)
        return trainer

    # ---- hooks on the hot path -------------------------------------------------------------------
    def _to_device(self, batch: dict[str, torch.Tensor]) -> None:
        dev = self._accelerator.device
        for k, v in batch.items():
            if torch.is_tensor(v) and v.device != dev:
                batch[k] = v.to(dev, non_blocking=True)

    @torch.no_grad()
    def _train_batch(self, engine: Any) -> None:
        batch = engine.state.batch
        self._to_device(batch)
        if batch["item"].dim() < 2:
            batch["item"] = batch["item"].unsqueeze(-1)
        num = batch["item"].size(-1)
        if self._adaptive:
            eng = self._model.logits_model.engine()
            batch["neg"] = eng.sample_adaptive_padded(batch["user"], batch["seen_items"], num,
                                                      self._adaptive_sampling_prob, self._neg_seed, self._neg_calls,
                                                      opt_step=getattr(self._model, "_opt_step", None))
        else:
            kind = native.SAMPLER_WEIGHTED if self._weighted else native.SAMPLER_UNIFORM
            batch["neg"] = self._sampler_ctx.sample_padded(batch["seen_items"], self._config["num_items"], num,
                                                           self._neg_seed, self._neg_calls, kind)
        self._neg_calls += 1

    @torch.no_grad()
    def _update_adaptive_stats(self) -> None:
        flush = getattr(self._model, "flush", None)
        if flush is not None:
            flush()
        self._model.logits_model.engine().adaptive_update_stats()

    @torch.no_grad()
    def _remove_seen_items(self, engine: Any) -> None:
        batch, output = engine.state.batch, engine.state.output
        if getattr(output, "fused", False) and dict.__contains__(batch, "seen_csr"):
            output.mask_seen(batch["seen_csr"])  # all-items batch: masking happens inside the scoring call
        elif (seen := batch.get("seen_items")) is not None:
            self._sampler_ctx.mask_seen_padded(output["logits"], seen)
        self._to_device(batch)  # metrics read batch["target"] on the device next

    def _sync_user_shards(self) -> None:
        self._model.sync_user_shards()

    def _poll_after_first_batch(self, engine: Any) -> None:
        polled = self.__dict__.setdefault("_polled_engines", set())
        if engine.state.name not in polled:
            polled.add(engine.state.name)
            self._poll_device_errors()

    def _poll_device_errors(self, *_: Any) -> None:
        """Raise (RuntimeError) what the kernels flagged since the last poll: out-of-range user / item
        / triple ids, a user without any negative left, a metric target outside {0,1}."""
        lm = getattr(self._model, "logits_model", None)
        eng = getattr(lm, "_engine", None)
        if eng is not None:
            eng.sync_check()
        self._sampler_ctx.sync_check()
        from revisit_bpr.metrics import metric as _metric_mod
        for ctx in _metric_mod._contexts.values():
            ctx.sync_check()

    def _reset_metrics(self, engine: Any) -> None:
        if engine.state.was_interrupted:
            return
        for m in ("bpr_loss", "l2_reg", "logits_diff"):
            engine.state.metrics[f"_{m}"] = torch.tensor(0.0, device=self._accelerator.device)

    @torch.no_grad()
    def _update_metrics(self, engine: Any) -> None:
        state = engine.state
        out = state.output
        for m in ("bpr_loss", "l2_reg"):
            state.metrics[f"_{m}"] += out[m]
            state.metrics[m] = state.metrics[f"_{m}"] / state.epoch_iteration
        state.metrics["_logits_diff"] += out["logits"].abs().mean()
        state.metrics["logits_diff"] = state.metrics["_logits_diff"] / state.epoch_iteration
