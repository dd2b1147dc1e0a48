"""Data-parallel split of the BPR step (SURVEY.md §8 e): one process per GPU, users sharded by
owner, item table replicated, ONE all-reduce(sum) of the dense item-gradient buffer per step.

Replaces the reference's DDP launch (experiments/launcher.py:35-73) and the dense all-parameter
gradient all-reduce inside `accelerator.backward` (experiments/trainer.py:76): user rows, their
optimizer state and the user-gradient accumulator never leave their owner.
"""
from __future__ import annotations

from typing import Protocol

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(indptr: np.ndarray, world: int) -> np.ndarray:
    """User-row cut points (world+1,) of contiguous owner blocks balanced by interaction count."""
    nnz = int(indptr[-1])
    cuts = np.searchsorted(indptr, np.linspace(0, nnz, world + 1), side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, indptr.size - 1
    return np.maximum.accumulate(cuts)


def owned_triples(indptr: np.ndarray, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) range of triple ids (COO positions) whose user is owned by `rank`."""
    cuts = shard_bounds(indptr, world)
    return int(indptr[cuts[rank]]), int(indptr[cuts[rank + 1]])


class StepEngine(Protocol):
    def grad_step(self, triple_idx: torch.Tensor, seed: int, step: int, neg_in=None, want_neg: bool = False): ...
    def item_grad_tensor(self) -> torch.Tensor: ...
    def apply_item_grads(self, step: int) -> None: ...


class DataParallelTrainer:
    """Drives `rbpr.engine.Engine` (or anything with its grad_step / item_grad_tensor /
    apply_item_grads) across the default process group."""

    def __init__(self, engine: StepEngine, group: dist.ProcessGroup | None = None) -> None:
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._grad = engine.item_grad_tensor() if self.world > 1 else None

    def step(self, local_triples: torch.Tensor, seed: int, step: int, neg_in=None, want_neg: bool = False):
        """One global step: this rank contributes `local_triples` (ids of triples of ITS users)."""
        stats, negs = self.engine.grad_step(local_triples, seed, step, neg_in=neg_in, want_neg=want_neg)
        if self.world > 1:
            dist.all_reduce(self._grad, op=dist.ReduceOp.SUM, group=self.group)
        self.engine.apply_item_grads(step)
        return stats, negs

    def reduce_stats(self, stats: torch.Tensor) -> torch.Tensor:
        """Sum per-step statistics over ranks (the loss is a sum over triples)."""
        if self.world > 1:
            stats = stats.clone()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return stats
