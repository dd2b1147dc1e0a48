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


class OwnerBatchSampler:
    """Index batches of ONE rank for `DataLoader(batch_sampler=...)`: a fresh permutation of the
    triples owned by this rank every epoch, cut into `batch_size` batches.  Every rank yields the same
    number of batches — ceil(nnz / (batch_size * world)) — because each step ends in a collective:
    a shard a little shorter than that wraps around, a longer one drops its tail (owner blocks are
    balanced by interaction count, so the difference is less than one user's row)."""

    def __init__(self, indptr: np.ndarray, world: int, rank: int, batch_size: int,
                 generator: torch.Generator | None = None) -> None:
        self.lo, self.hi = owned_triples(indptr, world, rank)
        self.batch_size, self.generator = int(batch_size), generator
        nnz = int(indptr[-1])
        self.steps = (nnz + self.batch_size * world - 1) // (self.batch_size * world)

    def __len__(self) -> int:
        return self.steps

    def __iter__(self):
        n = self.hi - self.lo
        need = self.steps * self.batch_size
        perm = torch.randperm(n, generator=self.generator) + self.lo
        if 0 < n < need:
            perm = perm.repeat((need + n - 1) // n)
        perm = perm[:need].tolist()
        for a in range(0, need, self.batch_size):
            yield perm[a:a + self.batch_size]


class RoundRobinBatches:
    """Eval loader of one rank: every world-th batch of the underlying loader (what
    accelerate.prepare_data_loader does to the reference's loaders, experiments/bpr/exp.py:110),
    without padding the tail — per-rank (sum, count) pairs are reduced instead of means."""

    def __init__(self, loader, world: int, rank: int) -> None:
        self._loader, self._world, self._rank = loader, int(world), int(rank)

    def __iter__(self):
        for i, batch in enumerate(self._loader):
            if i % self._world == self._rank:
                yield batch

    def __len__(self) -> int:
        n = len(self._loader)
        return n // self._world + (1 if self._rank < n % self._world else 0)

    def __getattr__(self, name: str):
        return getattr(self._loader, name)


def reduce_sum_count(pairs: torch.Tensor, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """ONE all-reduce of stacked (sum, count) pairs (n, 2) -> per-entry global mean (n,)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        pairs = pairs.clone()
        dist.all_reduce(pairs, op=dist.ReduceOp.SUM, group=group)
    return pairs[:, 0] / pairs[:, 1]


def sync_row_shards(tensors: list[torch.Tensor], cuts: np.ndarray, group: dist.ProcessGroup | None = None) -> None:
    """Make every rank hold every owner's rows: rows [cuts[r], cuts[r+1]) of each tensor are broadcast
    from rank r (user rows and their optimizer state live only on their owner between eval passes)."""
    world = dist.get_world_size(group)
    for r in range(world):
        a, b = int(cuts[r]), int(cuts[r + 1])
        if b <= a:
            continue
        src = dist.get_global_rank(group, r) if group is not None else r
        for t in tensors:
            dist.broadcast(t[a:b], src=src, group=group)
