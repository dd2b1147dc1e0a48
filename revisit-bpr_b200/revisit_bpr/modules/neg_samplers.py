"""Negative samplers with the reference's interface (revisit_bpr/modules/neg_samplers.py:9-132):
`Sampler.sample(batch) -> LongTensor (batch, num)` where `batch` carries `item` (B,num), `user`
(B,) and the 0-padded `seen_items` (B,S).

The draws are made on the device by a counter-based generator (Philox4x32-10):
negative = f(seed, call number, row, seen row).  `seed` is `neg_gen.initial_seed()`, so the
reference's `torch.Generator(device).manual_seed(seed)` argument keeps its meaning; the stream
itself is ours (torch.multinomial's B*I exponential draws are what the reference spends 85 % of
its step on).  The target distributions are the reference's: uniform (or popularity-weighted)
over the unseen non-padding items; factor/rank-adaptive for `AdaptiveSampler`.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from rbpr import native
from rbpr.engine import Context


class Sampler(ABC):
    @abstractmethod
    def sample(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        pass


def _require_cuda(gen: torch.Generator) -> torch.device:
    dev = torch.device(gen.device)
    if dev.type != "cuda":
        raise native.NativeError("the CUDA samplers need neg_gen = torch.Generator(device='cuda'): "
                                 "there is no CPU fallback")
    return dev


class UniformSampler(Sampler):
    """Uniform negatives over {1..num_items-1} minus the row's seen items."""

    def __init__(self, num_items: int, neg_gen: torch.Generator) -> None:
        self._num_items = int(num_items)
        self._neg_gen = neg_gen
        self._calls = 0
        self._ctx: Context | None = None

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(_require_cuda(self._neg_gen))
        return self._ctx

    def sample(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        num = batch["item"].size(-1) if batch["item"].dim() > 1 else 1
        ctx = self._context()
        out = ctx.sample_padded(batch["seen_items"], self._num_items, num, self._neg_gen.initial_seed(),
                                self._calls, native.SAMPLER_UNIFORM)
        self._calls += 1
        ctx.sync_check()  # a row without any unseen item raises here, like torch.multinomial would
        return out


class AdaptiveSampler(Sampler):
    """Adaptive (factor / rank) negative sampling, `AdaptiveSampler(model, num_items, sampling_prob,
    neg_gen, every)`: factor ~ |u_f| * std_f, rank ~ Geometric(sampling_prob) among the user's
    unseen items ordered by that factor of a (stale) item-table snapshot refreshed every `every`
    calls (reference neg_samplers.py:40-132)."""

    def __init__(self, model: torch.nn.Module, num_items: int, sampling_prob: float,
                 neg_gen: torch.Generator, every: int) -> None:
        self._model = model
        self._num_items = int(num_items)
        self._sampling_prob = float(sampling_prob)
        self._neg_gen = neg_gen
        self._every = int(every)
        self._iteration_cnt = 0
        self._stats = None

    def _engine(self):
        _require_cuda(self._neg_gen)
        return self._model.logits_model.engine()

    def sample(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        if self._stats is None:
            raise AttributeError("AdaptiveSampler.update_stats() must be called before sample()")
        self._iteration_cnt += 1
        num = batch["item"].size(-1) if batch["item"].dim() > 1 else 1
        eng = self._engine()
        out = eng.sample_adaptive_padded(batch["user"], batch["seen_items"], num, self._sampling_prob,
                                         self._neg_gen.initial_seed(), self._iteration_cnt - 1, self._stats)
        eng.sync_check()
        if self._iteration_cnt % self._every == 0:
            self.update_stats()
        return out

    @torch.no_grad()
    def update_stats(self) -> None:
        flush = getattr(self._model, "flush", None)
        if flush is not None:
            flush()
        self._stats = self._engine().adaptive_update_stats()
