"""Streaming per-user ranking metrics with the reference's interface
(revisit_bpr/metrics/metric.py:9-61): `m(output, target)` accumulates, `m.compute(output, target)`
returns the per-user vector, `m.get_metric(reset)` the running mean, `state_dict` /
`load_state_dict` / `set_accelerator` as in the reference.  `output` (B,I) are scores (seen items
already pushed to -1e13 by the caller), `target` (B,I) is multi-hot.

One CUDA kernel (rbpr_topk_metrics_dense) selects the top-k of every row once and derives NDCG,
Recall and Precision from the hit flags; the reference sorts the full (B,I) matrix per metric."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

import torch

from rbpr import native
from rbpr.engine import Context

_contexts: dict[torch.device, Context] = {}


def _context(device: torch.device) -> Context:
    device = torch.device(device)
    if device.type != "cuda":
        raise native.NativeError("revisit_bpr.metrics run on a B200 only (librbpr.so): pass CUDA tensors; "
                                 "there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def validate_metric_inputs(output: torch.Tensor, target: torch.Tensor) -> None:
    # the binary-target check of the reference (metric.py:100-107) happens inside the kernel and
    # surfaces as ValueError from `topk_metrics`
    if output.size() != target.size():
        raise IndexError("Different sizes in output and target tensors: "
                         f"output - {output.size()}, target - {target.size()}.")


def prepare_target(output: torch.Tensor, target: torch.Tensor, topk: int | None = None) -> torch.Tensor:
    """`target` re-ordered by descending `output` (reference metric.py:110-113), cut at `topk`.

    The metric classes never call this — ranking and metric arithmetic are one kernel
    (`topk_metrics`) — it exists for callers of the reference's helper.  The ranks come from the same
    top-k kernel, so at most RBPR_MAX_TOPK (128) columns can be returned: pass `topk` (the reference's
    callers all slice `[:, :topk]` right away) when the catalog is wider."""
    validate_metric_inputs(output, target)
    width = output.size(-1)
    k = width if topk is None else min(int(topk), width)
    if k > native.MAX_TOPK:
        raise NotImplementedError(f"prepare_target returns at most {native.MAX_TOPK} ranked columns: pass topk")
    ctx = _context(output.device)
    ranked = ctx.topk_metrics_dense(output, torch.zeros_like(output, dtype=torch.float32), [k], want_items=True)
    return torch.gather(target, -1, ranked["items"][:, :k].long())


def topk_metrics(output: torch.Tensor, target: torch.Tensor, topk: int, linear_gain: bool = False,
                 validate: bool = False, map_normalized: bool = True) -> dict[str, torch.Tensor]:
    """{'ndcg','recall','precision','map'} -> (B,) at cut-off min(topk, I)."""
    if output.dim() != 2:
        raise IndexError(f"metrics expect (users, items) tensors, got {tuple(output.shape)}")
    validate_metric_inputs(output, target)
    ctx = _context(output.device)
    res = ctx.topk_metrics_dense(output, target, [topk], linear_gain=linear_gain, map_normalized=map_normalized)
    if validate:
        try:
            ctx.sync_check()
        except native.NativeError as e:
            if "outside of 0 and 1" in str(e):
                raise ValueError(f"Target contains values outside of 0 and 1.\nTarget:\n{target}") from e
            raise
    return {k: v[:, 0] for k, v in res.items()}


class Metric(ABC):
    """Base class for all metrics."""

    @property
    def accelerator(self) -> Any | None:
        return getattr(self, "_accelerator", None)

    def set_accelerator(self, value: Any) -> None:
        self._accelerator = value

    @abstractmethod
    def state_dict(self) -> dict[str, Any]:
        pass

    @abstractmethod
    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        pass

    @abstractmethod
    def __call__(self, output: torch.Tensor, target: torch.Tensor) -> None:
        pass

    @abstractmethod
    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        pass

    @abstractmethod
    def get_metric(self, reset: bool = False) -> torch.Tensor:
        pass

    @abstractmethod
    def reset(self) -> None:
        pass


class MaskedMetric(Metric):
    """Interface for metrics with masked computation."""

    @abstractmethod
    def __call__(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> None:
        pass

    @abstractmethod
    def compute(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        pass


class _TopkMean(Metric):
    """Shared streaming state: sum of per-user values and user count (users without positives
    contribute 0 and still count, like the reference's nan_to_num)."""

    _key = "value"

    def __init__(self, topk: int) -> None:
        assert topk > 0, f"Invalid topk value: {topk}"
        self._topk = topk
        self._total = self._total_count = 0

    def state_dict(self) -> dict[str, Any]:
        return {f"total_{self._key}": self._total, "total_count": self._total_count}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._total, self._total_count = state_dict[f"total_{self._key}"], state_dict["total_count"]
        if self.accelerator is None:
            return
        self._total = self._total.to(self.accelerator.device)
        self._total_count = self._total_count.to(self.accelerator.device)

    def __call__(self, output: torch.Tensor, target: torch.Tensor) -> None:
        self._total_count += torch.tensor(target.size(0), device=output.device)
        self._total += self.compute(output, target).sum()

    # ---- fused evaluation (experiments.options.attach_metrics): one ranking pass feeds every metric --
    _fused_family: str | None = None  # which output of rbpr_score_metrics this metric reads

    def fused_request(self) -> tuple[str, ...] | None:
        """Names of the rbpr_score_metrics outputs this metric is computed from at cut-off `_topk`
        (None: not computable from the shared ranking pass)."""
        if self._fused_family is None or self._topk > native.MAX_TOPK:
            return None
        return (self._fused_family,)

    def fused_value(self, res: dict[str, torch.Tensor], col: int) -> torch.Tensor:
        return res[self._fused_family][:, col]

    def accumulate(self, values: torch.Tensor) -> None:
        """Add per-user values (B,) computed elsewhere (same bookkeeping as __call__)."""
        self._total_count += torch.tensor(float(values.size(0)), device=values.device)
        self._total += values.sum()

    def get_metric(self, reset: bool = False) -> torch.Tensor:
        metric = self._total / self._total_count
        if reset:
            self.reset()
        return metric

    def reset(self) -> None:
        device = torch.device("cpu") if self.accelerator is None else self.accelerator.device
        self._total = torch.tensor(0.0, device=device)
        self._total_count = torch.tensor(0.0, device=device)
