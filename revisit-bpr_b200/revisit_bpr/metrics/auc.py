"""`RocAucManySlow` (reference revisit_bpr/metrics/auc.py:122-179): per user, the fraction of
(positive, negative) score pairs ordered correctly; negatives are the items with target 0 whose
mask is non-zero (no mask = every item, seen ones included — they sit at -1e13).  A user without
positives yields NaN, as in the reference.  One CUDA kernel per batch (rbpr_auc_dense) instead of a
Python loop over users."""
from __future__ import annotations

from typing import Any

import torch

from revisit_bpr.metrics.metric import MaskedMetric, _context


class _AucBase(MaskedMetric):
    def __init__(self) -> None:
        self._total_auc = self._total_count = 0

    def state_dict(self) -> dict[str, Any]:
        return {"total_auc": self._total_auc, "total_count": self._total_count}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._total_auc, self._total_count = state_dict["total_auc"], state_dict["total_count"]
        if self.accelerator is None:
            return
        self._total_auc = self._total_auc.to(self.accelerator.device)
        self._total_count = self._total_count.to(self.accelerator.device)

    def __call__(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> None:
        self._total_count += torch.tensor(target.size(0), device=output.device)
        self._total_auc += self.compute(output, target, mask).sum()

    def compute(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        raise NotImplementedError

    def get_metric(self, reset: bool = False) -> torch.Tensor:
        metric = self._total_auc / self._total_count
        if reset:
            self.reset()
        return metric

    def reset(self) -> None:
        device = torch.device("cpu") if self.accelerator is None else self.accelerator.device
        self._total_auc = torch.tensor(0.0, device=device)
        self._total_count = torch.tensor(0.0, device=device)


class RocAucManySlow(_AucBase):
    def compute(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        if output.size() != target.size():
            raise IndexError(f"Different sizes in output and target tensors: output - {output.size()}, "
                             f"target - {target.size()}.")
        return _context(output.device).auc_dense(output, target, mask)


class RocAucMany(RocAucManySlow):
    """The reference's vectorised variant (auc.py:62-119) computes the same quantity."""


class RocAucOne(_AucBase):
    """One positive per row in column 0, negatives in columns 1.. (reference auc.py:10-59; the RQ1
    protocol with OnePosCollator): fraction of unmasked negatives scored below the positive."""

    def compute(self, output: torch.Tensor, _: torch.Tensor | None = None, mask: torch.Tensor | None = None) -> torch.Tensor:
        target = torch.zeros_like(output)
        target[:, 0] = 1.0
        return _context(output.device).auc_dense(output, target, mask)
