"""F-beta@k: (1 + b^2) * P * R / (b^2 * P + R + 1e-13) per user, from the Precision@k and Recall@k of
ONE top-k pass of the CUDA kernel (reference behaviour: revisit_bpr/metrics/fbeta.py:8-73, which
evaluates two separate metric objects and therefore sorts twice)."""
from __future__ import annotations

from typing import Any

import torch

from revisit_bpr.metrics.metric import _TopkMean, topk_metrics


class FBeta(_TopkMean):
    _key = "f"

    def __init__(self, topk: int, beta: float = 1.0) -> None:
        super().__init__(topk)
        self._beta = beta

    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        both = topk_metrics(output, target, self._topk, validate=True)
        b2 = self._beta * self._beta
        num = (1.0 + b2) * both["precision"] * both["recall"]
        return num / (b2 * both["precision"] + both["recall"] + 1e-13)

    def fused_request(self) -> tuple[str, ...] | None:
        from rbpr import native
        return None if self._topk > native.MAX_TOPK else ("precision", "recall")

    def fused_value(self, res: dict[str, torch.Tensor], col: int) -> torch.Tensor:
        b2 = self._beta * self._beta
        pr, rc = res["precision"][:, col], res["recall"][:, col]
        return (1.0 + b2) * pr * rc / (b2 * pr + rc + 1e-13)

    # The reference keeps (never updated) Precision / Recall sub-metrics in its state; their entries
    # are reproduced so that checkpoints written by either implementation load in the other.
    def state_dict(self) -> dict[str, Any]:
        state = super().state_dict()
        state["precision"] = {"total_precision": 0, "total_count": 0}
        state["recall"] = {"total_recall": 0, "total_count": 0}
        return state

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        super().load_state_dict({k: v for k, v in state_dict.items() if k in ("total_f", "total_count")})
