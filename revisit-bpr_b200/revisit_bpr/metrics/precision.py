"""Precision@k (reference revisit_bpr/metrics/precision.py:6-64): hits in the top-k over k."""
import torch

from revisit_bpr.metrics.metric import _TopkMean, topk_metrics


class Precision(_TopkMean):
    _key = "precision"
    _fused_family = "precision"

    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return topk_metrics(output, target, self._topk, validate=True)["precision"]
