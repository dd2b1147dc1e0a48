"""Recall@k (reference revisit_bpr/metrics/recall.py:6-64): hits in the top-k over all positives
of the user; validates sizes (IndexError) and a binary target (ValueError)."""
import torch

from revisit_bpr.metrics.metric import _TopkMean, topk_metrics


class Recall(_TopkMean):
    _key = "recall"
    _fused_family = "recall"

    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return topk_metrics(output, target, self._topk, validate=True)["recall"]
