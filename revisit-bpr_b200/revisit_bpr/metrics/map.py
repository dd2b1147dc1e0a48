"""MAP@k (reference revisit_bpr/metrics/map.py:8-77): mean over users of the average precision of
the top-k list, normalised by min(#positives, k) (`normalized=True`) or by the hits in the list."""
import torch

from revisit_bpr.metrics.metric import _TopkMean, topk_metrics


class MAP(_TopkMean):
    _key = "map"

    def __init__(self, topk: int, normalized: bool = True) -> None:
        super().__init__(topk)
        self._normalized = normalized
        self._fused_family = "map"  # one normalisation per ranking pass: attach_metrics groups by it

    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return topk_metrics(output, target, self._topk, validate=True, map_normalized=self._normalized)["map"]
