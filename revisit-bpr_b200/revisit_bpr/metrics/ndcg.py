"""NDCG@k (reference revisit_bpr/metrics/ndcg.py:27-91): gain 2^t-1 over log2(rank+2) ("exp") or
t/(rank+1) ("linear"); ideal DCG over min(k, #positives); users without positives score 0."""
import torch

from revisit_bpr.metrics.metric import _TopkMean, topk_metrics


class NDCG(_TopkMean):
    _key = "ndcg"

    def __init__(self, topk: int, gain_function: str = "exp") -> None:
        assert gain_function in ("exp", "linear"), f"Invalid gain_function value: {gain_function}"
        super().__init__(topk)
        self._linear = gain_function == "linear"
        self._fused_family = "ndcg_linear" if self._linear else "ndcg"

    def compute(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return topk_metrics(output, target, self._topk, linear_gain=self._linear)["ndcg"]
