"""`Loss` of the reference (revisit_bpr/models/bpr/loss.py:5-21): -log sigmoid(logits), averaged
when `size_average`.  Inside `Model` the loss is evaluated by the fused training kernel; this
module exists for the callers that use it on its own, and is plain elementwise torch."""
import torch


class Loss(torch.nn.Module):
    def __init__(self, size_average: bool = True) -> None:
        super().__init__()
        self.size_average = size_average

    def forward(self, logits: torch.Tensor) -> torch.Tensor:
        per_pair = torch.nn.functional.softplus(-logits)  # == -logsigmoid(logits)
        return per_pair.mean() if self.size_average else per_pair
