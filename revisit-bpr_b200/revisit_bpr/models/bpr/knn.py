"""`ItemKNN` and `FreeItemKNN` logits models with the reference's constructor signatures, parameter
names (`_weights`, `_bias`), initialisation and `get_features()` keys (reference:
revisit_bpr/models/bpr/model.py:156-251), evaluated by librbpr.so:

    ItemKNN      logits[b,i] = w[item[b,i]] . SUM_{s kept} w[seen_items[b,s]]   (+ bias[item[b,i]])
    FreeItemKNN  logits[b,i] = SUM_{s kept} W[item[b,i], seen_items[b,s]]       (+ bias[item[b,i]])

with a seen entry dropped when its id occurs among item[b,:] ("discard current items").  Forward
and backward are CUDA kernels (rbpr_knn_forward/backward, rbpr_freeknn_forward/backward) behind
autograd Functions, so these models train through `loss.backward()` + any torch optimizer exactly
like the reference's; `Model` does not fuse their step.  No config of the reference uses them.

There is no CPU path: parameters and inputs must live on a CUDA device.
"""
from __future__ import annotations

from typing import Any

import torch

from rbpr import native
from rbpr.engine import Context
from revisit_bpr.models.bpr.model import BaseLogitModel

_contexts: dict[torch.device, Context] = {}


def _context(t: torch.Tensor) -> Context:
    if not t.is_cuda:
        raise native.NativeError("ItemKNN / FreeItemKNN run on a B200 only (librbpr.so): move the model and "
                                 "the batch to a CUDA device; there is no CPU fallback")
    if t.device not in _contexts:
        _contexts[t.device] = Context(t.device)
    return _contexts[t.device]


class _ItemKNNLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, weights: torch.Tensor, bias: torch.Tensor | None, item: torch.Tensor,
                seen: torch.Tensor) -> torch.Tensor:
        lib = _context(weights)
        logits, profile, keep = lib.knn_forward(weights.detach(), None if bias is None else bias.detach(), item, seen)
        ctx.save_for_backward(weights, item, seen, keep, profile)
        ctx.with_bias = bias is not None
        return logits

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):  # noqa: ANN205
        weights, item, seen, keep, profile = ctx.saved_tensors
        grad_w = torch.zeros_like(weights)
        grad_b = torch.zeros(weights.size(0), dtype=torch.float32, device=weights.device) if ctx.with_bias else None
        _context(weights).knn_backward(weights.detach(), item, seen, keep, profile,
                                       grad.to(torch.float32).contiguous(), grad_w, grad_b)
        return grad_w, grad_b, None, None


class _FreeItemKNNLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, weights: torch.Tensor, bias: torch.Tensor | None, item: torch.Tensor,
                seen: torch.Tensor) -> torch.Tensor:
        logits, keep = _context(weights).freeknn_forward(weights.detach(), None if bias is None else bias.detach(),
                                                         item, seen)
        ctx.save_for_backward(weights, item, seen, keep)
        ctx.with_bias = bias is not None
        return logits

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):  # noqa: ANN205
        weights, item, seen, keep = ctx.saved_tensors
        grad_w = torch.zeros_like(weights)
        grad_b = torch.zeros(weights.size(0), dtype=torch.float32, device=weights.device) if ctx.with_bias else None
        _context(weights).freeknn_backward(weights.size(0), item, seen, keep, grad.to(torch.float32).contiguous(),
                                           grad_w, grad_b)
        return grad_w, grad_b, None, None


class _Neighbourhood(BaseLogitModel):
    """Shared shell: an (num_items, width) weight matrix, an optional per-item bias, U(0,1) init with
    the padding row zeroed (reference model.py:169-174, 214-222)."""
    _zero_padding_bias = False

    def __init__(self, num_items: int, width: int, padding_idx: int, bias: bool) -> None:
        super().__init__()
        self._padding_idx = padding_idx
        self._weights = torch.nn.Parameter(torch.empty(num_items, width))
        self._bias = torch.nn.Parameter(torch.empty(num_items)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self._weights.uniform_()
            self._weights[self._padding_idx].zero_()
            if self._bias is not None:
                self._bias.zero_()

    def get_features(self) -> dict[str, torch.Tensor]:
        return {"item": self._weights, "bias": self._bias}

    def _ids(self, item: torch.Tensor, seen: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if item.dim() != 2 or seen.dim() != 2:
            raise IndexError("item must be (batch size, num items) and seen_items (batch size, seen items)")
        dev = self._weights.device
        return item.to(dev), seen.to(dev)


class ItemKNN(_Neighbourhood):
    def __init__(self, num_items: int, hidden_dim: int, padding_idx: int = 0, bias: bool = False) -> None:
        super().__init__(num_items, hidden_dim, padding_idx, bias)

    def forward(self, _: torch.Tensor, item: torch.Tensor, other: dict[str, torch.Tensor]) -> torch.Tensor:
        item, seen = self._ids(item, other["seen_items"])
        return _ItemKNNLogits.apply(self._weights, self._bias, item, seen)


class FreeItemKNN(_Neighbourhood):
    def __init__(self, num_items: int, padding_idx: int = 0, bias: bool = False) -> None:
        super().__init__(num_items, num_items, padding_idx, bias)

    def forward(self, _: torch.Tensor, item: torch.Tensor, other: dict[str, torch.Tensor]) -> torch.Tensor:
        if "seen_items" not in (other or {}):
            raise ValueError("seen_items should be present")
        item, seen = self._ids(item, other["seen_items"])
        return _FreeItemKNNLogits.apply(self._weights, self._bias, item, seen)
