"""Stand-in for the handful of `accelerate.Accelerator` members the BPR trainer path touches
(device, prepare, backward, accumulate, reduce, wait_for_everyone, is_local_main_process); the real
accelerate package is preferred when importable."""
from __future__ import annotations

import contextlib
import os
from typing import Any

import torch
import torch.distributed as dist


class Accelerator:
    def __init__(self, device: torch.device | str | None = None, **_: Any) -> None:
        if device is None:
            local = int(os.environ.get("LOCAL_RANK", "0"))
            device = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)

    @property
    def is_local_main_process(self) -> bool:
        return int(os.environ.get("LOCAL_RANK", "0")) == 0

    @property
    def is_main_process(self) -> bool:
        return int(os.environ.get("RANK", "0")) == 0

    def prepare(self, *objs: Any) -> Any:
        out = [o.to(self.device) if isinstance(o, torch.nn.Module) else o for o in objs]
        return out[0] if len(out) == 1 else tuple(out)

    def prepare_data_loader(self, loader: Any, **_: Any) -> Any:
        return loader

    @contextlib.contextmanager
    def accumulate(self, *_: Any):
        yield

    def backward(self, loss: torch.Tensor, **kwargs: Any) -> None:
        if loss.requires_grad:
            loss.backward(**kwargs)

    def reduce(self, tensor: torch.Tensor, reduction: str = "sum") -> torch.Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            tensor = tensor.clone()
            dist.all_reduce(tensor)
            if reduction == "mean":
                tensor = tensor / dist.get_world_size()
        return tensor

    def wait_for_everyone(self) -> None:
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    def log(self, *_: Any, **__: Any) -> None:
        pass

    def end_training(self) -> None:
        pass
