"""Stand-in for the `accelerate.Accelerator` members the BPR trainer path and its checkpoint glue
touch — device, prepare, backward, accumulate, reduce, wait_for_everyone, is_local_main_process,
and the checkpoint family (project_configuration / project_dir / save_iteration,
register_for_checkpointing, save_state, load_state, skip_first_batches) with accelerate's on-disk
layout (`<project_dir>/checkpoints/checkpoint_<n>/{pytorch_model.bin, optimizer.bin,
custom_checkpoint_<i>.pkl, random_states_<rank>.pkl}`, oldest folders removed beyond
`total_limit`).  The real accelerate package is preferred when importable."""
from __future__ import annotations

import contextlib
import dataclasses
import itertools
import os
import pickle
import random
import re
import shutil
from pathlib import Path
from typing import Any

import numpy as np
import torch
import torch.distributed as dist

CHECKPOINTS = "checkpoints"


@dataclasses.dataclass
class ProjectConfiguration:
    project_dir: str | None = None
    automatic_checkpoint_naming: bool = False
    total_limit: int | None = None
    iteration: int = 0


def _suffix(i: int) -> str:
    return "" if i == 0 else f"_{i}"


def numbered_checkpoints(root: Path) -> list[tuple[int, Path]]:
    """(number, folder) of every `..._<number>` folder under root, ascending."""
    found = []
    if root.is_dir():
        for d in root.iterdir():
            m = re.search(r"(\d+)$", d.name)
            if d.is_dir() and m:
                found.append((int(m.group(1)), d))
    return sorted(found, key=lambda x: x[0])


class _SkipFirst:
    """A loader whose first `skip` batches are dropped (accelerate.skip_first_batches)."""

    def __init__(self, loader: Any, skip: int) -> None:
        self._loader, self._skip = loader, int(skip)

    def __iter__(self):
        return itertools.islice(iter(self._loader), self._skip, None)

    def __len__(self) -> int:
        return max(0, len(self._loader) - self._skip)

    def __getattr__(self, name: str) -> Any:
        return getattr(self._loader, name)


class Accelerator:
    def __init__(self, device: torch.device | str | None = None, log_with: Any = None,
                 mixed_precision: str | None = None, project_dir: str | None = None,
                 project_config: ProjectConfiguration | None = None, **_: Any) -> None:
        if device is None:
            local = int(os.environ.get("LOCAL_RANK", "0"))
            device = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        if mixed_precision not in (None, "no"):
            raise NotImplementedError("fp32 only")
        self.device = torch.device(device)
        self.project_configuration = project_config or ProjectConfiguration(project_dir=project_dir)
        self.trackers: list[Any] = []
        self._models: list[torch.nn.Module] = []
        self._optimizers: list[torch.optim.Optimizer] = []
        self._custom: list[Any] = []
        del log_with

    # ---- processes ---------------------------------------------------------------------------------
    @property
    def is_local_main_process(self) -> bool:
        return int(os.environ.get("LOCAL_RANK", "0")) == 0

    @property
    def is_main_process(self) -> bool:
        return int(os.environ.get("RANK", "0")) == 0

    @property
    def process_index(self) -> int:
        return int(os.environ.get("RANK", "0"))

    def wait_for_everyone(self) -> None:
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    def reduce(self, tensor: torch.Tensor, reduction: str = "sum") -> torch.Tensor:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            tensor = tensor.clone()
            dist.all_reduce(tensor)
            if reduction == "mean":
                tensor = tensor / dist.get_world_size()
        return tensor

    # ---- training ----------------------------------------------------------------------------------
    def prepare(self, *objs: Any) -> Any:
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if not any(o is m for m in self._models):
                    self._models.append(o)
            elif isinstance(o, torch.optim.Optimizer) and not any(o is m for m in self._optimizers):
                self._optimizers.append(o)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def prepare_data_loader(self, loader: Any, **_: Any) -> Any:
        return loader

    @contextlib.contextmanager
    def accumulate(self, *_: Any):
        yield

    def backward(self, loss: torch.Tensor, **kwargs: Any) -> None:
        if loss.requires_grad:
            loss.backward(**kwargs)

    def skip_first_batches(self, dataloader: Any, num_batches: int = 0) -> Any:
        return dataloader if num_batches <= 0 else _SkipFirst(dataloader, num_batches)

    # ---- trackers (control plane of the reference: accepted, nothing recorded) ---------------------
    def init_trackers(self, *_: Any, **__: Any) -> None:
        pass

    def log(self, *_: Any, **__: Any) -> None:
        pass

    def end_training(self) -> None:
        pass

    def free_memory(self) -> None:
        self._models.clear()
        self._optimizers.clear()
        self._custom.clear()

    # ---- checkpoints -------------------------------------------------------------------------------
    @property
    def project_dir(self) -> str | None:
        return self.project_configuration.project_dir

    @property
    def save_iteration(self) -> int:
        return self.project_configuration.iteration

    def register_for_checkpointing(self, *objects: Any) -> None:
        bad = [o for o in objects if not (hasattr(o, "state_dict") and hasattr(o, "load_state_dict"))]
        if bad:
            raise ValueError(f"objects registered for checkpointing need state_dict / load_state_dict: {bad}")
        self._custom.extend(objects)

    def save_state(self, output_dir: str | None = None) -> str:
        cfg = self.project_configuration
        if cfg.automatic_checkpoint_naming:
            root = Path(cfg.project_dir) / CHECKPOINTS
            root.mkdir(parents=True, exist_ok=True)
            if cfg.total_limit is not None and self.is_main_process:
                kept = numbered_checkpoints(root)
                while len(kept) + 1 > cfg.total_limit and kept:
                    shutil.rmtree(kept.pop(0)[1], ignore_errors=True)
            out = root / f"checkpoint_{cfg.iteration}"
            if out.exists():
                raise ValueError(f"checkpoint directory {out} already exists")
        else:
            if output_dir is None and cfg.project_dir is None:
                raise ValueError("save_state needs an output_dir or a project_dir")
            out = Path(output_dir if output_dir is not None else cfg.project_dir)
        self.wait_for_everyone()
        out.mkdir(parents=True, exist_ok=True)
        if self.is_main_process:
            for i, m in enumerate(self._models):
                torch.save(m.state_dict(), out / f"pytorch_model{_suffix(i)}.bin")
            for i, o in enumerate(self._optimizers):
                torch.save(o.state_dict(), out / f"optimizer{_suffix(i)}.bin")
            for i, c in enumerate(self._custom):
                with open(out / f"custom_checkpoint_{i}.pkl", "wb") as fh:
                    torch.save(c.state_dict(), fh)
        rng = {"random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
               "torch_manual_seed": torch.get_rng_state()}
        if self.device.type == "cuda":
            rng["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
        with open(out / f"random_states_{self.process_index}.pkl", "wb") as fh:
            pickle.dump(rng, fh)
        if cfg.automatic_checkpoint_naming:
            cfg.iteration += 1
        return str(out)

    def load_state(self, input_dir: str | None = None) -> None:
        cfg = self.project_configuration
        if input_dir is None:
            if not cfg.automatic_checkpoint_naming:
                raise ValueError("load_state needs an input_dir without automatic checkpoint naming")
            found = numbered_checkpoints(Path(cfg.project_dir) / CHECKPOINTS)
            if not found:
                raise ValueError(f"no checkpoint under {Path(cfg.project_dir) / CHECKPOINTS}")
            src = found[-1][1]
        else:
            src = Path(input_dir)
            if not src.is_dir():
                raise ValueError(f"tried to find {src} but the folder does not exist")
        load = lambda p: torch.load(p, map_location=self.device, weights_only=False)  # noqa: E731
        for i, m in enumerate(self._models):
            m.load_state_dict(load(src / f"pytorch_model{_suffix(i)}.bin"))
        for i, o in enumerate(self._optimizers):
            o.load_state_dict(load(src / f"optimizer{_suffix(i)}.bin"))
        for i, c in enumerate(self._custom):
            c.load_state_dict(load(src / f"custom_checkpoint_{i}.pkl"))
        rng_file = src / f"random_states_{self.process_index}.pkl"
        if rng_file.exists():
            with contextlib.suppress(Exception), open(rng_file, "rb") as fh:  # best effort, like accelerate
                rng = pickle.load(fh)  # noqa: S301
                random.setstate(rng["random_state"])
                np.random.set_state(rng["numpy_random_seed"])
                torch.set_rng_state(rng["torch_manual_seed"])
                if "torch_cuda_manual_seed" in rng and self.device.type == "cuda":
                    torch.cuda.set_rng_state_all(rng["torch_cuda_manual_seed"])
