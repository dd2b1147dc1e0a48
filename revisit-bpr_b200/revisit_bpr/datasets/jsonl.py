"""JsonLines datasets and the dict collator of the reference library
(revisit_bpr/datasets/jsonl.py:12-95), used by the RQ1 configs (`revisit_bpr.datasets.jsonl.Iter`,
`Collator(pad=["seen_items"])`): one JSON object per line, batches are dicts of tensors, keys in
`pad` are right-padded and get a `<key>_mask`.  Host code (the path into the kernels starts at the
batch dict these produce)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Iterator

import torch
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import Dataset, IterableDataset, get_worker_info


def _objects(path: Path, first: int = 0, stride: int = 1) -> Iterator[dict[str, Any]]:
    """Objects of lines first, first+stride, ... of a JsonLines file."""
    with path.open("r", encoding="utf-8") as fh:
        for number, line in enumerate(fh):
            if number >= first and (number - first) % stride == 0:
                yield json.loads(line)


class InMemory(Dataset):
    def __init__(self, path: Path | str) -> None:
        self._samples = list(_objects(Path(path)))

    def __len__(self) -> int:
        return len(self._samples)

    def __getitem__(self, idx: int) -> dict[str, Any]:
        return self._samples[idx]


class Iter(IterableDataset):
    """Streams the file; with W DataLoader workers, worker w reads lines w, w+W, w+2W, ..."""

    def __init__(self, path: Path | str) -> None:
        self._path = Path(path)

    def __iter__(self) -> Iterator[dict[str, Any]]:
        info = get_worker_info()
        if info is None or info.num_workers <= 0:
            return _objects(self._path)
        return _objects(self._path, info.id, info.num_workers)


class Collator:
    def __init__(self, pad: list[str] | None = None, padding_value: float = 0) -> None:
        self._pad = set(pad or [])
        self._padding_value = padding_value

    def _column(self, key: str, values: list[Any]) -> torch.Tensor:
        if key not in self._pad:
            return torch.tensor(values)
        rows = [torch.as_tensor(v) for v in values]
        return pad_sequence(rows, batch_first=True, padding_value=self._padding_value)

    def __call__(self, instances: list[dict[str, Any]]) -> dict[str, torch.Tensor]:
        columns: dict[str, list[Any]] = {}
        for inst in instances:
            for key, value in inst.items():
                columns.setdefault(key, []).append(value)
        batch = {key: self._column(key, values) for key, values in columns.items()}
        for key in self._pad:
            batch[f"{key}_mask"] = batch[key].ne(self._padding_value).float()
        return batch
