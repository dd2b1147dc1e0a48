"""JsonLines datasets and the dict collator of the reference library
(revisit_bpr/datasets/jsonl.py:12-95), used by the RQ1 configs (`revisit_bpr.datasets.jsonl.Iter`,
`Collator(pad=["seen_items"])`): one JSON object per line, batches are dicts of tensors, keys in
`pad` are right-padded and get a `<key>_mask`.  Host code (the path into the kernels starts at the
batch dict these produce)."""
from __future__ import annotations

import json
from itertools import islice
from pathlib import Path
from typing import Any, Iterator

import torch
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import Dataset, IterableDataset, get_worker_info


class InMemory(Dataset):
    def __init__(self, path: Path | str) -> None:
        with Path(path).open("r", encoding="utf-8") as fh:
            self._samples = [json.loads(line) for line in fh]

    def __len__(self) -> int:
        return len(self._samples)

    def __getitem__(self, idx: int) -> dict[str, Any]:
        return self._samples[idx]


class Iter(IterableDataset):
    """Streams the file; with DataLoader workers, worker w reads lines w, w+W, w+2W, ..."""

    def __init__(self, path: Path | str) -> None:
        self._path = Path(path)

    def __iter__(self) -> Iterator[dict[str, Any]]:
        info = get_worker_info()
        start, step = (info.id, info.num_workers) if info is not None and info.num_workers > 0 else (0, 1)
        with self._path.open("r", encoding="utf-8") as fh:
            for line in islice(fh, start, None, step):
                yield json.loads(line)


class Collator:
    def __init__(self, pad: list[str] | None = None, padding_value: float = 0) -> None:
        self._pad = set(pad or [])
        self._padding_value = padding_value

    def __call__(self, instances: list[dict[str, Any]]) -> dict[str, torch.Tensor]:
        columns: dict[str, list[Any]] = {}
        for inst in instances:
            for key, value in inst.items():
                columns.setdefault(key, []).append(value)
        batch: dict[str, torch.Tensor] = {}
        for key, values in columns.items():
            if key in self._pad:
                batch[key] = pad_sequence([torch.as_tensor(v) for v in values], batch_first=True,
                                          padding_value=self._padding_value)
            else:
                batch[key] = torch.tensor(values)
        for key in self._pad:
            batch[f"{key}_mask"] = batch[key].ne(self._padding_value).float()
        return batch
