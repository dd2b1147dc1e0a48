"""Datasets / collators the BPR configs name (reference experiments/bpr/dataset.py), same class
names and constructor arguments, same batch dictionaries.

On-disk format (reference bin/datasets/format-repro.sh:56-81): `<split>.jsonl` lines
{"user": u, "item": i}; `<split>-user-seen-items.jsonl` lines {"user": u, "seen_items": [...]};
`<split>-grouped.jsonl` lines {"user": u, "item": [...]}.  Ids are 1-based, 0 is the padding id.

`SparseSamplingInMemoryWithCollator` additionally exposes the interaction matrix as CSR
(`csr()`), which is what the fast path binds (rbpr_bind_csr) — the reference's dense padded
`(num_users, max_seen)` matrix is still built because `collate_fn` must return `seen_items`.
"""
from __future__ import annotations

import json
from collections import defaultdict
from itertools import islice
from pathlib import Path
from typing import Any, Iterator

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import Dataset, IterableDataset, get_worker_info


def _read_seen(path: Path | str) -> dict[int, list[int]]:
    with Path(path).open("r", encoding="utf-8") as fh:
        return {rec["user"]: rec["seen_items"] for rec in map(json.loads, fh)}


class InMemory(Dataset):
    def __init__(self, path: Path | str, seen_items_path: Path | str) -> None:
        with Path(path).open("r", encoding="utf-8") as fh:
            self._samples = [json.loads(line) for line in fh]
        self._seen_items = _read_seen(seen_items_path)

    def __len__(self) -> int:
        return len(self._samples)

    def __getitem__(self, idx: int) -> dict[str, Any]:
        sample = self._samples[idx]
        return {**sample, "seen_items": self._seen_items[sample["user"]]}


class Iter(IterableDataset):
    def __init__(self, path: Path | str, seen_items_path: Path | str) -> None:
        self._path = Path(path)
        self._seen_items = _read_seen(seen_items_path)

    def __iter__(self) -> Iterator[dict[str, Any]]:
        info = get_worker_info()
        start, step = (info.id, info.num_workers) if info is not None and info.num_workers > 0 else (0, 1)
        with self._path.open("r", encoding="utf-8") as fh:
            for line in islice(fh, start, None, step):
                sample = json.loads(line)
                sample["seen_items"] = self._seen_items[sample["user"]]
                yield sample


class SparseSamplingInMemoryWithCollator(Dataset):
    """Train set as COO triples; `__getitem__` is the triple index, `collate_fn` fancy-indexes."""

    def __init__(self, path: Path | str, seen_items_path: Path | str, num_users: int, num_items: int,
                 padding_value: float = 0, put_on_cuda: bool = False) -> None:
        from rbpr import ingest  # native mmap parser (csrc/ingest.cu), no per-line json.loads
        users_a, items_a = ingest.read_pairs(path, "user", "item")
        # CSR of the de-duplicated matrix, rows ascending: the COO flattening the reference takes
        # from scipy's dok -> csr conversion
        self._indptr, self._indices, coo_u = ingest.pairs_to_csr(users_a, items_a, num_users, num_items)
        self._user_ids = torch.from_numpy(coo_u)
        self._item_ids = torch.from_numpy(self._indices.astype(np.int64))
        # dense 0-padded (num_users, max_seen) matrix of the reference; users without a line: [0]
        su, soff, svals = ingest.read_lists(seen_items_path, "user", "seen_items")
        if su.size and (su.min() < 0 or su.max() >= num_users):
            raise IndexError("user id outside num_users in the seen-items file")
        lens = np.diff(soff)
        width = max(1, int(lens.max()) if lens.size else 1)
        dense = np.full((num_users, width), padding_value, dtype=np.int64)
        if svals.size:
            rows = np.repeat(su, lens)
            cols = np.arange(svals.size) - np.repeat(soff[:-1], lens)
            dense[rows, cols] = svals
        self._seen_items = torch.from_numpy(dense)
        self.num_users, self.num_items = num_users, num_items
        if put_on_cuda and torch.cuda.is_available():
            self._user_ids = self._user_ids.cuda()
            self._item_ids = self._item_ids.cuda()
            self._seen_items = self._seen_items.cuda()

    def __len__(self) -> int:
        return len(self._user_ids)

    def __getitem__(self, idx: int) -> int:
        return idx

    def collate_fn(self, indices: list[int]) -> dict[str, torch.Tensor]:
        idx = torch.as_tensor(indices, device=self._user_ids.device)
        users, items = self._user_ids[idx], self._item_ids[idx]
        return {"user": users, "item": items, "seen_items": self._seen_items[users]}

    def csr(self) -> tuple[np.ndarray, np.ndarray]:
        """(indptr (num_users+1,) int64, indices (nnz,) int32 ascending per row)."""
        return self._indptr, self._indices


class AllItemsCollator:
    """Eval batches: every item scored for every user, multi-hot target, padded seen items."""

    def __init__(self, num_items: int, padding_value: float = 0) -> None:
        self._num_items = num_items
        self._padding_value = padding_value

    def __call__(self, instances: list[dict[str, Any]]) -> dict[str, torch.Tensor]:
        cols = defaultdict(list)
        for inst in instances:
            for k, v in inst.items():
                cols[k].append(v)
        n = len(instances)
        target = torch.zeros(n, self._num_items)
        for r, pos in enumerate(cols["item"]):
            target[r, torch.as_tensor(pos, dtype=torch.long)] = 1.0
        return {
            "user": torch.as_tensor(cols["user"]),
            "item": torch.arange(self._num_items, dtype=torch.long).unsqueeze(0).repeat(n, 1),
            "target": target,
            "seen_items": pad_sequence([torch.as_tensor(s) for s in cols["seen_items"]], batch_first=True,
                                       padding_value=self._padding_value),
        }


class OnePosCollator:
    """RQ1 / AUC protocol (reference dataset.py:193-225): each instance names its positive as an
    index into its own `seen_items`; the batch scores that positive (column 0) against every item
    the user has not seen (columns 1..), target = one-hot column 0.  One instance per batch."""

    def __init__(self, num_items: int) -> None:
        self._num_items = num_items

    def __call__(self, instances: list[dict[str, Any]]) -> dict[str, torch.Tensor]:
        cols = defaultdict(list)
        for inst in instances:
            for k, v in inst.items():
                cols[k].append(v)
        batch = {k: torch.tensor(v) for k, v in cols.items()}
        seen = batch["seen_items"].view(-1)
        positive = seen[batch["item"]]
        unseen = torch.ones(self._num_items, dtype=torch.bool)
        unseen[0] = False  # padding item
        unseen[seen] = False
        batch["item"] = torch.hstack((positive.unsqueeze(0), torch.arange(self._num_items)[unseen].unsqueeze(0)))
        target = torch.zeros_like(batch["item"], dtype=torch.float)
        target[:, 0] = 1.0
        batch["target"] = target
        return batch


class EpochChunks:
    """Loader for the fast path (our extension): instead of B-sized index batches it yields, per
    epoch, a shuffled permutation of all triple ids in chunks of `steps_per_chunk` steps —
    `{"triple_idx": ids, "batch_size": B}` — which `revisit_bpr.models.BPR.forward` turns into that
    many complete training steps in one library call.  The permutation is what
    `DataLoader(shuffle=True, generator=g)` would draw (reference exp.py:111-115)."""

    def __init__(self, dataset: SparseSamplingInMemoryWithCollator, batch_size: int, steps_per_chunk: int = 64,
                 generator: torch.Generator | None = None, device: torch.device | str | None = None) -> None:
        self.dataset, self.batch_size, self.steps_per_chunk = dataset, int(batch_size), int(steps_per_chunk)
        self.generator, self.device = generator, device
        self.total_batch_size = self.batch_size

    @property
    def steps_per_epoch(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __len__(self) -> int:
        return (self.steps_per_epoch + self.steps_per_chunk - 1) // self.steps_per_chunk

    def __iter__(self) -> Iterator[dict[str, Any]]:
        perm = torch.randperm(len(self.dataset), generator=self.generator)
        if self.device is not None:
            perm = perm.to(self.device)
        chunk = self.batch_size * self.steps_per_chunk
        for a in range(0, perm.numel(), chunk):
            yield {"triple_idx": perm[a:a + chunk], "batch_size": self.batch_size}
