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
        if svals.size and (svals.min() < 0 or svals.max() >= num_items):
            raise IndexError("item id outside num_items in the seen-items file")
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


def _lists_to_csr(rows: list[Any]) -> tuple[torch.Tensor, torch.Tensor]:
    """(indptr (n+1,) int64, indices int32 ascending per row, duplicates dropped) of a list of id lists."""
    arrs = [np.unique(np.asarray(r, dtype=np.int64).reshape(-1)) for r in rows]
    indptr = np.zeros(len(arrs) + 1, dtype=np.int64)
    np.cumsum([a.size for a in arrs], out=indptr[1:])
    flat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.int64)
    return torch.from_numpy(indptr), torch.from_numpy(flat.astype(np.int32))


class AllItemsBatch(dict):
    """The batch `AllItemsCollator` returns.  Same keys and tensors as the reference's
    (experiments/bpr/dataset.py:274-296) — `user` (B,), `item` (B,I) = arange, `target` (B,I)
    multi-hot, `seen_items` (B,S) 0-padded — but the three wide ones are only built when read: the
    fused eval path (Model.forward -> AllItemsEval -> rbpr_score_metrics) works from the compact forms
    `target_csr` / `seen_csr` and the `all_items` marker, so a (B,I) int64 + a (B,I) float matrix per
    batch (30 MB at ML-20M, batch 128) never cross PCIe."""

    _LAZY = ("item", "target", "seen_items")

    def __init__(self, users: torch.Tensor, num_items: int, positives: list[Any], seen: list[Any],
                 padding_value: float) -> None:
        super().__init__(user=users, all_items=True, target_csr=_lists_to_csr(positives), seen_csr=_lists_to_csr(seen))
        self._num_items, self._positives, self._seen_lists, self._pad = num_items, positives, seen, padding_value

    def __missing__(self, key: str) -> torch.Tensor:
        n, dev = self["user"].numel(), self["user"].device
        if key == "item":
            val = torch.arange(self._num_items, dtype=torch.long, device=dev).unsqueeze(0).repeat(n, 1)
        elif key == "target":
            indptr, indices = self["target_csr"]
            val = torch.zeros(n, self._num_items, device=dev)
            rows = torch.repeat_interleave(torch.arange(n), indptr[1:] - indptr[:-1])
            val[rows.to(dev), indices.long().to(dev)] = 1.0
        elif key == "seen_items":
            val = pad_sequence([torch.as_tensor(s) for s in self._seen_lists], batch_first=True,
                               padding_value=self._pad).to(dev)
        else:
            raise KeyError(key)
        self[key] = val
        return val

    def __contains__(self, key: object) -> bool:
        return key in self._LAZY or dict.__contains__(self, key)

    def get(self, key: str, default: Any = None) -> Any:
        try:
            return self[key]
        except KeyError:
            return default


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
        return AllItemsBatch(torch.as_tensor(cols["user"]), self._num_items, cols["item"], cols["seen_items"],
                             self._padding_value)


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
                 generator: torch.Generator | None = None, device: torch.device | str | None = None,
                 owned: tuple[int, int] | None = None, world: int = 1) -> None:
        self.dataset, self.batch_size, self.steps_per_chunk = dataset, int(batch_size), int(steps_per_chunk)
        self.generator, self.device = generator, device
        self.total_batch_size = self.batch_size
        # data parallel: this rank permutes only the triples [lo, hi) of its own users, and every rank
        # runs ceil(nnz / (B * world)) steps per epoch (each step ends in a collective)
        self.owned, self.world = owned, int(world)

    @property
    def steps_per_epoch(self) -> int:
        per = self.batch_size * self.world
        return (len(self.dataset) + per - 1) // per

    def __len__(self) -> int:
        return (self.steps_per_epoch + self.steps_per_chunk - 1) // self.steps_per_chunk

    def __iter__(self) -> Iterator[dict[str, Any]]:
        if self.owned is None:
            perm = torch.randperm(len(self.dataset), generator=self.generator)
        else:
            lo, hi = self.owned
            need = self.steps_per_epoch * self.batch_size
            perm = torch.randperm(hi - lo, generator=self.generator) + lo
            if 0 < perm.numel() < need:
                perm = perm.repeat((need + perm.numel() - 1) // perm.numel())
            perm = perm[:need]
        if self.device is not None:
            perm = perm.to(self.device)
        chunk = self.batch_size * self.steps_per_chunk
        for a in range(0, perm.numel(), chunk):
            yield {"triple_idx": perm[a:a + chunk], "batch_size": self.batch_size}
