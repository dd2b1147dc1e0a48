"""JSONL ingest through the native parser of librbpr.so (csrc/ingest.cu): the reference's on-disk
files -> numpy arrays / CSR, one mmap'ed pass, no per-line json.loads."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from rbpr import native


def _take(ptr: C.c_void_p, n: int, lib: C.CDLL) -> np.ndarray:
    try:
        if n == 0:
            return np.zeros(0, dtype=np.int64)
        buf = (C.c_int64 * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.int64).copy()
    finally:
        lib.rbpr_ingest_free(ptr)


def read_pairs(path: Path | str, key_a: str = "user", key_b: str = "item") -> tuple[np.ndarray, np.ndarray]:
    """Lines {"user": u, "item": i} -> (users, items) int64 arrays in file order."""
    lib = native.load()
    a, b, n = C.c_void_p(), C.c_void_p(), C.c_int64()
    rc = lib.rbpr_ingest_pairs(str(path).encode(), key_a.encode(), key_b.encode(), C.byref(a), C.byref(b), C.byref(n))
    if rc != 0:
        raise ValueError(lib.rbpr_ingest_last_error().decode())
    return _take(a, n.value, lib), _take(b, n.value, lib)


def read_lists(path: Path | str, key_a: str = "user", key_list: str = "seen_items"):
    """Lines {"user": u, "<key_list>": [...]} -> (users (rows,), offsets (rows+1,), values) int64."""
    lib = native.load()
    a, off, vals, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
    rc = lib.rbpr_ingest_lists(str(path).encode(), key_a.encode(), key_list.encode(), C.byref(a), C.byref(off),
                               C.byref(vals), C.byref(n))
    if rc != 0:
        raise ValueError(lib.rbpr_ingest_last_error().decode())
    offsets = _take(off, n.value + 1, lib)
    return _take(a, n.value, lib), offsets, _take(vals, int(offsets[-1]), lib)


def pairs_to_csr(users: np.ndarray, items: np.ndarray, num_users: int, num_items: int):
    """De-duplicated binary interaction matrix as CSR (indptr int64, indices int32 ascending per
    row) — what scipy's dok -> csr conversion yields in the reference (dataset.py:183-190)."""
    if users.size and (users.min() < 0 or users.max() >= num_users or items.min() < 0 or items.max() >= num_items):
        raise IndexError("user / item id outside the (num_users, num_items) matrix")
    key = np.unique(users * np.int64(num_items) + items)
    coo_u, coo_i = key // num_items, key % num_items
    indptr = np.zeros(num_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(coo_u, minlength=num_users), out=indptr[1:])
    return indptr, coo_i.astype(np.int32), coo_u
