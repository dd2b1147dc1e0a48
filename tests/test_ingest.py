"""Native JSONL ingest (csrc/ingest.cu) against Python's json module on the reference's on-disk
formats, including the shapes real files take: other keys, key order, whitespace, blank lines,
empty lists, negative and large ids, missing trailing newline; malformed lines raise with the
line number.  CPU only (host code of librbpr.so)."""
import json
import time

import numpy as np
import pytest


def test_pairs_match_json(tmp_path):
    from rbpr import ingest
    rng = np.random.default_rng(0)
    recs = [{"user": int(u), "item": int(i)} for u, i in zip(rng.integers(0, 10**6, 5000), rng.integers(0, 10**5, 5000))]
    lines = [json.dumps(r) for r in recs]
    lines[3] = '{"item": 7, "rating": 4.5, "user": 12, "tags": ["a,b", "}{"], "meta": {"user": 99}}'
    lines[4] = '  {  "user" :  -3 ,\t"item":9007199254740993 }  '
    lines.insert(10, "")
    p = tmp_path / "train.jsonl"
    p.write_text("\n".join(lines))  # no trailing newline
    u, i = ingest.read_pairs(p)
    exp = [json.loads(x) for x in lines if x.strip()]
    assert u.tolist() == [r["user"] for r in exp] and i.tolist() == [r["item"] for r in exp]
    assert u[3] == 12 and i[3] == 7 and u[4] == -3 and i[4] == 9007199254740993


def test_lists_match_json_and_grouped_files(tmp_path):
    from rbpr import ingest
    rows = [{"user": 5, "seen_items": [3, 1, 2]}, {"user": 1, "seen_items": []}, {"seen_items": [10], "user": 7, "x": "y"}]
    p = tmp_path / "seen.jsonl"
    p.write_text("\n".join(json.dumps(r) for r in rows) + "\n")
    users, off, vals = ingest.read_lists(p, "user", "seen_items")
    assert users.tolist() == [5, 1, 7] and off.tolist() == [0, 3, 3, 4] and vals.tolist() == [3, 1, 2, 10]
    g = tmp_path / "test-grouped.jsonl"
    g.write_text('{"user": 2, "item": [4, 6]}\n{"user": 9, "item": [1]}\n')
    users, off, vals = ingest.read_lists(g, "user", "item")
    assert users.tolist() == [2, 9] and off.tolist() == [0, 2, 3] and vals.tolist() == [4, 6, 1]
    empty = tmp_path / "empty.jsonl"
    empty.write_text("")
    assert ingest.read_pairs(empty)[0].size == 0


@pytest.mark.parametrize("bad", ['{"user": 1}', '{"user": 1.5, "item": 2}', '{"user": 1, "item": 2', 'user,item',
                                 '{"user": "1", "item": 2}'])
def test_malformed_lines_raise_with_line_number(tmp_path, bad):
    from rbpr import ingest
    p = tmp_path / "bad.jsonl"
    p.write_text('{"user": 1, "item": 2}\n' + bad + "\n")
    with pytest.raises(ValueError, match="line 2"):
        ingest.read_pairs(p)
    with pytest.raises(ValueError, match="cannot open"):
        ingest.read_pairs(tmp_path / "missing.jsonl")


def test_pairs_to_csr_and_throughput(tmp_path):
    from rbpr import ingest
    rng = np.random.default_rng(1)
    n = 300_000
    u, i = rng.integers(1, 5000, n), rng.integers(1, 2000, n)
    p = tmp_path / "big.jsonl"
    with open(p, "w") as f:
        f.write("".join(f'{{"user": {a}, "item": {b}}}\n' for a, b in zip(u.tolist(), i.tolist())))
    t0 = time.perf_counter()
    uu, ii = ingest.read_pairs(p)
    native_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    with open(p) as f:
        ref = [json.loads(line) for line in f]
    json_s = time.perf_counter() - t0
    assert uu.tolist() == [r["user"] for r in ref] and ii.tolist() == [r["item"] for r in ref]
    assert native_s < json_s  # typically 20-50x
    indptr, indices, coo_u = ingest.pairs_to_csr(uu, ii, 5000, 2000)
    dense = np.zeros((5000, 2000), dtype=bool)
    dense[u, i] = True
    assert indptr[-1] == dense.sum() == indices.size
    for r in (1, 17, 4999):
        assert indices[indptr[r]:indptr[r + 1]].tolist() == np.nonzero(dense[r])[0].tolist()
    assert (coo_u == np.repeat(np.arange(5000), np.diff(indptr))).all()
    with pytest.raises(IndexError):
        ingest.pairs_to_csr(uu, ii, 100, 2000)
    print(f"ingest: native {n / native_s / 1e6:.1f} M lines/s vs json {n / json_s / 1e6:.2f} M lines/s")
