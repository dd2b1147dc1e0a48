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


def _random_json_value(rng, depth=0):
    kind = rng.randint(0, 7 if depth < 2 else 4)
    if kind == 0:
        return rng.randint(-10**6, 10**12)
    if kind == 1:
        alphabet = ["a", " ", '"', "\\", "{", "}", "[", "]", ":", ",", "é", "\n", "\t", "user", "item", "/", "0"]
        return "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 8)))
    if kind == 2:
        return rng.choice([None, True, False])
    if kind == 3:
        return rng.random() * rng.choice([1, 1e10, 1e-10, -1])
    if kind == 4:
        return rng.randint(0, 9)
    if kind == 5:
        return [_random_json_value(rng, depth + 1) for _ in range(rng.randint(0, 3))]
    return {str(_random_json_value(rng, 3)): _random_json_value(rng, depth + 1) for _ in range(rng.randint(0, 3))}


def _dump_with_random_layout(rng, obj):
    ws = lambda: "".join(rng.choice([" ", "", "", "\t"]) for _ in range(rng.randint(0, 2)))  # noqa: E731
    if isinstance(obj, dict):
        body = ("," + ws()).join(json.dumps(k, ensure_ascii=rng.random() < 0.5) + ws() + ":" + ws()
                                 + _dump_with_random_layout(rng, v) for k, v in obj.items())
        return "{" + ws() + body + ws() + "}"
    if isinstance(obj, list):
        return "[" + ws() + ("," + ws()).join(_dump_with_random_layout(rng, v) for v in obj) + ws() + "]"
    return json.dumps(obj, ensure_ascii=rng.random() < 0.5)


def test_scanner_agrees_with_json_on_randomised_valid_files(tmp_path):
    """Key order, whitespace, CRLF, blank lines, missing final newline, and arbitrary OTHER keys whose
    values nest objects / arrays / strings full of quotes, escapes and braces: same values as json.loads."""
    import random
    from rbpr import ingest
    rng = random.Random(20241)
    for trial in range(250):
        mode = rng.choice(["pairs", "lists"])
        lines, users, second = [], [], []
        for _ in range(rng.randint(0, 10)):
            extra = {str(_random_json_value(rng, 3)): _random_json_value(rng) for _ in range(rng.randint(0, 3))}
            fields = [(k, v) for k, v in extra.items() if k not in ("user", "item", "seen_items")]
            u = rng.randint(0, 10**9)
            val = rng.randint(0, 10**9) if mode == "pairs" else [rng.randint(0, 10**6) for _ in range(rng.randint(0, 5))]
            fields += [("user", u), ("item" if mode == "pairs" else "seen_items", val)]
            rng.shuffle(fields)
            users.append(u)
            second.append(val)
            lines.append(_dump_with_random_layout(rng, dict(fields)))
        sep = rng.choice(["\n", "\r\n"])
        text = sep.join(lines) + (sep if rng.random() < 0.7 else "")
        if lines and rng.random() < 0.2:
            text = text.replace(sep, sep + sep, 1)
        path = tmp_path / f"f{trial}.jsonl"
        path.write_bytes(text.encode("utf-8"))
        if mode == "pairs":
            a, b = ingest.read_pairs(path)
            assert a.tolist() == users and b.tolist() == second, text
        else:
            a, off, vals = ingest.read_lists(path)
            assert a.tolist() == users, text
            assert [vals[off[k]:off[k + 1]].tolist() for k in range(len(a))] == second, text


def test_scanner_survives_corrupted_files(tmp_path):
    """Random deletions / insertions / substitutions / truncations: either a ValueError or arrays that
    are consistent with each other — and whenever json.loads also accepts the file, the same values."""
    import random
    from rbpr import ingest
    rng = random.Random(977)
    base = {"pairs": '{"user": 12, "item": 7, "ts": 1.5e3, "tag": "a\\"b{}", "x": [1, {"y": null}]}\n{"item": 3, "user": 4}\n',
            "lists": '{"user": 1, "seen_items": [1, 2, 3], "z": {"a": [true, false]}}\n{"seen_items": [], "user": 9}\n'}
    chars = list('{}[]",:\\ \n\r\t0123456789-+.eEuseritm') + ["é", "\x00", "\x7f"]
    accepted = rejected = 0
    for trial in range(1500):
        mode = rng.choice(["pairs", "lists"])
        t = list(base[mode])
        for _ in range(rng.randint(1, 5)):
            op, pos = rng.randint(0, 3), rng.randint(0, max(0, len(t) - 1))
            if op == 0 and t:
                del t[pos]
            elif op == 1:
                t.insert(pos, rng.choice(chars))
            elif op == 2 and t:
                t[pos] = rng.choice(chars)
            else:
                t = t[:pos]
        text = "".join(t)
        path = tmp_path / "c.jsonl"
        path.write_bytes(text.encode("utf-8"))
        try:
            if mode == "pairs":
                a, b = ingest.read_pairs(path)
                assert len(a) == len(b)
                got = (a.tolist(), b.tolist())
            else:
                a, off, vals = ingest.read_lists(path)
                assert len(off) == len(a) + 1 and off[-1] == len(vals) and (np.diff(off) >= 0).all()
                got = (a.tolist(), [vals[off[k]:off[k + 1]].tolist() for k in range(len(a))])
        except ValueError:
            rejected += 1
            continue
        accepted += 1
        try:
            recs = [json.loads(line) for line in text.splitlines() if line.strip()]
            want = ([r["user"] for r in recs], [r["item" if mode == "pairs" else "seen_items"] for r in recs])
        except Exception:  # noqa: BLE001  (the scanner does not validate the values it skips)
            continue
        if all(isinstance(v, int) and not isinstance(v, bool) for v in want[0]):
            assert got == want, text
    assert accepted > 50 and rejected > 500
