"""Input canonicalisation: libmyrrix_ingest.so against the line-by-line
restatement of InputFilesReader.readInputFiles (oracle/ingest_oracle.py). CPU only.

The C++ side sorts and folds; the oracle mutates maps of maps per line like the reference.
Bit-exact: IDs, pattern, fp32 sums (same additions in the same order), counters."""
import gzip
import hashlib
import os

import numpy as np
import pytest

import myrrix_recommender_b200 as M
from myrrix_recommender_b200 import ingest as ING
from oracle import ingest_oracle as O


def _compare(res, ora):
    rby_row, rby_col, known, item_tags, user_tags, lines, bad = ora
    assert res.lines == lines and res.bad_lines == bad
    assert set(res.user_ids.tolist()) == set(rby_row) and len(res.user_ids) == len(rby_row)
    assert set(res.item_ids.tolist()) == set(rby_col) and len(res.item_ids) == len(rby_col)
    assert res.item_tag_ids.tolist() == item_tags and res.user_tag_ids.tolist() == user_tags
    n_cells = 0
    for u, uid in enumerate(res.user_ids.tolist()):
        a, b = res.row_ptr[u], res.row_ptr[u + 1]
        cols = res.col_idx[a:b]
        assert np.all(np.diff(cols) > 0)
        got = {int(res.item_ids[c]): res.val[a + k] for k, c in enumerate(cols)}
        exp = rby_row[uid]
        assert set(got) == set(exp), uid
        for i in exp:   # bit-exact fp32 sums
            assert np.float32(got[i]).view(np.uint32) == np.float32(exp[i]).view(np.uint32), (uid, i)
        n_cells += len(exp)
        ka, kb = res.known_ptr[u], res.known_ptr[u + 1]
        assert {int(res.item_ids[c]) for c in res.known_idx[ka:kb]} == set(known.get(uid, {}))
    assert n_cells == res.col_idx.size
    # the by-column map is the transpose of the by-row map
    t = {}
    for uid, row in rby_row.items():
        for i, v in row.items():
            t.setdefault(i, {})[uid] = v
    assert {i: r for i, r in rby_col.items() if r} == t
    assert set(known) <= set(rby_row)


def test_symbols_and_tag_hash():
    lib = ING.load()
    for name, _, _ in ING.SYMBOLS:   # every symbol include/myrrix_ingest.h declares
        assert hasattr(lib, name)
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "myrrix_ingest.h")).read()
    for name, _, _ in ING.SYMBOLS:
        assert name + "(" in hdr
    for tag in ["", "a", "foo", "tag with spaces", "x" * 55, "y" * 56, "z" * 64, "w" * 119, "é€"]:
        v = int.from_bytes(hashlib.md5(tag.encode()).digest()[:8], "big")
        v = v - (1 << 64) if v >= (1 << 63) else v
        assert ING.tag_id(tag) == v == O.to_long_id(tag)


def test_reference_line_semantics():
    """Every branch of InputFilesReader.java:93-165 on a hand-written file."""
    csv = b"\n".join([
        b"user,item,value",              # line 1 unparseable: header, not a bad line (:135-141)
        b"# comment", b"",               # skipped (:101-103)
        b"1,10,2.5", b"1,10,0.25",       # duplicates sum in fp32 (FastByIDFloatMap.increment)
        b"1,11", b" 2 , 10 , 3 ",        # no value = 1.0 (:127-129); fields are trimmed
        b"2,12,1e-5",                    # pruned as near-zero, row and column keys stay
        b"3,13,4", b"3,13,",             # empty value deletes (:125, :160-165): user 3 disappears
        b"4,14,1", b"4,14,", b"4,14,7",  # delete then re-add starts from 7
        b"5,15,1,extra,fields",          # further fields ignored
        b'"red",16,1', b'6,"blue",2',    # tags (:107-121, :150-158)
        b'"a","b",1',                    # two tags: bad line (:144-148)
        b"7", b"x,1,1", b"8,9,abc", b"8,9,NaN", b"8,9,Infinity", b"9,9,1e39",   # bad lines
        b"10,17,0x1.8p1", b"10,18,2.5f", b"10,19,-.5", b"10,20,+3.", b"-11,21,1E2",
        b"9223372036854775807,22,1", b"9223372036854775808,22,1",   # Long range
    ]) + b"\n"
    res = ING.read_csv_bytes(csv)
    _compare(res, O.read_input([csv]))
    ids = res.user_ids.tolist()
    assert 3 not in ids and 2 in ids and 7 not in ids
    assert res.bad_lines == 8 and res.lines == 30
    row = lambda uid: {int(res.item_ids[c]): float(res.val[k]) for u in [ids.index(uid)]
                       for k, c in zip(range(res.row_ptr[u], res.row_ptr[u + 1]),
                                       res.col_idx[res.row_ptr[u]:res.row_ptr[u + 1]])}
    assert row(1) == {10: 2.75, 11: 1.0} and row(4) == {14: 7.0}
    assert row(2) == {10: 3.0}                       # 1e-5 pruned, |v| < 1e-4
    assert 12 in res.item_ids.tolist()               # ...but item 12 keeps its (empty) column
    assert row(10) == {17: 3.0, 18: 2.5, 19: -0.5, 20: 3.0} and row(-11) == {21: 100.0}
    assert res.item_tag_ids.tolist() == [ING.tag_id("red")] and res.user_tag_ids.tolist() == [ING.tag_id("blue")]
    assert (1 << 63) - 1 in ids


def test_too_many_bad_lines_and_line_endings():
    bad = b"\n".join([b"1,2,3"] + [b"oops"] * 101 + [b"4,5,6"])
    with pytest.raises(ING.TooManyBadLines):
        ING.read_csv_bytes(bad)
    with pytest.raises(O.TooManyBadLines):
        O.read_input([bad])
    ok = b"\n".join([b"1,2,3"] + [b"oops"] * 101)      # the check runs when the NEXT line arrives
    assert ING.read_csv_bytes(ok).bad_lines == 101 == O.read_input([ok])[6]
    mixed = b"1,2,3\r\n1,3,4\r5,6\n\n#x\r\n5,6,\r\n"
    _compare(ING.read_csv_bytes(mixed), O.read_input([mixed]))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_streams_bit_exact(seed):
    """Long random event streams over a small ID space: many duplicates, deletions, re-adds,
    near-zero sums, several files; fp32 sums must match bit for bit."""
    rng = np.random.default_rng(seed)
    files = []
    for _ in range(3):
        out = []
        for _ in range(20000):
            u, i = int(rng.integers(-50, 300)), int(rng.integers(0, 120))
            r = rng.random()
            if r < 0.08:
                out.append("%d,%d," % (u, i))
            elif r < 0.16:
                out.append("%d,%d" % (u, i))
            elif r < 0.20:
                out.append("%d,%d,%g" % (u, i, rng.normal() * 1e-4))
            elif r < 0.23:
                out.append("#%d" % u)
            elif r < 0.231:
                out.append("%d;%d" % (u, i))
            else:
                out.append("%d, %d ,%s" % (u, i, repr(float(np.float32(rng.normal() * 3)))))
        files.append("\n".join(out).encode() + b"\n")
    res = ING.read_csv_bytes(*files)
    ora = O.read_input(files)
    _compare(res, ora)
    assert res.bad_lines > 0 and res.col_idx.size > 1000
    # chunked parsing on many threads gives the same bits, whatever the cut points
    for threads, chunk in ((8, 1), (3, 1000), (16, 77)):
        par = ING.read_csv_bytes(*files, max_threads=threads, min_chunk_bytes=chunk)
        for f in ("user_ids", "item_ids", "row_ptr", "col_idx", "known_ptr", "known_idx"):
            assert np.array_equal(getattr(par, f), getattr(res, f)), (threads, chunk, f)
        assert np.array_equal(par.val.view(np.uint32), res.val.view(np.uint32))
        assert (par.lines, par.bad_lines) == (res.lines, res.bad_lines)


def test_chunk_boundaries_line_endings_header_and_abort():
    """Cuts inside \r\n pairs, the header rule on the very first line only, and the bad-line
    abort replayed across chunks -- with one-byte minimum chunks on 8 threads."""
    kw = dict(max_threads=8, min_chunk_bytes=1)
    mixed = b"head,er\r\n1,2,3\r\n1,3,4\r5,6\n\n#x\r\n5,6,\r\n7,8,9\r\n\r\n9,9\r"
    _compare(ING.read_csv_bytes(mixed, **kw), O.read_input([mixed]))
    two = [b"1,1,1\nbad\n", b"nothdr\n2,2,2\n"]     # a second file's first line is not forgiven
    r = ING.read_csv_bytes(*two, **kw)
    _compare(r, O.read_input(two))
    assert r.bad_lines == 2
    for n_bad, tail, aborts in ((100, b"4,5,6\n", False), (101, b"", False), (101, b"4,5,6\n", True),
                                (101, b"#c\n", True), (150, b"", True)):
        body = b"1,2,3\n" + b"oops\n" * n_bad + tail
        if aborts:
            with pytest.raises(ING.TooManyBadLines):
                ING.read_csv_bytes(body, **kw)
            with pytest.raises(O.TooManyBadLines):
                O.read_input([body])
        else:
            _compare(ING.read_csv_bytes(body, **kw), O.read_input([body]))
    split = [b"1,2,3\n" + b"oops\n" * 101, b"4,5,6\n"]    # the next line arrives in the next file
    with pytest.raises(ING.TooManyBadLines):
        ING.read_csv_bytes(*split, **kw)
    with pytest.raises(O.TooManyBadLines):
        O.read_input(split)


def test_parse_float_matches_correct_rounding():
    """Float.parseFloat rounds the decimal string to float32 once (no double rounding)."""
    vals = ["0.1", "16777217", "1.00000017881393432617187500001", "3.4028235e38", "1e-45", "7.038531e-26",
            "1.17549435E-38", "0.000099999", "123456.7890123", "4.35", "8.41e21", "2.5e-5d", "0x1.fffffep127"]
    rng = np.random.default_rng(9)
    for _ in range(3000):   # short decimals (the single-division fast path) and long ones
        m = str(int(rng.integers(0, 1 << int(rng.integers(1, 40)))))
        f = int(rng.integers(0, 13))
        txt = (m[:-f] or "0") + "." + m[-f:].rjust(f, "0") if f else m
        vals.append(("-" if rng.random() < 0.3 else "") + txt)
    csv = "\n".join("1,%d,%s" % (k, v) for k, v in enumerate(vals)).encode()
    res = ING.read_csv_bytes(csv, zero_threshold=0.0)
    assert res.col_idx.size == len(vals)
    for k, v in enumerate(vals):
        assert np.float32(res.val[k]).view(np.uint32) == O.parse_float(v).view(np.uint32), v


def test_directory_reader_order_and_feeds_als_shape(tmp_path):
    """read_input_files: *.csv / *.csv.gz in last-modified order (a later file's deletion wins),
    other files skipped; the CSR has the shape als_set_interactions takes."""
    a, b, c = tmp_path / "b_first.csv", tmp_path / "a_second.csv.gz", tmp_path / "notes.txt"
    a.write_bytes(b"1,1,5\n2,1,1\n")
    with gzip.open(b, "wb") as f:
        f.write(b"1,1,\n1,2,2\n")
    c.write_bytes(b"9,9,9\n")
    os.utime(a, (1000, 1000)); os.utime(b, (2000, 2000))
    res = ING.read_input_files(str(tmp_path))
    assert res.user_ids.tolist() == [1, 2] and res.item_ids.tolist() == [1, 2]
    assert res.row_ptr.tolist() == [0, 1, 2] and res.col_idx.tolist() == [1, 0] and res.val.tolist() == [2.0, 1.0]
    assert res.row_ptr.dtype == np.int64 and res.col_idx.dtype == np.int32 and res.val.dtype == np.float32
    os.utime(a, (3000, 3000))    # now the deletion comes first and the 5 survives
    res = ING.read_input_files(str(tmp_path))
    assert sorted(zip(res.col_idx.tolist(), res.val.tolist())) == [(0, 1.0), (0, 5.0), (1, 2.0)]
