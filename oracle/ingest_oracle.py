"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input reading, line by
line, with the reference's own data structures (maps of maps mutated per line):

  InputFilesReader.readInputFiles  online-local/src/net/myrrix/online/generation/InputFilesReader.java:64-196
  InputFilesReader.removeSmall     :198-211
  MatrixUtils.addTo / remove       common/src/net/myrrix/common/math/MatrixUtils.java:64-121
  FastByIDFloatMap.increment       common/src/net/myrrix/common/collection/FastByIDFloatMap.java:129-138
  LangUtils.parseFloat             common/src/net/myrrix/common/LangUtils.java:42-46
  OneWayMigrator.toLongID          Mahout 0.8 AbstractIDMigrator.hash (third party, pom.xml:85): first 8
                                   bytes of MD5(UTF-8 bytes), big-endian

Parity unpinned: the reference holds no test with inline data for this path (its tests read
`testdata/`, absent from the tree) and no JVM exists here; this file is the specification the
C++ implementation (csrc_host/ingest.cpp) is held to, statement by statement.
"""
import hashlib
import re
from fractions import Fraction

import numpy as np

_WS = " \t\n\x0b\x0c\r"  # Splitter.trimResults() (ASCII subset of CharMatcher.WHITESPACE)
_LONG = re.compile(r"[+-]?[0-9]+$")
_DEC = re.compile(r"[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?[fFdD]?$")
_HEX = re.compile(r"[+-]?0[xX]([0-9a-fA-F]+\.?[0-9a-fA-F]*|\.[0-9a-fA-F]+)[pP][+-]?[0-9]+[fFdD]?$")


class TooManyBadLines(IOError):
    pass


def to_long_id(tag):
    v = int.from_bytes(hashlib.md5(tag.encode("utf-8")).digest()[:8], "big")
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_long(s):
    """Long.parseLong."""
    if not _LONG.match(s):
        raise ValueError(s)
    v = int(s)
    if not -(1 << 63) <= v < (1 << 63):
        raise ValueError(s)
    return v


def _round_to_f32(exact):
    """Nearest float32 to an exact rational, ties to even (what Float.parseFloat returns)."""
    with np.errstate(over="ignore"):
        g = np.float32(float(exact))
    if not np.isfinite(g):
        return g
    with np.errstate(over="ignore"):
        cands = [np.nextafter(g, np.float32(-np.inf)), g, np.nextafter(g, np.float32(np.inf))]
    best = None
    for c in cands:
        if not np.isfinite(c):
            continue
        err = abs(Fraction(float(c)) - exact)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        key = (err, 0 if even else 1)
        if best is None or key < best[0]:
            best = (key, c)
    # overflow: halfway beyond the largest finite float rounds to infinity
    fmax = Fraction(float(np.finfo(np.float32).max))
    if abs(exact) >= fmax + Fraction(2) ** 103:
        return np.float32(np.inf if exact > 0 else -np.inf)
    return np.float32(best[1])


def parse_float(s):
    """LangUtils.parseFloat = Float.parseFloat + finiteness ("NaN"/"Infinity" are rejected)."""
    if _HEX.match(s):
        t = s[:-1] if s[-1] in "fFdD" else s
        sign = -1 if t[0] == "-" else 1
        mant, exp = re.split("[pP]", t.lstrip("+-")[2:])
        ip, _, fp = mant.partition(".")
        exact = Fraction(int((ip + fp) or "0", 16), 16 ** len(fp)) * Fraction(2) ** int(exp) * sign
    elif _DEC.match(s):
        t = s[:-1] if s[-1] in "fFdD" else s
        exact = Fraction(t)
    else:
        raise ValueError(s)
    v = _round_to_f32(exact)
    if not np.isfinite(v):
        raise ValueError(s)
    if v == 0 and s[0] == "-":
        v = np.float32(-0.0)  # Float.parseFloat("-0.0") keeps the sign
    return v


def _lines(data):
    """BufferedReader.readLine: \\n, \\r or \\r\\n end a line; no empty line after a final terminator."""
    out = re.split(r"\r\n|\n|\r", data.decode("utf-8"))
    if out and out[-1] == "":
        out.pop()
    return out


class _TooFewColumns(Exception):
    pass


def read_input(files, zero_threshold=1e-4):
    """files: list of bytes, one per input file, in last-modified order.
    Returns (rby_row, rby_column, known_item_ids, item_tag_ids, user_tag_ids, lines, bad_lines);
    the maps are dicts of dicts (the reference's hash maps: key order is immaterial)."""
    rby_row, rby_col, known = {}, {}, {}
    item_tags, user_tags = {}, {}
    lines = bad = 0
    zt = np.float32(zero_threshold)

    def add_to(m, a, b, v):  # MatrixUtils.addToByRow + FastByIDFloatMap.increment
        row = m.setdefault(a, {})
        row[b] = v if b not in row else np.float32(row[b] + v)

    def remove(m, a, b):     # MatrixUtils.removeByRow
        row = m.get(a)
        if row is not None:
            row.pop(b, None)
            if not row:
                del m[a]

    def field(tok):          # quoted = tag (hashed), else Long.parseLong
        if tok.startswith('"'):
            if len(tok) < 2:  # substring(1, 0): the reference dies; both ports count a bad line
                raise _TooFewColumns
            return to_long_id(tok[1:-1]), True
        return parse_long(tok), False

    for data in files:
        for line in _lines(data):
            if bad > 100:
                raise TooManyBadLines("Too many bad lines; aborting")
            lines += 1
            if line == "" or line[0] == "#":
                continue
            toks = [t.strip(_WS) for t in line.split(",")]
            try:
                uid, u_tag = field(toks[0])
                if len(toks) < 2:
                    raise _TooFewColumns       # NoSuchElementException
                iid, i_tag = field(toks[1])
                if len(toks) > 2:
                    value = np.float32("nan") if toks[2] == "" else parse_float(toks[2])
                else:
                    value = np.float32(1.0)
            except _TooFewColumns:
                bad += 1
                continue
            except ValueError:                 # IllegalArgumentException: header if first line
                if lines != 1:
                    bad += 1
                continue
            if u_tag and i_tag:
                bad += 1
                continue
            if u_tag:
                item_tags[uid] = True
            if i_tag:
                user_tags[iid] = True
            if np.isnan(value):
                remove(rby_row, uid, iid)
                remove(rby_col, iid, uid)
                s = known.get(uid)
                if s is not None:
                    s.pop(iid, None)
                    if not s:
                        del known[uid]
            else:
                add_to(rby_row, uid, iid, value)
                add_to(rby_col, iid, uid, value)
                known.setdefault(uid, {})[iid] = True
    for m in (rby_row, rby_col):  # removeSmall: inner entries only, empty rows keep their key
        for row in m.values():
            for k in [k for k, v in row.items() if abs(v) < zt]:
                del row[k]
    return rby_row, rby_col, known, list(item_tags), list(user_tags), lines, bad
