"""Host mirror of the reference's input reading for the ALS path, over libmyrrix_ingest.so
(include/myrrix_ingest.h).

    read_input_files(input_dir) -> Interactions
mirrors InputFilesReader.readInputFiles
(online-local/src/net/myrrix/online/generation/InputFilesReader.java:64-196): every
`*.csv`, `*.csv.gz`, `*.csv.zip` file of the directory in last-modified order, lines of
`user,item[,strength]`; the result is what `als_set_interactions` takes (CSR by user with
dense indices) plus the long IDs behind the indices, knownItemIDs and the two tag sets.
There is no Python fallback: the library does the parsing and the folding.
"""
import ctypes as C
import gzip
import os
import re
import zipfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmyrrix_ingest.so")

INGEST_OK, INGEST_E_ARG, INGEST_E_BAD_LINES, INGEST_E_STATE, INGEST_E_OOM, INGEST_E_RANGE = range(6)
(N_USERS, N_ITEMS, NNZ, KNOWN_NNZ, LINES, BAD_LINES, N_ITEM_TAGS, N_USER_TAGS) = range(8)

_H = C.c_void_p
_i64p, _i32p, _f32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)
# Every symbol include/myrrix_ingest.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("ingest_create", C.c_int, [C.c_float, C.POINTER(_H)]),
    ("ingest_destroy", None, [_H]),
    ("ingest_set_parallelism", C.c_int, [_H, C.c_int, C.c_size_t]),
    ("ingest_add_file", C.c_int, [_H, C.c_char_p, C.c_size_t]),
    ("ingest_finish", C.c_int, [_H]),
    ("ingest_count", C.c_int64, [_H, C.c_int]),
    ("ingest_get_ids", C.c_int, [_H, C.c_int, _i64p]),
    ("ingest_get_csr", C.c_int, [_H, _i64p, _i32p, _f32p]),
    ("ingest_get_known", C.c_int, [_H, _i64p, _i32p]),
    ("ingest_get_tags", C.c_int, [_H, C.c_int, _i64p]),
    ("ingest_tag_id", C.c_int64, [C.c_char_p, C.c_size_t]),
    ("ingest_last_error", C.c_char_p, [_H]),
]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmyrrix_ingest.so is not built (python myrrix-recommender_b200/build.py)")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


class TooManyBadLines(IOError):
    """'Too many bad lines; aborting' (InputFilesReader.java:95-97)."""


@dataclass
class Interactions:
    user_ids: np.ndarray      # int64 [n_users]: dense user index -> long ID (keys of RbyRow)
    item_ids: np.ndarray      # int64 [n_items]: dense item index -> long ID (keys of RbyColumn)
    row_ptr: np.ndarray       # int64 [n_users + 1]
    col_idx: np.ndarray       # int32 [nnz], ascending inside a row
    val: np.ndarray           # float32 [nnz]
    known_ptr: np.ndarray     # knownItemIDs as a CSR pattern (entries before pruning)
    known_idx: np.ndarray
    item_tag_ids: np.ndarray  # itemTagIDs (users given as tags)
    user_tag_ids: np.ndarray  # userTagIDs (items given as tags)
    lines: int
    bad_lines: int


def tag_id(tag):
    """OneWayMigrator.toLongID: first 8 bytes of MD5(UTF-8), big-endian, as a signed long."""
    b = tag.encode("utf-8")
    return int(load().ingest_tag_id(b, len(b)))


class Ingest:
    """One accumulation of input files (the C handle)."""

    def __init__(self, zero_threshold=1e-4, max_threads=0, min_chunk_bytes=1 << 20):
        self.lib = load()
        self.h = _H()
        rc = self.lib.ingest_create(zero_threshold, C.byref(self.h))
        if rc != INGEST_OK:
            raise RuntimeError("ingest_create failed (%d)" % rc)
        self.lib.ingest_set_parallelism(self.h, max_threads, min_chunk_bytes)

    def add_bytes(self, data):
        rc = self.lib.ingest_add_file(self.h, data, len(data))
        if rc == INGEST_E_BAD_LINES:
            raise TooManyBadLines(self.lib.ingest_last_error(self.h).decode())
        if rc != INGEST_OK:
            raise RuntimeError("ingest_add_file: %s (%d)" % (self.lib.ingest_last_error(self.h).decode(), rc))

    def finish(self):
        L, h = self.lib, self.h
        rc = L.ingest_finish(h)
        if rc != INGEST_OK:
            raise RuntimeError("ingest_finish: %s (%d)" % (L.ingest_last_error(h).decode(), rc))
        n = {k: int(L.ingest_count(h, k)) for k in range(8)}
        uid, iid = np.empty(n[N_USERS], np.int64), np.empty(n[N_ITEMS], np.int64)
        ptr, kptr = np.empty(n[N_USERS] + 1, np.int64), np.empty(n[N_USERS] + 1, np.int64)
        idx, val = np.empty(n[NNZ], np.int32), np.empty(n[NNZ], np.float32)
        kidx = np.empty(n[KNOWN_NNZ], np.int32)
        itag, utag = np.empty(n[N_ITEM_TAGS], np.int64), np.empty(n[N_USER_TAGS], np.int64)
        p64 = lambda a: a.ctypes.data_as(_i64p)
        for rc in (L.ingest_get_ids(h, 0, p64(uid)), L.ingest_get_ids(h, 1, p64(iid)),
                   L.ingest_get_csr(h, p64(ptr), idx.ctypes.data_as(_i32p), val.ctypes.data_as(_f32p)),
                   L.ingest_get_known(h, p64(kptr), kidx.ctypes.data_as(_i32p)),
                   L.ingest_get_tags(h, 0, p64(itag)), L.ingest_get_tags(h, 1, p64(utag))):
            if rc != INGEST_OK:
                raise RuntimeError("ingest getter failed (%d)" % rc)
        return Interactions(uid, iid, ptr, idx, val, kptr, kidx, itag, utag, n[LINES], n[BAD_LINES])

    def close(self):
        if self.h:
            self.lib.ingest_destroy(self.h)
            self.h = _H()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


_CSV = re.compile(r".+\.csv(\.(zip|gz))?$")  # PatternFilenameFilter (InputFilesReader.java:71)


def _file_bytes(path):
    """FileLineIterator.getFileInputStream (common/.../iterator/FileLineIterator.java:93-105)."""
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return f.read()
    if path.endswith(".zip"):
        with zipfile.ZipFile(path) as z:
            return z.read(z.namelist()[0])  # the reference reads the first entry
    with open(path, "rb") as f:
        return f.read()


def read_input_files(input_dir, zero_threshold=1e-4):
    names = [n for n in os.listdir(input_dir) if _CSV.match(n)]
    paths = sorted((os.path.join(input_dir, n) for n in names),
                   key=lambda p: (os.path.getmtime(p), p))  # ByLastModifiedComparator (:86)
    with Ingest(zero_threshold) as ing:
        for p in paths:
            ing.add_bytes(_file_bytes(p))
        return ing.finish()


def read_csv_bytes(*chunks, zero_threshold=1e-4, max_threads=0, min_chunk_bytes=1 << 20):
    """The same, from in-memory file contents (one bytes object per file, in order)."""
    with Ingest(zero_threshold, max_threads, min_chunk_bytes) as ing:
        for c in chunks:
            ing.add_bytes(c)
        return ing.finish()
