"""Host mirror of GenerationSerializer.writeGeneration / readGeneration
(online-local/src/net/myrrix/online/generation/GenerationSerializer.java:84-94) over
libmyrrix_model_io.so (include/myrrix_model_io.h): the library produces / parses the Java
object stream, this module adds the gzip layer of IOUtils.writeObjectToFile (file must end in
.gz, common/src/net/myrrix/common/io/IOUtils.java:275-283). No Python fallback for the codec."""
import ctypes as C
import gzip
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmyrrix_model_io.so")
(MODEL_IO_OK, MODEL_IO_E_ARG, MODEL_IO_E_FORMAT, MODEL_IO_E_NONFINITE, MODEL_IO_E_RAGGED,
 MODEL_IO_E_OOM) = range(6)

_i64p, _f32p, _u8p = C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint8)


class ModelIoDesc(C.Structure):
    _fields_ = [("features", C.c_int32),
                ("n_users", C.c_int64), ("user_ids", _i64p), ("x", _f32p),
                ("n_items", C.c_int64), ("item_ids", _i64p), ("y", _f32p),
                ("has_known", C.c_int32),
                ("n_known_users", C.c_int64), ("known_user_ids", _i64p),
                ("known_ptr", _i64p), ("known_item_ids", _i64p),
                ("n_item_tags", C.c_int64), ("item_tags", _i64p),
                ("n_user_tags", C.c_int64), ("user_tags", _i64p)]


_R = C.c_void_p
# Every symbol include/myrrix_model_io.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("model_io_write", C.c_int, [C.POINTER(ModelIoDesc), C.POINTER(_u8p), C.POINTER(C.c_size_t)]),
    ("model_io_free", None, [C.c_void_p]),
    ("model_io_read", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_R)]),
    ("model_io_reader_destroy", None, [_R]),
    ("model_io_count", C.c_int64, [_R, C.c_int]),
    ("model_io_get_matrix", C.c_int, [_R, C.c_int, _i64p, _f32p]),
    ("model_io_get_known", C.c_int, [_R, _i64p, _i64p, _i64p]),
    ("model_io_get_tags", C.c_int, [_R, C.c_int, _i64p]),
]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmyrrix_model_io.so is not built (python myrrix-recommender_b200/build.py)")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


@dataclass
class Generation:
    """The serialised part of net.myrrix.online.generation.Generation."""
    user_ids: np.ndarray                 # int64 [n_users]
    X: np.ndarray                        # float32 [n_users, features]
    item_ids: np.ndarray
    Y: np.ndarray
    known_user_ids: np.ndarray = None    # knownItemIDs as CSR over long IDs; None = not recorded
    known_ptr: np.ndarray = None
    known_item_ids: np.ndarray = None
    item_tag_ids: np.ndarray = field(default_factory=lambda: np.empty(0, np.int64))
    user_tag_ids: np.ndarray = field(default_factory=lambda: np.empty(0, np.int64))


def _i64(a):
    return np.ascontiguousarray(a, np.int64)


def to_bytes(g):
    """The object stream (before gzip)."""
    lib = load()
    X, Y = np.ascontiguousarray(g.X, np.float32), np.ascontiguousarray(g.Y, np.float32)
    k = X.shape[1] if X.ndim == 2 and X.size else (Y.shape[1] if Y.ndim == 2 else 0)
    uid, iid = _i64(g.user_ids), _i64(g.item_ids)
    assert len(uid) == len(X) and len(iid) == len(Y) and (Y.size == 0 or Y.shape[1] == k)
    it, ut = _i64(g.item_tag_ids), _i64(g.user_tag_ids)
    d = ModelIoDesc()
    d.features = k
    d.n_users, d.user_ids, d.x = len(uid), uid.ctypes.data_as(_i64p), X.ctypes.data_as(_f32p)
    d.n_items, d.item_ids, d.y = len(iid), iid.ctypes.data_as(_i64p), Y.ctypes.data_as(_f32p)
    keep = [uid, iid, X, Y, it, ut]
    if g.known_user_ids is not None:
        ku, kp, ki = _i64(g.known_user_ids), _i64(g.known_ptr), _i64(g.known_item_ids)
        assert len(kp) == len(ku) + 1 and kp[-1] == len(ki)
        d.has_known, d.n_known_users = 1, len(ku)
        d.known_user_ids, d.known_ptr, d.known_item_ids = (a.ctypes.data_as(_i64p) for a in (ku, kp, ki))
        keep += [ku, kp, ki]
    d.n_item_tags, d.item_tags = len(it), it.ctypes.data_as(_i64p)
    d.n_user_tags, d.user_tags = len(ut), ut.ctypes.data_as(_i64p)
    out, n = _u8p(), C.c_size_t(0)
    rc = lib.model_io_write(C.byref(d), C.byref(out), C.byref(n))
    if rc == MODEL_IO_E_NONFINITE:
        raise ValueError("non-finite feature value")     # Preconditions.checkState(isFinite(f))
    if rc != MODEL_IO_OK:
        raise RuntimeError("model_io_write failed (%d)" % rc)
    try:
        return C.string_at(out, n.value)
    finally:
        lib.model_io_free(out)


def from_bytes(data):
    lib = load()
    r = _R()
    rc = lib.model_io_read(data, len(data), C.byref(r))
    if rc != MODEL_IO_OK:
        raise IOError("Can't read model (%s)" % {MODEL_IO_E_FORMAT: "not a GenerationSerializer stream",
                                                 MODEL_IO_E_NONFINITE: "non-finite feature value",
                                                 MODEL_IO_E_RAGGED: "rows of different lengths"}.get(rc, rc))
    try:
        n = [int(lib.model_io_count(r, w)) for w in range(10)]
        k = n[0]
        out = []
        for which, rows in ((0, n[1]), (1, n[2])):
            ids, m = np.empty(rows, np.int64), np.empty((rows, k), np.float32)
            lib.model_io_get_matrix(r, which, ids.ctypes.data_as(_i64p), m.ctypes.data_as(_f32p))
            out += [ids, m]
        g = Generation(out[0], out[1], out[2], out[3])
        if n[3]:
            ku, kp, ki = np.empty(n[4], np.int64), np.empty(n[4] + 1, np.int64), np.empty(n[5], np.int64)
            lib.model_io_get_known(r, ku.ctypes.data_as(_i64p), kp.ctypes.data_as(_i64p), ki.ctypes.data_as(_i64p))
            g.known_user_ids, g.known_ptr, g.known_item_ids = ku, kp, ki
        g.item_tag_ids, g.user_tag_ids = np.empty(n[6], np.int64), np.empty(n[7], np.int64)
        lib.model_io_get_tags(r, 0, g.item_tag_ids.ctypes.data_as(_i64p))
        lib.model_io_get_tags(r, 1, g.user_tag_ids.ctypes.data_as(_i64p))
        return g
    finally:
        lib.model_io_reader_destroy(r)


def write_generation(g, path):
    """GenerationSerializer.writeGeneration: object stream inside gzip; the name must end in .gz."""
    if not path.endswith(".gz"):
        raise ValueError("File should end in .gz: %s" % path)
    with gzip.open(path, "wb") as f:
        f.write(to_bytes(g))


def read_generation(path):
    """GenerationSerializer.readGeneration (IOUtils.openMaybeDecompressing: gzip by extension)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        return from_bytes(f.read())
