"""Host mirror of the fold-in math over libmyrrix_foldin.so (include/myrrix_foldin.h):
Generation.recomputeState's two solvers and ServerRecommender.updateFeatures / foldInWeight /
buildAnonymousUserFeatures (online/src/net/myrrix/online/ServerRecommender.java:561-608,
865-907, 981-994). No Python fallback: the library does the arithmetic."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmyrrix_foldin.so")
(FOLDIN_OK, FOLDIN_E_ARG, FOLDIN_E_ILL_CONDITIONED, FOLDIN_E_SINGULAR, FOLDIN_E_NOT_READY,
 FOLDIN_E_NONFINITE, FOLDIN_E_OOM) = range(7)

_H = C.c_void_p
_f32p, _f64p, _i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
# Every symbol include/myrrix_foldin.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("foldin_create", C.c_int, [C.c_int32, _f64p, _f64p, C.c_double, C.c_double, C.POINTER(_H), _i32p, _i32p]),
    ("foldin_destroy", None, [_H]),
    ("foldin_weight", C.c_double, [_H, C.c_double, C.c_float]),
    ("foldin_solve", C.c_int, [_H, C.c_int32, _f32p, _f64p]),
    ("foldin_update_features", C.c_int, [_H, _f32p, _f32p, C.c_float]),
    ("foldin_update_many", C.c_int, [_H, _f32p, _f32p, _i32p, _i32p, _f32p, C.c_int64]),
    ("foldin_export_solver", C.c_int, [_H, C.c_int32, _f64p, _f64p, _i32p]),
    ("foldin_anonymous_user", C.c_int, [_H, _f32p, _f32p, C.c_int32, _f32p]),
]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmyrrix_foldin.so is not built (python myrrix-recommender_b200/build.py)")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


class IllConditionedSolverException(RuntimeError):
    """common/src/net/myrrix/common/math/IllConditionedSolverException.java"""


class SingularMatrixSolverException(RuntimeError):
    def __init__(self, apparent_rank, which):
        super().__init__("Apparent rank: %d" % apparent_rank)
        self.apparent_rank, self.which = apparent_rank, which


class NotReadyException(RuntimeError):
    pass


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, np.float64)


class FoldIn:
    """The solver state of one Generation (Generation.java:132-158) and the writes against it."""

    def __init__(self, features, xtx=None, yty=None, singularity_threshold=1e-5, learn_rate=1.0):
        self.lib, self.k = load(), int(features)
        self.h = _H()
        x, y = _f64(xtx), _f64(yty)
        for m in (x, y):
            assert m is None or m.shape == (self.k, self.k)
        which, rank = C.c_int32(-1), C.c_int32(0)
        rc = self.lib.foldin_create(self.k, None if x is None else x.ctypes.data_as(_f64p),
                                    None if y is None else y.ctypes.data_as(_f64p),
                                    singularity_threshold, learn_rate, C.byref(self.h), C.byref(which),
                                    C.byref(rank))
        if rc == FOLDIN_E_ILL_CONDITIONED:
            raise IllConditionedSolverException("infNorm < 1 (%s)" % ("X'X", "Y'Y")[which.value])
        if rc == FOLDIN_E_SINGULAR:
            raise SingularMatrixSolverException(rank.value, which.value)
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_create failed (%d)" % rc)

    def weight(self, estimate, value):
        return float(self.lib.foldin_weight(self.h, estimate, value))

    def solve(self, which, b):
        b = np.ascontiguousarray(b, np.float32)
        x = np.empty(self.k, np.float64)
        rc = self.lib.foldin_solve(self.h, which, b.ctypes.data_as(_f32p), x.ctypes.data_as(_f64p))
        if rc == FOLDIN_E_NOT_READY:
            raise NotReadyException()
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_solve failed (%d)" % rc)
        return x

    def update_features(self, user_features, item_features, value):
        """In place on two float32 arrays (rows of X and Y), like the reference."""
        for a in (user_features, item_features):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (self.k,)
        rc = self.lib.foldin_update_features(self.h, user_features.ctypes.data_as(_f32p),
                                             item_features.ctypes.data_as(_f32p), value)
        if rc == FOLDIN_E_NONFINITE:
            raise ArithmeticError("non-finite fold-in")
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_update_features failed (%d)" % rc)

    def update_many(self, X, Y, users, items, values=None):
        """A stream of writes applied in order, in place on X [n_users, k] and Y [n_items, k]."""
        for a in (X, Y):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape[1] == self.k
        u = np.ascontiguousarray(users, np.int32)
        i = np.ascontiguousarray(items, np.int32)
        assert u.shape == i.shape and (u.size == 0 or (0 <= u.min() and u.max() < len(X) and
                                                        0 <= i.min() and i.max() < len(Y)))
        v = None if values is None else np.ascontiguousarray(values, np.float32)
        rc = self.lib.foldin_update_many(self.h, X.ctypes.data_as(_f32p), Y.ctypes.data_as(_f32p),
                                         u.ctypes.data_as(_i32p), i.ctypes.data_as(_i32p),
                                         None if v is None else v.ctypes.data_as(_f32p), u.size)
        if rc == FOLDIN_E_NONFINITE:
            raise ArithmeticError("non-finite fold-in")
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_update_many failed (%d)" % rc)

    def anonymous_user(self, item_rows, values=None):
        rows = np.ascontiguousarray(item_rows, np.float32).reshape(-1, self.k)
        v = None if values is None else np.ascontiguousarray(values, np.float32)
        out = np.empty(self.k, np.float32)
        rc = self.lib.foldin_anonymous_user(self.h, rows.ctypes.data_as(_f32p),
                                            None if v is None else v.ctypes.data_as(_f32p), len(rows),
                                            out.ctypes.data_as(_f32p))
        if rc == FOLDIN_E_NOT_READY:
            raise NotReadyException()
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_anonymous_user failed (%d)" % rc)
        return out

    def export(self, which):
        """(qrt [k, k], rdiag [k], perm [k]) of solver `which` (0: X'X, 1: Y'Y), or None."""
        qrt = np.empty((self.k, self.k), np.float64)
        rdiag = np.empty(self.k, np.float64)
        perm = np.empty(self.k, np.int32)
        rc = self.lib.foldin_export_solver(self.h, which, qrt.ctypes.data_as(_f64p), rdiag.ctypes.data_as(_f64p),
                                           perm.ctypes.data_as(_i32p))
        if rc == FOLDIN_E_NOT_READY:
            return None
        if rc != FOLDIN_OK:
            raise RuntimeError("foldin_export_solver failed (%d)" % rc)
        return qrt, rdiag, perm

    def close(self):
        if self.h:
            self.lib.foldin_destroy(self.h)
            self.h = _H()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
