"""ctypes binding of oracle/als_oracle.c -- TEST INFRASTRUCTURE ONLY.

The oracle restates the reference's Java ALS arithmetic in fp64 on the CPU
(citations in als_oracle.c).  It is the checker for the CUDA path and the timed
CPU baseline in bench.py; it is never a fallback for the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "als_oracle.c")

OK, E_SINGULAR, E_OOM, E_ARG = 0, 1, 5, 6


def _cpu_has(*flags):
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    have = set(line.split(":", 1)[1].split())
                    return all(fl in have for fl in flags)
    except OSError:
        pass
    return False


def build(force=False):
    """Compile als_oracle.c -> libals_oracle.so (x86-64-v3) or a generic build when
    the host CPU lacks AVX2/FMA (the prebuilt .so travels to the GPU box)."""
    v3 = _cpu_has("avx2", "fma", "bmi2")
    name = "libals_oracle.so" if v3 else "libals_oracle_generic.so"
    out = os.path.join(_HERE, name)
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(_SRC):
        march = "x86-64-v3" if v3 else "x86-64"
        cmd = ["gcc", "-O3", "-march=" + march, "-ffp-contract=off", "-fPIC", "-Wall", "-pthread",
               "-shared", "-o", out, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i64p, i32p, f32p, f64p = (C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float), C.POINTER(C.c_double))
        L.oracle_transpose_times_self.argtypes = [f32p, C.c_int64, C.c_int, f64p]
        L.oracle_transpose_times_self.restype = None
        L.oracle_dot.argtypes = [f32p, f32p, C.c_int]
        L.oracle_dot.restype = C.c_double
        L.oracle_norm.argtypes = [f32p, C.c_int]
        L.oracle_norm.restype = C.c_double
        L.oracle_solve.argtypes = [f64p, f64p, C.c_int, C.c_double, f32p, C.POINTER(C.c_int)]
        L.oracle_solve.restype = C.c_int
        L.oracle_als_half.argtypes = [i64p, i32p, f32p, C.c_int64, f32p, f64p, C.c_int, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, f32p,
                                      C.POINTER(C.c_int)]
        L.oracle_als_half.restype = C.c_int
        u8p = C.POINTER(C.c_uint8)
        L.oracle_als_half_p.argtypes = L.oracle_als_half.argtypes + [u8p]
        L.oracle_als_half_p.restype = C.c_int
        L.oracle_convergence_probe.argtypes = [f32p, f32p, C.c_int, i32p, C.c_int, i32p, C.c_int, f64p]
        L.oracle_convergence_probe.restype = C.c_double
        L.oracle_als_run.argtypes = [i64p, i32p, f32p, C.c_int64, i64p, i32p, f32p, C.c_int64,
                                     C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                     C.c_double, C.c_int, C.c_int, i32p, C.c_int, i32p, C.c_int,
                                     C.c_int, f32p, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     f64p]
        L.oracle_als_run.restype = C.c_int
        L.oracle_als_run_p.argtypes = L.oracle_als_run.argtypes + [u8p, u8p]
        L.oracle_als_run_p.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class SingularMatrixError(Exception):
    """Mirrors SingularMatrixSolverException (common/.../math/SingularMatrixSolverException.java)."""

    def __init__(self, apparent_rank):
        super().__init__("Apparent rank: %d" % apparent_rank)
        self.apparent_rank = apparent_rank


def transpose_times_self(M):
    M = np.ascontiguousarray(M, dtype=np.float32)
    n, k = M.shape
    G = np.zeros((k, k), dtype=np.float64)
    lib().oracle_transpose_times_self(_p(M, C.c_float), n, k, _p(G, C.c_double))
    return G


def dot(x, y):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    return lib().oracle_dot(_p(x, C.c_float), _p(y, C.c_float), x.size)


def norm(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return lib().oracle_norm(_p(x, C.c_float), x.size)


def solve(W, b, threshold=1e-5):
    W = np.ascontiguousarray(W, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    k = b.size
    out = np.zeros(k, dtype=np.float32)
    rank = C.c_int(0)
    rc = lib().oracle_solve(_p(W, C.c_double), _p(b, C.c_double), k, threshold, _p(out, C.c_float),
                            C.byref(rank))
    if rc == E_SINGULAR:
        raise SingularMatrixError(rank.value)
    if rc != OK:
        raise RuntimeError("oracle_solve rc=%d" % rc)
    return out


def csr_transpose(row_ptr, col_idx, val, n_cols):
    """CSR -> CSR of the transpose; within each column entries keep ascending row order."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col_idx = np.asarray(col_idx, dtype=np.int32)
    val = np.asarray(val, dtype=np.float32)
    n_rows = row_ptr.size - 1
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(row_ptr))
    order = np.argsort(col_idx, kind="stable")
    t_ptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(np.bincount(col_idx, minlength=n_cols), out=t_ptr[1:])
    return t_ptr, np.ascontiguousarray(rows[order]), np.ascontiguousarray(val[order])


def _mask(present, n):
    if present is None:
        return None, None
    m = np.zeros(n, dtype=np.uint8)
    m[np.asarray(present, dtype=np.int64)] = 1
    return m, _p(m, C.c_uint8)


def als_half(row_ptr, col_idx, val, M, G, out, alpha=1.0, lam=0.1, reconstruct_r=False,
             loss_ignores_unspecified=False, threshold=1e-5, n_threads=1, present=None):
    """out[u] = solve(W_u, b_u) for every row u with entries (in place on `out`).  `present`:
    indices of rows without entries that are nevertheless keys of the reference's map
    (solved as W = G, b = 0, AlternatingLeastSquares.java:391-410)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    G = np.ascontiguousarray(G, dtype=np.float64)
    assert out.dtype == np.float32 and out.flags.c_contiguous
    k = M.shape[1]
    rank = C.c_int(0)
    keep, pm = _mask(present, row_ptr.size - 1)
    rc = lib().oracle_als_half_p(_p(row_ptr, C.c_int64), _p(col_idx, C.c_int32), _p(val, C.c_float),
                                 row_ptr.size - 1, _p(M, C.c_float), _p(G, C.c_double), k, alpha, lam,
                                 int(reconstruct_r), int(loss_ignores_unspecified), threshold,
                                 n_threads, _p(out, C.c_float), C.byref(rank), pm)
    if rc == E_SINGULAR:
        raise SingularMatrixError(rank.value)
    if rc != OK:
        raise RuntimeError("oracle_als_half rc=%d" % rc)
    return out


def als_run(row_ptr, col_idx, val, n_items, Y0, alpha=1.0, lam=0.1, reconstruct_r=False,
            loss_ignores_unspecified=False, threshold=1e-5, convergence_threshold=0.001,
            max_iterations=30, random_y=False, test_users=None, test_items=None, n_threads=1,
            t_csr=None, present_users=None, present_items=None):
    """Full AlternatingLeastSquares.call() restatement. Returns (X, Y, iterations, conv)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    n_users = row_ptr.size - 1
    Y = np.array(Y0, dtype=np.float32, order="C", copy=True)
    assert Y.shape[0] == n_items
    k = Y.shape[1]
    X = np.zeros((n_users, k), dtype=np.float32)
    if t_csr is None:
        t_csr = csr_transpose(row_ptr, col_idx, val, n_items)
    t_ptr, t_idx, t_val = [np.ascontiguousarray(a) for a in t_csr]
    if test_users is None:
        present = np.nonzero(np.diff(row_ptr) > 0)[0]
        test_users = present[:100] if present.size > 100 else present
    if test_items is None:
        present = np.nonzero(np.diff(t_ptr) > 0)[0]
        test_items = present[:100] if present.size > 100 else present
    test_users = np.ascontiguousarray(test_users, dtype=np.int32)
    test_items = np.ascontiguousarray(test_items, dtype=np.int32)
    its, rank, conv = C.c_int(0), C.c_int(0), C.c_double(float("nan"))
    keep_u, pu = _mask(present_users, n_users)
    keep_i, pi = _mask(present_items, n_items)
    rc = lib().oracle_als_run_p(_p(row_ptr, C.c_int64), _p(col_idx, C.c_int32), _p(val, C.c_float),
                              n_users, _p(t_ptr, C.c_int64), _p(t_idx, C.c_int32),
                              _p(t_val, C.c_float), n_items, k, alpha, lam, int(reconstruct_r),
                              int(loss_ignores_unspecified), threshold, convergence_threshold,
                              max_iterations, int(random_y), _p(test_users, C.c_int32),
                              test_users.size, _p(test_items, C.c_int32), test_items.size,
                              n_threads, _p(X, C.c_float), _p(Y, C.c_float), C.byref(its),
                              C.byref(rank), C.byref(conv), pu, pi)
    if rc == E_SINGULAR:
        raise SingularMatrixError(rank.value)
    if rc != OK:
        raise RuntimeError("oracle_als_run rc=%d" % rc)
    return X, Y, its.value, conv.value
