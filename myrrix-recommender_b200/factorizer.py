"""Host-side mirror of the reference's factorizer interface over the C ABI.

  MatrixFactorizer          online/.../factorizer/MatrixFactorizer.java:31-77
  AlternatingLeastSquares   online/.../factorizer/als/AlternatingLeastSquares.java:66-543

Same names, argument meaning and error behaviour as the Java so the parity tests read
like the reference's own (AlternatingLeastSquaresTest.java:39-117).  The maps at the
boundary are plain dicts: FastByIDMap<FastByIDFloatMap> -> {long: {long: float}},
FastByIDMap<float[]> -> {long: float32 ndarray}.  All arithmetic happens on the GPU
through libmyrrix_als.so; this file only flattens IDs, owns the stop rule
(ALS.java:227-257) and maps status codes back onto the reference's exceptions.
"""
import ctypes as C
import math

import numpy as np

from . import _native as N
from . import initial_y

# JVM system properties the reference reads on this path, same names.
properties = {}

DEFAULT_FEATURES = 30                      # MatrixFactorizer.java:34
DEFAULT_ALPHA = 1.0                        # AlternatingLeastSquares.java:71
DEFAULT_LAMBDA = 0.1                       # :73
DEFAULT_CONVERGENCE_THRESHOLD = 0.001      # :74
DEFAULT_MAX_ITERATIONS = 30                # :75
NUM_USER_ITEMS_TO_TEST_CONVERGENCE = 100   # :78


class SolverException(RuntimeError):
    """common/.../math/SolverException.java"""


class SingularMatrixSolverException(SolverException):
    """common/.../math/SingularMatrixSolverException.java:23-52 (unchecked, carries apparentRank)."""

    def __init__(self, apparent_rank=0, message=""):
        super().__init__(message)
        self.apparent_rank = apparent_rank

    def getApparentRank(self):
        return self.apparent_rank


class ExecutionException(RuntimeError):
    """java.util.concurrent.ExecutionException stand-in for CUDA/NCCL/OOM failures."""


def _prop_bool(name, default=False):
    v = properties.get(name)
    return default if v is None else str(v).lower() == "true"


def _prop_float(name, default):
    v = properties.get(name)
    if v is None:
        return default
    f = float(v)
    if not math.isfinite(f):  # LangUtils.parseDouble (common/.../LangUtils.java:60-64)
        raise ValueError("Bad value: %s" % v)
    return f


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class NativeALS:
    """Thin RAII wrapper of one als_handle. Raises on every non-OK status."""

    def __init__(self, features, alpha=DEFAULT_ALPHA, lam=DEFAULT_LAMBDA, reconstruct_r=False,
                 loss_ignores_unspecified=False, singularity_threshold=1.0e-5, device=0,
                 kernel=N.ALS_KERNEL_AUTO):
        self.lib = N.load()
        cfg = N.AlsConfig()
        self.lib.als_config_default(C.byref(cfg))
        cfg.features = int(features)
        cfg.alpha = float(alpha)
        cfg.lambda_ = float(lam)
        cfg.reconstruct_r = int(bool(reconstruct_r))
        cfg.loss_ignores_unspecified = int(bool(loss_ignores_unspecified))
        cfg.singularity_threshold = float(singularity_threshold)
        cfg.device = int(device)
        cfg.kernel = int(kernel)
        self.h = C.c_void_p()
        self.features = int(features)
        rc = self.lib.als_create(C.byref(cfg), C.byref(self.h))
        if rc != N.ALS_OK:
            msg = self.lib.als_last_error(self.h).decode() if self.h else ""
            if self.h:
                self.lib.als_destroy(self.h)
                self.h = C.c_void_p()
            if rc == N.ALS_E_ARG:
                raise ValueError("als_create: invalid argument %s" % msg)
            raise ExecutionException("als_create failed: %s %s" % (N.STATUS_NAMES[rc], msg))
        self.n_users = self.n_items = 0
        self.rank, self.world = 0, 1

    def close(self):
        if getattr(self, "h", None):
            self.lib.als_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc == N.ALS_OK:
            return
        msg = self.lib.als_last_error(self.h).decode()
        if rc == N.ALS_E_SINGULAR:
            raise SingularMatrixSolverException(self.lib.als_singular_rank(self.h), msg)
        if rc == N.ALS_E_NONFINITE:
            raise SolverException(msg)
        if rc == N.ALS_E_ARG:
            raise ValueError(msg)
        raise ExecutionException("%s: %s" % (N.STATUS_NAMES[rc], msg))

    # -- data ---------------------------------------------------------------
    def set_interactions(self, n_users, n_items, row_ptr, col_idx, val, by_column=None):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float32)
        if col_idx.size and (col_idx.min() < 0 or col_idx.max() >= n_items):
            raise ValueError("column index out of range")
        if row_ptr.size < 1 or row_ptr[0] != 0 or np.any(np.diff(row_ptr) < 0):
            raise ValueError("row_ptr must start at 0 and be non-decreasing")
        self.check(self.lib.als_set_interactions(
            self.h, n_users, n_items, row_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
            col_idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(val)))
        if by_column is not None:
            cp, ri, cv = by_column
            cp = np.ascontiguousarray(cp, dtype=np.int64)
            ri = np.ascontiguousarray(ri, dtype=np.int32)
            cv = np.ascontiguousarray(cv, dtype=np.float32)
            if ri.size and (ri.min() < 0 or ri.max() >= n_users):
                raise ValueError("row index out of range")
            if cp.size < 1 or cp[0] != 0 or np.any(np.diff(cp) < 0):
                raise ValueError("col_ptr must start at 0 and be non-decreasing")
            self.check(self.lib.als_set_interactions_by_column(
                self.h, cp.ctypes.data_as(C.POINTER(C.c_int64)),
                ri.ctypes.data_as(C.POINTER(C.c_int32)), _fp(cv)))
        self.n_users, self.n_items = int(n_users), int(n_items)

    def set_present_empty_rows(self, which, rows):
        """Rows that are keys of RbyRow (which=0) / RbyColumn (which=1) with no entries left
        (InputFilesReader.removeSmall): the reference solves them to the zero vector."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.check(self.lib.als_set_present_empty_rows(
            self.h, int(which), rows.ctypes.data_as(C.POINTER(C.c_int32)), rows.size))

    def set_interactions_device(self, n_users, n_items, d_row_ptr, d_col_idx, d_val):
        self.check(self.lib.als_set_interactions_device(self.h, n_users, n_items, d_row_ptr,
                                                        d_col_idx, d_val))
        self.n_users, self.n_items = int(n_users), int(n_items)

    def synth_interactions(self, n_users, n_items, nnz_per_user, seed=1234567890, neg_fraction=0.0):
        self.check(self.lib.als_synth_interactions(self.h, n_users, n_items, nnz_per_user, seed,
                                                   neg_fraction))
        self.n_users, self.n_items = int(n_users), int(n_items)

    def synth_interactions_powerlaw(self, n_users, n_items, mean_nnz, max_nnz=20000, seed=1234567890,
                                    neg_fraction=0.0):
        self.check(self.lib.als_synth_interactions_powerlaw(self.h, n_users, n_items, float(mean_nnz), int(max_nnz),
                                                            seed, neg_fraction))
        self.n_users, self.n_items = n_users, n_items

    def synth_y0(self, seed=1234567890):
        self.check(self.lib.als_synth_y0(self.h, seed))

    def set_y(self, y):
        y = np.ascontiguousarray(y, dtype=np.float32)
        if y.shape != (self.n_items, self.features):
            raise ValueError("Y must be n_items x features")
        self.check(self.lib.als_set_y(self.h, _fp(y)))

    def set_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.shape != (self.n_users, self.features):
            raise ValueError("X must be n_users x features")
        self.check(self.lib.als_set_x(self.h, _fp(x)))

    def get_x(self, out=None):
        if out is None:
            out = np.empty((self.n_users, self.features), dtype=np.float32)
        self.check(self.lib.als_get_x(self.h, _fp(out)))
        return out

    def get_y(self, out=None):
        if out is None:
            out = np.empty((self.n_items, self.features), dtype=np.float32)
        self.check(self.lib.als_get_y(self.h, _fp(out)))
        return out

    def get_rows(self, which, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.empty((rows.size, self.features), dtype=np.float32)
        self.check(self.lib.als_get_rows(self.h, 0 if which in (0, "x", "X") else 1,
                                         rows.ctypes.data_as(C.POINTER(C.c_int32)), rows.size, _fp(out)))
        return out

    def get_interactions(self, by_column=False):
        from .sharding import local_block
        info = self.info()
        b, e = local_block(self.n_items if by_column else self.n_users, self.rank, self.world)
        rows = e - b
        if self.world > 1:  # shard sizes differ per orientation: go through the slice getter
            return self.get_interaction_rows(0, rows, by_column=by_column,
                                             capacity=int(info.nnz) * 4 + 1024)
        ptr = np.empty(rows + 1, dtype=np.int64)
        idx = np.empty(info.nnz, dtype=np.int32)
        val = np.empty(info.nnz, dtype=np.float32)
        fn = self.lib.als_get_interactions_by_column if by_column else self.lib.als_get_interactions
        self.check(fn(self.h, ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                      idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(val)))
        return ptr, idx, val

    def get_interaction_rows(self, first_row, n_rows, by_column=False, capacity=None):
        ptr = np.empty(n_rows + 1, dtype=np.int64)
        if capacity is None:
            capacity = self.info().nnz
        idx = np.empty(capacity, dtype=np.int32)
        val = np.empty(capacity, dtype=np.float32)
        self.check(self.lib.als_get_interaction_rows(
            self.h, int(by_column), first_row, n_rows, ptr.ctypes.data_as(C.POINTER(C.c_int64)),
            idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(val), capacity))
        n = int(ptr[-1])
        return ptr, idx[:n].copy(), val[:n].copy()

    # -- compute ------------------------------------------------------------
    def half_x(self):
        self.check(self.lib.als_half_x(self.h))

    def half_y(self):
        self.check(self.lib.als_half_y(self.h))

    def iterate(self, n):
        self.check(self.lib.als_iterate(self.h, n))

    def sync(self):
        self.check(self.lib.als_sync(self.h))

    def call(self, test_users, test_items, max_iterations, convergence_threshold, random_y, x_is_empty=True):
        """The iteration loop of AlternatingLeastSquares.call with the stop rule on the device
        (als_call): (iterations run, last convergence value)."""
        tu, pu = self._i32(test_users)
        ti, pi = self._i32(test_items)
        n_it, val = C.c_int32(0), C.c_double(float("nan"))
        self.check(self.lib.als_call(self.h, pu, tu.size, pi, ti.size, int(max_iterations),
                                     float(convergence_threshold), int(bool(random_y)), int(bool(x_is_empty)),
                                     C.byref(n_it), C.byref(val)))
        return n_it.value, val.value

    def probe(self, users, items):
        users = np.ascontiguousarray(users, dtype=np.int32)
        items = np.ascontiguousarray(items, dtype=np.int32)
        out = np.zeros((users.size, items.size), dtype=np.float64)
        self.check(self.lib.als_probe(self.h, users.ctypes.data_as(C.POINTER(C.c_int32)), users.size,
                                      items.ctypes.data_as(C.POINTER(C.c_int32)), items.size,
                                      out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    # -- the live model between builds (fold-in on the resident rows) ----------------------------
    def recompute_state(self, singularity_threshold=1.0e-5, learn_rate=1.0, xtx=True, yty=True):
        """Generation.recomputeState: Gramians from the resident factors (GPU), the generation's
        solvers (libmyrrix_foldin.so: infNorm guard, pivoted QR), a copy of their factors on the device.
        Returns the host-side FoldIn (anonymous users, single solves)."""
        from .foldin import FoldIn
        fi = FoldIn(self.features, self.gramian("x") if xtx else None, self.gramian("y") if yty else None,
                    singularity_threshold=singularity_threshold, learn_rate=learn_rate)
        for which in (0, 1):
            st = fi.export(which)
            if st is None:
                self.check(self.lib.als_set_fold_in_state(self.h, which, None, None, None, learn_rate))
            else:
                qrt, rdiag, perm = st
                self.check(self.lib.als_set_fold_in_state(
                    self.h, which, qrt.ctypes.data_as(C.POINTER(C.c_double)),
                    rdiag.ctypes.data_as(C.POINTER(C.c_double)), perm.ctypes.data_as(C.POINTER(C.c_int32)),
                    learn_rate))
        return fi

    def fold_in(self, users, items, values=None):
        """ServerRecommender.updateFeatures for a stream of writes, in order, on the resident rows."""
        u, pu = self._i32(users)
        i, pi = self._i32(items)
        v = None if values is None else np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        if u.size != i.size or (v is not None and v.size != u.size):
            raise ValueError("users / items / values differ in length")
        self.check(self.lib.als_fold_in(self.h, pu, pi, None if v is None else _fp(v), u.size))

    # -- top-N scoring on the resident model (include/myrrix_als.h, csrc/topn.cuh) --------------
    @staticmethod
    def _i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
        return a, a.ctypes.data_as(C.POINTER(C.c_int32))

    def recommend(self, users, how_many, consider_known_items=False, exclude=()):
        """ServerRecommender.recommendToMany on dense indices: (items, values), best first."""
        users, pu = self._i32(np.atleast_1d(users))
        ex, pe = self._i32(exclude)
        items = np.empty(how_many, dtype=np.int32)
        values = np.empty(how_many, dtype=np.float32)
        n = C.c_int32(0)
        self.check(self.lib.als_recommend(self.h, pu, users.size, how_many, int(bool(consider_known_items)),
                                          pe if ex.size else None, ex.size,
                                          items.ctypes.data_as(C.POINTER(C.c_int32)), _fp(values), C.byref(n)))
        return items[:n.value].copy(), values[:n.value].copy()

    def recommend_batch(self, users, how_many, consider_known_items=False):
        """One single-user query per entry (AllRecommendations): items [n][how_many] (-1 beyond
        counts[n]), values, counts."""
        users, pu = self._i32(users)
        items = np.empty((users.size, how_many), dtype=np.int32)
        values = np.empty((users.size, how_many), dtype=np.float32)
        counts = np.empty(users.size, dtype=np.int32)
        self.check(self.lib.als_recommend_batch(self.h, pu, users.size, how_many, int(bool(consider_known_items)),
                                                items.ctypes.data_as(C.POINTER(C.c_int32)), _fp(values),
                                                counts.ctypes.data_as(C.POINTER(C.c_int32))))
        return items, values, counts

    def top_n(self, which, features, how_many, exclude=()):
        """Best rows of Y (which = 1 / "y") or X for the mean score of the given feature vectors."""
        f = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, self.features)
        ex, pe = self._i32(exclude)
        ids = np.empty(how_many, dtype=np.int32)
        values = np.empty(how_many, dtype=np.float32)
        n = C.c_int32(0)
        self.check(self.lib.als_top_n(self.h, 0 if which in (0, "x", "X") else 1, _fp(f), f.shape[0],
                                      pe if ex.size else None, ex.size, how_many,
                                      ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(values), C.byref(n)))
        return ids[:n.value].copy(), values[:n.value].copy()

    def gramian(self, which):
        out = np.zeros((self.features, self.features), dtype=np.float64)
        self.check(self.lib.als_gramian(self.h, 0 if which in (0, "x", "X") else 1,
                                        out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def set_stream(self, cuda_stream):
        self.check(self.lib.als_set_stream(self.h, C.c_void_p(cuda_stream)))

    def info(self):
        info = N.AlsInfo()
        self.check(self.lib.als_get_info(self.h, C.byref(info)))
        return info

    def profile(self, on=True):
        self.check(self.lib.als_profile_enable(self.h, int(on)))

    def timings(self, reset=False):
        t = N.AlsTimings()
        self.check(self.lib.als_get_timings(self.h, C.byref(t), int(reset)))
        return t

    def comm_init(self, rank, world_size, unique_id):
        buf = None if unique_id is None else C.create_string_buffer(bytes(unique_id), len(unique_id))
        self.check(self.lib.als_comm_init(self.h, rank, world_size, buf))
        self.rank, self.world = int(rank), int(world_size)


def comm_unique_id():
    lib = N.load()
    n = lib.als_comm_unique_id_size()
    buf = C.create_string_buffer(n)
    rc = lib.als_comm_get_unique_id(buf)
    if rc != N.ALS_OK:
        raise ExecutionException("als_comm_get_unique_id: %s" % N.STATUS_NAMES[rc])
    return buf.raw


class DoubleWeightedMean:
    """common/.../stats/DoubleWeightedMean.java:33-101 (increment only)."""

    def __init__(self):
        self.totalWeight = 0.0
        self.mean = float("nan")

    def increment(self, datum, weight=1.0):
        old = self.totalWeight
        self.totalWeight += weight
        if old <= 0:
            self.mean = datum
        else:
            self.mean = self.mean * old / self.totalWeight + datum * weight / self.totalWeight

    def getResult(self):
        return self.mean


class MatrixFactorizer:
    """online/.../factorizer/MatrixFactorizer.java:31-77"""
    DEFAULT_FEATURES = DEFAULT_FEATURES

    def call(self):
        raise NotImplementedError

    def setPreviousX(self, previousX):
        raise NotImplementedError

    def setPreviousY(self, previousY):
        raise NotImplementedError

    def getX(self):
        raise NotImplementedError

    def getY(self):
        raise NotImplementedError


class AlternatingLeastSquares(MatrixFactorizer):
    """Drop-in for net.myrrix.online.factorizer.als.AlternatingLeastSquares.

    `RbyRow` / `RbyColumn`: {userID: {itemID: strength}} / {itemID: {userID: strength}}
    (FastByIDMap<FastByIDFloatMap>, ctor at AlternatingLeastSquares.java:132-147).
    """

    def __init__(self, RbyRow, RbyColumn, features=DEFAULT_FEATURES,
                 estimateErrorConvergenceThreshold=DEFAULT_CONVERGENCE_THRESHOLD,
                 maxIterations=DEFAULT_MAX_ITERATIONS, device=0, kernel=N.ALS_KERNEL_AUTO,
                 random=None):
        # Preconditions, ALS.java:137-141
        if RbyRow is None or RbyColumn is None:
            raise TypeError("RbyRow/RbyColumn must not be null")
        if not features > 0:
            raise ValueError("features must be positive: %s" % features)
        if not (0.0 < estimateErrorConvergenceThreshold < 1.0):
            raise ValueError("threshold must be in (0,1): %s" % estimateErrorConvergenceThreshold)
        self.RbyRow = RbyRow
        self.RbyColumn = RbyColumn
        self.features = int(features)
        self.estimateErrorConvergenceThreshold = float(estimateErrorConvergenceThreshold)
        self.maxIterations = int(maxIterations)
        self.X = None
        self.Y = None
        self.previousY = None
        self.device = device
        self.kernel = kernel
        self.random = random if random is not None else initial_y.RandomManager.getRandom()
        self.iterationsRun = 0
        self.lastConvergenceValue = float("nan")

    def getX(self):
        return self.X

    def getY(self):
        return self.Y

    def setPreviousX(self, previousX):
        pass  # "Does nothing." ALS.java:162-165

    def setPreviousY(self, previousY):
        self.previousY = previousY

    # -- constructInitialY, ALS.java:264-335 ---------------------------------
    def _construct_initial_y(self, previousY):
        """Dense rows for every key of previousY and of RbyColumn, in their maps' iteration order,
        then libmyrrix_init.so (include/myrrix_init.h) draws from the MersenneTwister stream exactly
        like constructInitialY / randomUnitVectorFarFrom; the dict comes back keyed like the maps."""
        k = self.features
        if not previousY:
            previousY = {}
        keys = list(previousY.keys())
        index = {key: i for i, key in enumerate(keys)}
        for item_id in self.RbyColumn.keys():
            if item_id not in index:
                index[item_id] = len(keys)
                keys.append(item_id)
        prev = None
        if previousY:
            old = len(next(iter(previousY.values())))
            prev = np.zeros((len(keys), old), dtype=np.float32)
            for key, vec in previousY.items():
                prev[index[key]] = vec
            if old == k:
                randomY = previousY  # adopted in place (ALS.java:304-308): only new keys are added
            else:
                randomY = {}
        else:
            randomY = {}
        y, has = initial_y.construct_initial_y(
            self.random, k, len(keys), [index[i] for i in self.RbyColumn.keys()], prev,
            [index[key] for key in previousY.keys()])
        for key in keys:
            if key not in randomY and has[index[key]]:
                randomY[key] = y[index[key]].copy()
        return randomY

    # -- flattening: long IDs -> dense indices, maps -> CSR --------------------
    @staticmethod
    def _flatten(R, row_index, col_index):
        n_rows = len(row_index)
        counts = np.zeros(n_rows + 1, dtype=np.int64)
        for rid, row in R.items():
            counts[row_index[rid] + 1] = len(row)
        ptr = np.cumsum(counts)
        idx = np.empty(int(ptr[-1]), dtype=np.int32)
        val = np.empty(int(ptr[-1]), dtype=np.float32)
        for rid, row in R.items():
            o = int(ptr[row_index[rid]])
            for j, (cid, v) in enumerate(row.items()):
                idx[o + j] = col_index[cid]
                val[o + j] = v
        return ptr, idx, val

    def _choose_about_n(self, keys, n):
        # RandomUtils.chooseAboutNFromStream (RandomUtils.java:202-217)
        keys = list(keys)
        if n < len(keys):
            rate = float(n) / len(keys)
            keys = [key for key in keys if self.random.nextDouble() < rate]
        return keys

    def call(self):
        k = self.features
        randomY = not self.previousY  # ALS.java:181
        Ymap = self._construct_initial_y(self.previousY)
        user_ids = list(self.RbyRow.keys())
        # every row of Y counts in Y^T Y, including stale rows absent from RbyColumn
        item_ids = list(Ymap.keys())
        uindex = {u: i for i, u in enumerate(user_ids)}
        iindex = {it: i for i, it in enumerate(item_ids)}
        for it in self.RbyColumn.keys():
            if it not in iindex:  # cannot happen after constructInitialY; keep the invariant explicit
                raise AssertionError("item without a Y row")
        r_ptr, r_idx, r_val = self._flatten(self.RbyRow, uindex, iindex)
        c_full = {it: self.RbyColumn.get(it, {}) for it in item_ids}
        c_ptr, c_idx, c_val = self._flatten(c_full, iindex, uindex)
        if r_ptr[-1] != c_ptr[-1]:
            raise ValueError("RbyRow and RbyColumn disagree")

        alpha = _prop_float("model.als.alpha", DEFAULT_ALPHA)
        lam = _prop_float("model.als.lambda", DEFAULT_LAMBDA)
        als = NativeALS(k, alpha=alpha, lam=lam,
                        reconstruct_r=_prop_bool("model.reconstructRMatrix"),
                        loss_ignores_unspecified=_prop_bool("model.lossIgnoresUnspecified"),
                        singularity_threshold=_prop_float("common.matrix.singularityThreshold", 1e-5),
                        device=self.device, kernel=self.kernel)
        try:
            n_users, n_items = len(user_ids), len(item_ids)
            if n_users == 0 or n_items == 0:
                self.X, self.Y = {}, dict(Ymap)
                return None
            als.set_interactions(n_users, n_items, r_ptr, r_idx, r_val, by_column=(c_ptr, c_idx, c_val))
            # keys whose maps were emptied by removeSmall (InputFilesReader.java:202-211) are still
            # walked by addWorkers (ALS.java:391-410): W = G, b = 0 -> the zero vector, from the
            # first half on (so they stop feeding X^T X / Y^T Y like in the reference)
            pe_u = [uindex[u] for u, row in self.RbyRow.items() if len(row) == 0]
            pe_i = [iindex[it] for it, row in self.RbyColumn.items() if len(row) == 0]
            if pe_u:
                als.set_present_empty_rows(0, pe_u)
            if pe_i:
                als.set_present_empty_rows(1, pe_i)
            Y0 = np.stack([np.asarray(Ymap[it], dtype=np.float32) for it in item_ids])
            als.set_y(Y0)

            def publish():
                Xd, Yd = als.get_x(), als.get_y()
                self.X = {u: Xd[i].copy() for i, u in enumerate(user_ids)}
                newY = {it: Yd[i].copy() for i, it in enumerate(item_ids)}
                if Ymap is self.previousY:  # adopted in place (ALS.java:304-308)
                    for it, v in newY.items():
                        Ymap[it] = v
                    self.Y = Ymap
                else:
                    self.Y = newY

            if not _prop_bool("model.als.iterate", True):  # ALS.java:196-204
                als.half_x()
                als.sync()
                publish()
                return None

            present_items = [it for it in self.RbyColumn.keys()]
            test_users = self._choose_about_n(user_ids, NUM_USER_ITEMS_TO_TEST_CONVERGENCE)
            test_items = self._choose_about_n(present_items, NUM_USER_ITEMS_TO_TEST_CONVERGENCE)
            tu = np.array([uindex[u] for u in test_users], dtype=np.int32)
            ti = np.array([iindex[it] for it in test_items], dtype=np.int32)
            estimates = np.zeros((tu.size, ti.size), dtype=np.float64)  # X empty: stay 0 (:215-223)

            if not _prop_bool("model.als.hostStopRule"):
                # the loop below, run by the library with the statistic computed on the device
                self.iterationsRun, conv = als.call(tu, ti, self.maxIterations,
                                                    self.estimateErrorConvergenceThreshold, randomY)
                if not (self.maxIterations > 0 and self.iterationsRun >= self.maxIterations):
                    self.lastConvergenceValue = conv
                publish()
                return None

            iterationNumber = 0
            while True:
                als.half_x()   # iterateXFromY, :228
                als.half_y()   # iterateYFromX, :229
                als.sync()     # surfaces SingularMatrixSolverException like Future.get() (:349)
                new = als.probe(tu, ti)
                mean = DoubleWeightedMean()
                for i in range(tu.size):
                    for j in range(ti.size):
                        nv = float(new[i, j])
                        mean.increment(abs(nv - estimates[i, j]), max(0.0, nv))
                estimates = new
                iterationNumber += 1
                self.iterationsRun = iterationNumber
                if self.maxIterations > 0 and iterationNumber >= self.maxIterations:
                    break
                conv = mean.getResult()
                self.lastConvergenceValue = conv
                if not math.isfinite(conv):
                    break
                if not (randomY and iterationNumber == 1) and conv < self.estimateErrorConvergenceThreshold:
                    break
            publish()
        finally:
            als.close()
        return None
