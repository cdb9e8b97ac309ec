"""Host mirror of the cold-start path over libmyrrix_init.so (include/myrrix_init.h):
RandomManager.getRandom (common/src/net/myrrix/common/random/RandomManager.java:41-68), the
commons-math3 MersenneTwister stream behind it, RandomUtils.randomUnitVector(FarFrom)
(common/.../random/RandomUtils.java:82-140) and AlternatingLeastSquares.constructInitialY
(online/.../als/AlternatingLeastSquares.java:264-335).  No Python fallback: the library draws."""
import ctypes as C
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmyrrix_init.so")
MYRRIX_INIT_OK, MYRRIX_INIT_E_ARG, MYRRIX_INIT_E_OOM = range(3)

_R = C.c_void_p
_f32p, _i32p, _i64p, _u8p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                             C.POINTER(C.c_uint8))
# Every symbol include/myrrix_init.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("myrrix_rng_create", _R, [C.c_int64]),
    ("myrrix_rng_create_by_array", _R, [_i32p, C.c_int32]),
    ("myrrix_rng_destroy", None, [_R]),
    ("myrrix_rng_next_bits", C.c_uint32, [_R, C.c_int32]),
    ("myrrix_rng_next_double", C.c_double, [_R]),
    ("myrrix_rng_next_gaussian", C.c_double, [_R]),
    ("myrrix_rng_next_int", C.c_int32, [_R, C.c_int32]),
    ("myrrix_random_unit_vector", C.c_int, [_R, C.c_int32, _f32p]),
    ("myrrix_random_unit_vector_far_from", C.c_int, [_R, C.c_int32, _f32p, C.c_int64, _f32p]),
    ("myrrix_construct_initial_y", C.c_int, [_R, C.c_int32, C.c_int64, C.c_int32, _f32p, _i64p, C.c_int64,
                                             _i64p, C.c_int64, _f32p, _u8p]),
]
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmyrrix_init.so is not built (python myrrix-recommender_b200/build.py)")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f32p)


class MersenneTwister:
    """org.apache.commons.math3.random.MersenneTwister(long seed) / (int[] seed)."""

    def __init__(self, seed=None):
        self.lib = load()
        if seed is None:  # MersenneTwister(): time + identity hash
            seed = int(time.time() * 1000) + id(self)
        if isinstance(seed, (list, tuple, np.ndarray)):
            key = np.ascontiguousarray(np.asarray(seed, dtype=np.int64).astype(np.uint32).view(np.int32))
            self.r = self.lib.myrrix_rng_create_by_array(key.ctypes.data_as(_i32p), key.size)
        else:
            seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            self.r = self.lib.myrrix_rng_create(seed - (1 << 64) if seed >= (1 << 63) else seed)
        if not self.r:
            raise MemoryError("myrrix_rng_create")

    def __del__(self):
        if getattr(self, "r", None):
            self.lib.myrrix_rng_destroy(self.r)
            self.r = None

    def next(self, bits):
        return int(self.lib.myrrix_rng_next_bits(self.r, bits))

    def nextDouble(self):
        return float(self.lib.myrrix_rng_next_double(self.r))

    def nextGaussian(self):
        return float(self.lib.myrrix_rng_next_gaussian(self.r))

    def nextInt(self, n):
        if n <= 0:
            raise ValueError("n must be strictly positive")  # NotStrictlyPositiveException
        return int(self.lib.myrrix_rng_next_int(self.r, n))


class RandomManager:
    """net.myrrix.common.random.RandomManager."""
    TEST_SEED = 1234567890
    _use_test_seed = False

    @classmethod
    def getRandom(cls):
        return MersenneTwister(cls.TEST_SEED) if cls._use_test_seed else MersenneTwister()

    @classmethod
    def useTestSeed(cls, on=True):
        cls._use_test_seed = bool(on)


def randomUnitVector(dimensions, random):
    out = np.empty(dimensions, dtype=np.float32)
    if load().myrrix_random_unit_vector(random.r, dimensions, _fp(out)) != MYRRIX_INIT_OK:
        raise ValueError("randomUnitVector")
    return out


def randomUnitVectorFarFrom(dimensions, farFrom, random):
    far = np.ascontiguousarray(np.asarray(farFrom, dtype=np.float32).reshape(-1, dimensions))
    out = np.empty(dimensions, dtype=np.float32)
    rc = load().myrrix_random_unit_vector_far_from(random.r, dimensions, _fp(far) if far.size else None,
                                                   far.shape[0], _fp(out))
    if rc != MYRRIX_INIT_OK:
        raise ValueError("randomUnitVectorFarFrom")
    return out


def construct_initial_y(random, features, n_rows, column_order, prev=None, prev_order=()):
    """constructInitialY on dense rows (see include/myrrix_init.h): (Y [n_rows][features], has_vector)."""
    y = np.zeros((n_rows, features), dtype=np.float32)
    has = np.zeros(n_rows, dtype=np.uint8)
    col = np.ascontiguousarray(column_order, dtype=np.int64)
    po = np.ascontiguousarray(prev_order, dtype=np.int64)
    pk = 0
    pv = None
    if po.size:
        pv = np.ascontiguousarray(prev, dtype=np.float32)
        pk = pv.shape[1]
    rc = load().myrrix_construct_initial_y(
        random.r, features, n_rows, pk, _fp(pv) if pv is not None else None,
        po.ctypes.data_as(_i64p) if po.size else None, po.size,
        col.ctypes.data_as(_i64p) if col.size else None, col.size, _fp(y), has.ctypes.data_as(_u8p))
    if rc == MYRRIX_INIT_E_OOM:
        raise MemoryError("constructInitialY")
    if rc != MYRRIX_INIT_OK:
        raise ValueError("constructInitialY: invalid argument")
    return y, has.astype(bool)
