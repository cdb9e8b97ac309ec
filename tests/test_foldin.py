"""Fold-in math: libmyrrix_foldin.so against oracle/foldin_oracle.py (numpy
LU in fp64 -- an independent route from the product's RRQR) and against the ALS oracle's RRQR
restatement for the solver semantics (singular threshold, apparent rank). CPU only."""
import os

import numpy as np
import pytest

from myrrix_recommender_b200 import foldin as FI
from oracle import foldin_oracle as FO
from oracle import oracle as O


def _model(rng, n_users, n_items, k):
    X = (rng.standard_normal((n_users, k)) * 0.3).astype(np.float32)
    Y = (rng.standard_normal((n_items, k)) * 0.3).astype(np.float32)
    return X, Y, O.transpose_times_self(X), O.transpose_times_self(Y)


def test_symbols():
    lib = FI.load()
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "myrrix_foldin.h")).read()
    for name, _, _ in FI.SYMBOLS:
        assert hasattr(lib, name) and name + "(" in hdr


@pytest.mark.parametrize("k", [2, 3, 16, 30, 64, 100])
def test_solve_matches_lu_and_rrqr_oracle(k):
    rng = np.random.default_rng(k)
    X, Y, xtx, yty = _model(rng, 500 + 10 * k, 300 + 5 * k, k)
    with FI.FoldIn(k, xtx, yty) as f:
        for which, G in ((0, xtx), (1, yty)):
            b = rng.standard_normal(k).astype(np.float32)
            x = f.solve(which, b)
            ref = np.linalg.solve(G, b.astype(np.float64))
            assert np.abs(x - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300)
            xo = O.solve(G, b.astype(np.float64))          # RRQR restatement, cast to float32
            assert np.abs(x.astype(np.float32) - xo).max() <= 1e-6 * np.abs(xo).max()


def test_fold_in_weight_branches():
    with FI.FoldIn(2, np.eye(2) * 2, np.eye(2) * 2, learn_rate=0.5) as f:
        for est in (-0.5, 0.0, 0.3, 0.999, 1.0, 1.7):
            for val in (-3.0, -0.25, 0.0, 0.25, 1.0, 40.0):
                assert f.weight(est, val) == FO.fold_in_weight(est, val, 0.5), (est, val)
        assert f.weight(0.0, 1.0) == 0.25 and f.weight(1.0, 1.0) == 0.0 and f.weight(0.5, -1.0) == -0.125


@pytest.mark.parametrize("k", [3, 30, 64])
def test_update_features_stream(k):
    """A stream of online writes against one generation: rows drift exactly as the restatement says."""
    rng = np.random.default_rng(100 + k)
    X, Y, xtx, yty = _model(rng, 400, 250, k)
    Xo, Yo = X.copy(), Y.copy()
    with FI.FoldIn(k, xtx, yty) as f:
        for _ in range(300):
            u, i = int(rng.integers(len(X))), int(rng.integers(len(Y)))
            v = float(rng.choice([-2.0, -1.0, 0.5, 1.0, 3.0]))
            f.update_features(X[u], Y[i], v)
            Xo[u], Yo[i] = FO.update_features(Xo[u], Yo[i], v, xtx, yty)
    for a, b in ((X, Xo), (Y, Yo)):
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    # the same stream through the batched entry point gives the same bits as one call per write
    rng = np.random.default_rng(100 + k)
    X2, Y2, _, _ = _model(rng, 400, 250, k)
    us, its, vs = [], [], []
    for _ in range(300):
        us.append(int(rng.integers(len(X2)))); its.append(int(rng.integers(len(Y2))))
        vs.append(float(rng.choice([-2.0, -1.0, 0.5, 1.0, 3.0])))
    with FI.FoldIn(k, xtx, yty) as f:
        f.update_many(X2, Y2, us, its, vs)
    assert np.array_equal(X2, X) and np.array_equal(Y2, Y)
    assert np.abs(X - _model(np.random.default_rng(100 + k), 400, 250, k)[0]).max() > 1e-4   # it did move


def test_one_sided_solvers_and_anonymous_user():
    rng = np.random.default_rng(5)
    k = 16
    X, Y, xtx, yty = _model(rng, 300, 200, k)
    with FI.FoldIn(k, None, yty) as f:                     # model.solver.xtx.compute=false
        u, i = X[0].copy(), Y[0].copy()
        f.update_features(u, i, 1.0)
        uo, io = FO.update_features(X[0], Y[0], 1.0, None, yty)
        assert np.array_equal(i, Y[0]) and np.abs(u - uo).max() <= 1e-6 * np.abs(uo).max()
        with pytest.raises(FI.NotReadyException):
            f.solve(0, X[0])
        rows, vals = Y[[3, 17, 42]], np.array([1.0, -2.0, 5.0], np.float32)
        a = f.anonymous_user(rows, vals)
        assert np.abs(a - FO.anonymous_user(rows, vals, yty)).max() <= 1e-6 * np.abs(a).max()
        a1 = f.anonymous_user(rows)
        assert np.abs(a1 - FO.anonymous_user(rows, None, yty)).max() <= 1e-6 * np.abs(a1).max()
    with FI.FoldIn(k, xtx, None) as f:
        with pytest.raises(FI.NotReadyException):
            f.anonymous_user(Y[:2])


def test_ill_conditioned_and_singular_reports():
    k = 8
    rng = np.random.default_rng(9)
    tiny = np.eye(k) * 0.01                                 # infNorm < 1 (Generation.java:147-151)
    with pytest.raises(FI.IllConditionedSolverException):
        FI.FoldIn(k, tiny, None)
    with pytest.raises(FO.IllConditioned):
        FO.check_mtm(tiny)
    B = rng.standard_normal((200, 3)).astype(np.float32)    # rank 3 in 8 features
    M = np.hstack([B, B @ rng.standard_normal((3, 5)).astype(np.float32)]).astype(np.float32)
    G = O.transpose_times_self(M)
    with pytest.raises(FI.SingularMatrixSolverException) as e:
        FI.FoldIn(k, np.eye(k) * 2, G)
    assert e.value.which == 1
    with pytest.raises(O.SingularMatrixError) as eo:
        O.solve(G, np.ones(k))
    assert e.value.apparent_rank == eo.value.apparent_rank == 3


# ---- fold-in on the resident rows (csrc/foldin_dev.cuh) against the host library ------------
@pytest.mark.gpu
@pytest.mark.parametrize("k", [30, 64, 100])
def test_device_fold_in_matches_the_host_library(k):
    import myrrix_recommender_b200 as M
    from oracle import topn_oracle as T
    U, I = 1500, 900
    with M.NativeALS(k, device=0) as als:
        als.synth_interactions(U, I, 20, seed=11, neg_fraction=0.05)
        als.synth_y0(seed=11)
        als.iterate(3)
        als.sync()
        fi = als.recompute_state()                 # Gramians on the GPU, solvers on the host, copy on the device
        Xh, Yh = als.get_x(), als.get_y()
        X0 = Xh.copy()
        rng = np.random.default_rng(k)
        n = 400
        users = rng.integers(0, 40, n).astype(np.int32)      # few rows: every write sees earlier ones
        items = rng.integers(0, 25, n).astype(np.int32)
        values = rng.choice(np.array([1, 2, 5, -1, -3, 0.5], np.float32), n)
        fi.update_many(Xh, Yh, users, items, values)         # the host library, in order
        als.fold_in(users, items, values)                    # the device, in order, on the resident rows
        Xd, Yd = als.get_x(), als.get_y()
        for A, B in ((Xd, Xh), (Yd, Yh)):
            assert np.isfinite(A).all()
            # same factorisation, sums reassociated across lanes: identical except on rounding boundaries
            assert np.abs(A - B).max() <= 4e-7 * np.abs(B).max()
            assert (A == B).mean() > 0.995
        assert np.abs(Xd[:40] - X0[:40]).max() > 1e-4          # the writes did move the rows
        # a write moves the score it is about, and queries see the updated rows at once
        ptr, idx, _ = als.get_interactions()
        got_i, got_v = als.recommend([int(users[0])], 10)
        want_i, want_v = T.recommend(Yd, Xd, ptr, idx, [int(users[0])], 10)
        assert np.array_equal(got_i, want_i) and np.array_equal(got_v, want_v)
        # no write, no change; bad indices are rejected
        als.fold_in([], [], [])
        with pytest.raises(ValueError):
            als.fold_in([U], [0], [1.0])


@pytest.mark.gpu
def test_device_fold_in_one_sided_and_unset_state():
    import myrrix_recommender_b200 as M
    from myrrix_recommender_b200.factorizer import ExecutionException
    k = 16
    with M.NativeALS(k, device=0) as als:
        als.synth_interactions(300, 200, 10, seed=3)
        als.synth_y0(seed=3)
        als.iterate(2)
        als.sync()
        with pytest.raises(ExecutionException):
            als.fold_in([0], [0], [1.0])                     # no solver state yet
        fi = als.recompute_state(xtx=False)                  # model.solver.xtx.compute=false
        Xh, Yh = als.get_x(), als.get_y()
        fi.update_many(Xh, Yh, np.array([1, 2, 1], np.int32), np.array([5, 5, 6], np.int32), None)
        als.fold_in([1, 2, 1], [5, 5, 6])
        assert np.array_equal(als.get_y(), Yh)               # items untouched without the X'X solver
        assert np.abs(als.get_x() - Xh).max() <= 4e-7 * np.abs(Xh).max()
