"""Parity of the CUDA path against the CPU oracle and the reference's golden vectors.
All calls go through the C ABI (libmyrrix_als.so) via the host mirror.

Tolerance (BASELINE.json north_star): factor matrices within 1e-4 relative of the
reference arithmetic after a fixed iteration count, identical lambda/alpha/k/Y0:
  ||A-B||_F/||B||_F <= 1e-4  and  max|A-B|/max|B| <= 1e-4       
"""
import os

import numpy as np
import pytest

from conftest import dense_als_numpy, dense_to_csr, dense_to_maps, random_problem, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def M():
    import myrrix_recommender_b200 as m
    m._native.load()
    return m


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


KERNELS = [1, 0]  # ALS_KERNEL_SIMT, ALS_KERNEL_AUTO


def _run_gpu(M, ptr, idx, val, n_items, Y0, iters, kernel=0, **cfg):
    k = Y0.shape[1]
    with M.NativeALS(k, kernel=kernel, **cfg) as als:
        als.set_interactions(ptr.size - 1, n_items, ptr, idx, val)
        als.set_y(Y0)
        als.iterate(iters)
        als.sync()
        return als.get_x(), als.get_y(), als.info().kernel


def _mirror_product(M, g, reconstruct=False, kernel=0):
    by_row, by_col = dense_to_maps(g["R"])
    prevY = {i: np.array(v, np.float32) for i, v in enumerate(g["Y0"])}
    M.properties.clear()
    if reconstruct:
        M.properties["model.reconstructRMatrix"] = "true"
    try:
        als = M.AlternatingLeastSquares(by_row, by_col, g["features"], g["threshold"],
                                        g["max_iterations"], kernel=kernel)
        als.setPreviousY(prevY)
        als.call()
    finally:
        M.properties.clear()
    X, Y = als.getX(), als.getY()
    # MatrixUtils.multiplyXYT (MatrixUtils.java:153-163)
    P = np.array([[float(np.dot(X[r].astype(np.float64), Y[c].astype(np.float64)))
                   for c in range(len(Y))] for r in range(len(X))])
    return P, als.iterationsRun


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_als(M, goldens, kernel):
    """AlternatingLeastSquaresTest.testALS (:39-57) through the CUDA path. The reference
    asserts 1e-6 on its own fp64 arithmetic; the fp32 device path is held to the 1e-4 bar."""
    P, its = _mirror_product(M, goldens["als"], kernel=kernel)
    ref = np.array(goldens["als"]["product"])
    assert np.abs(P - ref).max() <= TOL * np.abs(ref).max()
    assert abs(its - 28) <= 1


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_als_reconstruct_r(M, goldens, kernel):
    """AlternatingLeastSquaresTest.testALSPredictingR (:60-78)."""
    P, its = _mirror_product(M, goldens["als"], reconstruct=True, kernel=kernel)
    ref = np.array(goldens["als"]["product_reconstruct_r"])
    assert np.abs(P - ref).max() <= TOL * np.abs(ref).max()
    assert abs(its - 34) <= 1


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_negative_input(M, goldens, kernel):
    """NegativeInputTest.testALS (:37-80)."""
    P, its = _mirror_product(M, goldens["negative_input"], kernel=kernel)
    ref = np.array(goldens["negative_input"]["product"])
    assert np.abs(P - ref).max() <= TOL * np.abs(ref).max()
    assert abs(its - 19) <= 1


def test_gramian_golden_and_parity(M, O, goldens):
    """MatrixUtilsTest.testTransposeTimesSelf (:63-73) + random parity (fp64 accumulate)."""
    g = goldens["transpose_times_self"]
    Mx = np.array(g["M"], np.float32)
    with M.NativeALS(3) as als:
        ptr = np.array([0, 1, 2], np.int64)
        als.set_interactions(2, 2, ptr, np.array([0, 1], np.int32), np.ones(2, np.float32))
        als.set_y(Mx)
        assert np.abs(als.gramian("y") - np.array(g["MTM"])).max() <= 1e-12
    rng = np.random.default_rng(5)
    for k, n in ((16, 1000), (30, 777), (64, 5000), (128, 300)):
        Y = rng.standard_normal((n, k)).astype(np.float32)
        with M.NativeALS(k) as als:
            als.set_interactions(1, n, np.array([0, 1], np.int64), np.array([0], np.int32),
                                 np.ones(1, np.float32))
            als.set_y(Y)
            G = als.gramian("y")
        Go = O.transpose_times_self(Y)
        assert np.abs(G - Go).max() <= 1e-6 * np.abs(Go).max()
        assert np.array_equal(G, G.T)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("k,n_users,n_items,nnz", [
    (2, 50, 40, 5), (3, 64, 33, 7), (16, 2000, 400, 20), (30, 500, 300, 25),
    (32, 1500, 500, 50), (50, 800, 500, 60), (64, 1200, 600, 70), (100, 300, 400, 40),
    (128, 400, 500, 90),
])
def test_factor_parity_fixed_iterations(M, O, kernel, k, n_users, n_items, nnz):
    """5 iterations from the same Y0; X and Y within 1e-4 relative of the oracle; 5% negative
    strengths (the r>0 gate, ALS.java:480-482), a few empty users and stale items (8b)."""
    ptr, idx, val, Y0 = random_problem(n_users, n_items, nnz, k, seed=100 + k, neg_fraction=0.05,
                                       empty_users=3, stale_items=2)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=5,
                             convergence_threshold=1e-12, n_threads=8)
    X, Y, used = _run_gpu(M, ptr, idx, val, n_items, Y0, 5, kernel=kernel)
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (k, used, fro, mx)
    assert np.all(X[:3] == 0)                      # users without entries: not in the map
    assert np.array_equal(Y[-2:], Y0[-2:])         # stale item rows pass through unchanged


@pytest.mark.parametrize("kernel", KERNELS)
def test_config1_parity(M, O, kernel):
    """BASELINE.json configs[0]: 10k x 2k, 20 nnz/user, k=16, device-generated workload."""
    from oracle import synth
    with M.NativeALS(16, kernel=kernel) as als:
        als.synth_interactions(10000, 2000, 20, seed=1234567890, neg_fraction=0.05)
        als.synth_y0(seed=1234567890)
        ptr, idx, val = als.get_interactions()
        cptr, cidx, cval = als.get_interactions(by_column=True)
        Y0 = als.get_y()
        als.iterate(5)
        als.sync()
        X, Y = als.get_x(), als.get_y()
    # the device generator is bit-identical to its numpy twin, the transpose to a stable sort
    p2, i2, v2 = synth.synth_rows(0, 10000, 2000, 20, seed=1234567890, neg_fraction=0.05)
    assert np.array_equal(ptr, p2) and np.array_equal(idx, i2) and np.array_equal(val, v2)
    tp, ti, tv = O.csr_transpose(ptr, idx, val, 2000)
    assert np.array_equal(cptr, tp) and np.array_equal(cidx, ti) and np.array_equal(cval, tv)
    assert np.abs(np.linalg.norm(Y0.astype(np.float64), axis=1) - 1).max() < 1e-6
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, 2000, Y0, max_iterations=5,
                             convergence_threshold=1e-12, n_threads=8, t_csr=(tp, ti, tv))
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (fro, mx)


@pytest.mark.parametrize("kernel", KERNELS)
def test_variants_alpha_lambda_reconstruct_ignore(M, O, kernel):
    ptr, idx, val, Y0 = random_problem(400, 150, 12, 8, seed=7, neg_fraction=0.2)
    cases = [dict(alpha=40.0, lam=0.01), dict(reconstruct_r=True),
             dict(loss_ignores_unspecified=True), dict(alpha=0.5, lam=2.0, reconstruct_r=True,
                                                       loss_ignores_unspecified=True)]
    for cfg in cases:
        Xo, Yo, _, _ = O.als_run(ptr, idx, val, 150, Y0, max_iterations=3,
                                 convergence_threshold=1e-12, **cfg)
        X, Y, _ = _run_gpu(M, ptr, idx, val, 150, Y0, 3, kernel=kernel, **cfg)
        for a, b in ((X, Xo), (Y, Yo)):
            fro, mx = rel_err(a, b)
            assert fro <= TOL and mx <= TOL, (cfg, fro, mx)


@pytest.mark.parametrize("kernel", KERNELS)
def test_ragged_rows_one_to_thousands(M, O, kernel):
    """Row lengths from 1 to 3000 entries (power-law-like), k=32."""
    rng = np.random.default_rng(11)
    n_items, k = 4000, 32
    lens = np.concatenate([[1, 2, 3, 31, 32, 33, 63, 64, 65, 3000],
                           np.minimum(3000, (rng.pareto(1.2, 300) * 5 + 1).astype(int))])
    ptr, idx, val = [0], [], []
    for n in lens:
        idx += list(np.sort(rng.choice(n_items, size=n, replace=False)))
        val += list(rng.integers(1, 6, size=n).astype(np.float32))
        ptr.append(len(idx))
    ptr, idx, val = np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)
    d = rng.standard_normal((n_items, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=2,
                             convergence_threshold=1e-12, n_threads=8)
    X, Y, _ = _run_gpu(M, ptr, idx, val, n_items, Y0, 2, kernel=kernel)
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (fro, mx)


@pytest.mark.parametrize("mix", ["4", "8"])
@pytest.mark.parametrize("k", [32, 64])
def test_warp_role_mixes(M, O, monkeypatch, mix, k):
    """Both role mixes of the tensor-core kernel (8 Cholesky + 7 producer warps for short rows,
    4 + 11 for long rows; chosen by row length in production, forced here) on the same
    problem: short user rows, long item rows, a few rows past one accumulation segment."""
    monkeypatch.setenv("MYRRIX_ALS_MIX", mix)
    rng = np.random.default_rng(23 + k)
    n_users, n_items = 1500, 100  # item rows: ~1200 entries (> the 1024-entry segment)
    lens = np.concatenate([[1, 16, 17, 100], rng.integers(60, 100, size=n_users - 4)])
    ptr, idx, val = [0], [], []
    for n in lens:
        idx += list(np.sort(rng.choice(n_items, size=n, replace=False)))
        v = rng.integers(1, 6, size=n).astype(np.float32)
        v[rng.random(n) < 0.05] *= -1
        val += list(v)
        ptr.append(len(idx))
    ptr, idx, val = np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)
    d = rng.standard_normal((n_items, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=3,
                             convergence_threshold=1e-12, n_threads=8)
    X, Y, used = _run_gpu(M, ptr, idx, val, n_items, Y0, 3, kernel=2)
    assert used == 2
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (mix, k, fro, mx)


@pytest.mark.parametrize("limit", ["64", "200", "1000"])
@pytest.mark.parametrize("k", [32, 64, 50])
def test_long_rows_split_into_chunks(M, O, monkeypatch, limit, k):
    """Rows of more than MYRRIX_ALS_SPLIT_ROWS entries are walked as several virtual rows by the
    tensor-core kernel (partial -D and rhs summed in a per-row record, the last chunk to arrive
    assembles W_u and solves): the result is the one of the whole row.  The limit is forced low here
    so that ragged rows of 65..3000 entries split into 2..47 chunks, next to whole rows, in both halves."""
    monkeypatch.setenv("MYRRIX_ALS_SPLIT_ROWS", limit)
    rng = np.random.default_rng(5 + k)
    n_items = 900
    lens = np.concatenate([[1, 2, 63, 64, 65, 97, 128, 129, 199, 200, 201, 257, 900, 3000 % 900 + 300],
                           np.minimum(900, (rng.pareto(1.1, 700) * 8 + 1).astype(int))])
    ptr, idx, val = [0], [], []
    for n in lens:
        idx += list(np.sort(rng.choice(n_items, size=n, replace=False)))
        v = rng.integers(1, 6, size=n).astype(np.float32)
        v[rng.random(n) < 0.05] *= -1
        val += list(v)
        ptr.append(len(idx))
    ptr, idx, val = np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)
    d = rng.standard_normal((n_items, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=3,
                             convergence_threshold=1e-12, n_threads=8)
    X, Y, used = _run_gpu(M, ptr, idx, val, n_items, Y0, 3, kernel=2)
    assert used == 2
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (limit, k, fro, mx)


def _ill_conditioned_problem(k):
    """Ragged user rows against item vectors with a strong common component: for ~85 % of the user rows
    max diagonal / min pivot of W_u is 300..1000 (numpy), past the fp32 solve's conditioning gate of 256."""
    rng = np.random.default_rng(31 + k)
    n_users, n_items = 2500, 1200
    lens = np.concatenate([[1, 2, 17, 129, 257, 600],
                           np.minimum(600, (rng.pareto(1.3, n_users - 6) * 10 + 2).astype(int))])
    ptr, idx, val = [0], [], []
    for n in lens:
        idx += list(np.sort(rng.choice(n_items, size=n, replace=False)))
        v = rng.integers(1, 6, size=n).astype(np.float32)
        v[rng.random(n) < 0.05] *= -1
        val += list(v)
        ptr.append(len(idx))
    ptr, idx, val = np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)
    d = 0.03 * rng.standard_normal((n_items, k))
    d[:, 0] += 1.0   # (along one axis, so that the largest DIAGONAL entry -- what the gate looks at -- is ~n_items)
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    return n_users, n_items, ptr, idx, val, Y0


@pytest.mark.parametrize("k", [64, 32, 50])
def test_stash_mode_resolves_rows_against_an_ill_conditioned_gramian(M, O, monkeypatch, k):
    """Stash mode forced from the first launch: the tensor-core kernel hands the data term of the rows its
    fp32 solve refuses to the fp64 re-solve (fp64 Gramian + tensor-core D, no second gather) instead of
    the CUDA-core fp64 kernel.  Rows of more than 128 entries are split here, so rows assembled from chunks
    take the path too.  Three iterations against the oracle."""
    monkeypatch.setenv("MYRRIX_ALS_STASH", "1")
    monkeypatch.setenv("MYRRIX_ALS_SPLIT_ROWS", "128")
    n_users, n_items, ptr, idx, val, Y0 = _ill_conditioned_problem(k)
    iters = 3
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, n_items, Y0, max_iterations=iters,
                             convergence_threshold=1e-12, n_threads=8)
    with M.NativeALS(k, kernel=2) as als:
        als.set_interactions(n_users, n_items, ptr, idx, val)
        als.set_y(Y0)
        als.timings(reset=True)
        als.iterate(iters)
        als.sync()
        X, Y = als.get_x(), als.get_y()
        tm = als.timings()
    assert tm.fp64_resolve_rows > n_users // 2, tm.fp64_resolve_rows   # the gate did refuse them
    assert tm.fp64_retry_rows == tm.fp64_resolve_rows                   # and none was gathered again
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (k, fro, mx)


def test_stash_mode_switches_on_after_a_launch_that_refused_many_rows(M, O, monkeypatch):
    """Default policy: the first X-half sends the refused rows to the CUDA-core fp64 kernel and reports how
    many there were; the next X-half (same Y, so the same rows fail) runs in stash mode.  Same result."""
    monkeypatch.delenv("MYRRIX_ALS_STASH", raising=False)
    k = 64
    n_users, n_items, ptr, idx, val, Y0 = _ill_conditioned_problem(k)
    out = np.zeros((n_users, k), np.float32)
    O.als_half(ptr, idx, val, Y0, O.transpose_times_self(Y0), out, n_threads=8)
    with M.NativeALS(k, kernel=2) as als:
        als.set_interactions(n_users, n_items, ptr, idx, val)
        als.set_y(Y0)
        als.timings(reset=True)
        als.half_x(); als.sync()
        X1 = als.get_x()
        t1 = als.timings(reset=True)
        als.half_x(); als.sync()
        X2 = als.get_x()
        t2 = als.timings(reset=True)
    assert t1.fp64_retry_rows > n_users // 2 and t1.fp64_resolve_rows == 0
    assert t2.fp64_resolve_rows > n_users // 2 and t2.fp64_retry_rows == t2.fp64_resolve_rows
    for X in (X1, X2):
        fro, mx = rel_err(X, out)
        assert fro <= TOL and mx <= TOL, (fro, mx)


def test_csv_files_to_factors(M, O):
    """The step in front of the path joined to it: CSV lines with duplicates,
    deletions, near-zero sums and comments -> libmyrrix_ingest.so -> als_set_interactions ->
    3 iterations, against the ALS oracle run on the same canonical matrix (the canonicalisation
    itself is held to its own line-by-line oracle in tests/test_ingest.py)."""
    from myrrix_recommender_b200 import ingest as ING
    rng = np.random.default_rng(77)
    lines = ["user,item,strength"]
    for _ in range(30000):
        u, i = int(rng.integers(1000, 1400)), int(rng.integers(-40, 60))
        r = rng.random()
        if r < 0.05:
            lines.append("%d,%d," % (u, i))
        elif r < 0.15:
            lines.append("%d,%d" % (u, i))
        elif r < 0.17:
            lines.append("# %d" % u)
        else:
            lines.append("%d,%d,%.3f" % (u, i, rng.integers(1, 6) * (1 if rng.random() > 0.05 else -1)))
    res = ING.read_csv_bytes(("\n".join(lines) + "\n").encode())
    n_users, n_items, k = len(res.user_ids), len(res.item_ids), 16
    assert n_users > 300 and n_items > 90 and res.col_idx.size > 10000
    d = rng.standard_normal((n_items, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    Xo, Yo, _, _ = O.als_run(res.row_ptr, res.col_idx, res.val, n_items, Y0, max_iterations=3,
                             convergence_threshold=1e-12, n_threads=8)
    X, Y, _ = _run_gpu(M, res.row_ptr, res.col_idx, res.val, n_items, Y0, 3)
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (fro, mx)


@pytest.mark.parametrize("kernel", KERNELS)
def test_singular_reports_rank_like_reference(M, O, kernel):
    """lambda=0 and fewer independent rows than features: the reference throws
    SingularMatrixSolverException(apparentRank) (CommonsMathLinearSystemSolver.java:43-54);
    the device path must return ALS_E_SINGULAR, never NaN/Inf."""
    k = 6
    ptr, idx, val, _ = random_problem(20, 3, 2, k, seed=9, neg_fraction=0.0)
    rng = np.random.default_rng(9)
    Y0 = rng.standard_normal((3, k)).astype(np.float32)
    with pytest.raises(O.SingularMatrixError) as eo:
        O.als_run(ptr, idx, val, 3, Y0, lam=0.0, max_iterations=1)
    with M.NativeALS(k, lam=0.0, kernel=kernel) as als:
        als.set_interactions(20, 3, ptr, idx, val)
        als.set_y(Y0)
        als.half_x()
        with pytest.raises(M.SingularMatrixSolverException) as eg:
            als.sync()
        assert eg.value.getApparentRank() == eo.value.apparent_rank == 3
        with pytest.raises(M.SingularMatrixSolverException):  # sticky until destroyed
            als.sync()


def test_probe_matches_simple_vector_math_dot(M, O):
    ptr, idx, val, Y0 = random_problem(100, 60, 6, 30, seed=13)
    with M.NativeALS(30) as als:
        als.set_interactions(100, 60, ptr, idx, val)
        als.set_y(Y0)
        als.iterate(1)
        users, items = np.arange(0, 100, 7), np.arange(0, 60, 5)
        P = als.probe(users, items)
        X, Y = als.get_x(), als.get_y()
    ref = np.array([[O.dot(X[u], Y[i]) for i in items] for u in users])
    assert np.array_equal(P, ref)  # fp32-rounded products, fp64 sum in index order: bit-exact


def test_mirror_stale_and_warm_start(M, O):
    """setPreviousY with an extra stale item row and a present-but-empty user."""
    by_row, by_col = dense_to_maps([[0, 2, 3, 1], [4, 0, 0, 5], [1, 1, 0, 0]])
    by_row[99] = {}
    rng = np.random.default_rng(3)
    prev = {i: rng.standard_normal(3).astype(np.float32) for i in (0, 1, 2, 3, 42)}
    prev_copy = {i: v.copy() for i, v in prev.items()}
    als = M.AlternatingLeastSquares(by_row, by_col, 3, 1e-9, 4)
    als.setPreviousY(prev)
    als.call()
    assert als.getY() is prev and set(als.getX()) == {0, 1, 2, 99}
    assert np.array_equal(als.getY()[42], prev_copy[42])
    assert np.all(als.getX()[99] == 0)
    # oracle on the same dense problem (user 99 -> row 3 with no entries, item 42 -> row 4)
    ptr, idx, val = dense_to_csr([[0, 2, 3, 1, 0], [4, 0, 0, 5, 0], [1, 1, 0, 0, 0], [0, 0, 0, 0, 0]])
    Y0 = np.stack([prev_copy[i] for i in (0, 1, 2, 3, 42)])
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, 5, Y0, max_iterations=4, convergence_threshold=1e-9)
    Xg = np.stack([als.getX()[u] for u in (0, 1, 2)])
    Yg = np.stack([als.getY()[i] for i in (0, 1, 2, 3, 42)])
    assert rel_err(Xg, Xo[:3])[0] <= TOL and rel_err(Yg, Yo)[0] <= TOL


def test_large_config_sampled_row_parity(M, O):
    """Size-independent property at a BASELINE-scale shape (1M x 100k, 50/user, k=32 = configs[1]):
    after 2 device iterations, re-solve 200 sampled user rows and 50 item rows on the CPU oracle
    from the device's own opposite factor; each must match to 1e-4."""
    k, U, I, nnz = 32, 1000000, 100000, 50
    with M.NativeALS(k) as als:
        als.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
        als.synth_y0(seed=1234567890)
        als.iterate(1)
        als.half_x()
        als.sync()
        X, Y = als.get_x(), als.get_y()
        ptr, idx, val = als.get_interactions()
        als.half_y()
        als.sync()
        Y2 = als.get_y()
        cptr, cidx, cval = als.get_interactions(by_column=True)
    assert np.isfinite(X).all() and np.isfinite(Y2).all()
    rng = np.random.default_rng(0)
    G = O.transpose_times_self(Y)
    rows = np.sort(rng.choice(U, 200, replace=False))
    sp = np.concatenate([[0], np.cumsum(ptr[rows + 1] - ptr[rows])]).astype(np.int64)
    si = np.concatenate([idx[ptr[r]:ptr[r + 1]] for r in rows])
    sv = np.concatenate([val[ptr[r]:ptr[r + 1]] for r in rows])
    out = np.zeros((200, k), np.float32)
    O.als_half(sp, si, sv, Y, G, out)
    fro, mx = rel_err(X[rows], out)
    assert fro <= TOL and mx <= TOL, (fro, mx)
    G = O.transpose_times_self(X)
    rows = np.sort(rng.choice(I, 50, replace=False))
    sp = np.concatenate([[0], np.cumsum(cptr[rows + 1] - cptr[rows])]).astype(np.int64)
    si = np.concatenate([cidx[cptr[r]:cptr[r + 1]] for r in rows])
    sv = np.concatenate([cval[cptr[r]:cptr[r + 1]] for r in rows])
    out = np.zeros((50, k), np.float32)
    O.als_half(sp, si, sv, X, G, out)
    fro, mx = rel_err(Y2[rows], out)
    assert fro <= TOL and mx <= TOL, (fro, mx)


def test_sharded_synthesis_matches_slices_of_the_whole(M, O):
    """Partition-only handles (no communicator) on one GPU: every rank's by-user and by-item
    shard equals the matching slice of the single-GPU matrix, and a sharded half-iteration
    writes exactly its own block."""
    from myrrix_recommender_b200 import sharding as S
    U, I, nnz, k, world = 5003, 701, 23, 32, 3
    with M.NativeALS(k) as whole:
        whole.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
        whole.synth_y0(seed=1234567890)
        ptr, idx, val = whole.get_interactions()
        cptr, cidx, cval = whole.get_interactions(by_column=True)
        Y0 = whole.get_y()
        whole.half_x(); whole.sync()
        X_full = whole.get_x()
        whole.half_y(); whole.sync()
        Y_full = whole.get_y()
    for rank in range(world):
        with M.NativeALS(k) as als:
            als.comm_init(rank, world, None)
            als.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
            als.synth_y0(seed=1234567890)
            p, i, v = als.get_interactions()
            ep, ei, ev = S.shard_rows(ptr, idx, val, rank, world)
            assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(v, ev)
            p, i, v = als.get_interactions(by_column=True)
            ep, ei, ev = S.shard_rows(cptr, cidx, cval, rank, world)
            assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(v, ev)
            als.half_x(); als.sync()
            X = als.get_x()
            ub, ue = S.local_block(U, rank, world)
            assert np.array_equal(X[ub:ue], X_full[ub:ue])
            assert not X[:ub].any() and not X[ue:].any()
            als.set_x(X_full)  # what the all-gather would deliver
            als.half_y(); als.sync()
            Y = als.get_y()
            ib, ie = S.local_block(I, rank, world)
            # same arithmetic, but the rhs partial sums are grouped by the CTA's stage stream,
            # which depends on the row set: equal to fp32 rounding, not bit for bit
            assert np.allclose(Y[ib:ie], Y_full[ib:ie], rtol=1e-5, atol=1e-6)
            assert np.array_equal(Y[:ib], Y0[:ib]) and np.array_equal(Y[ie:], Y0[ie:])


def test_sharded_powerlaw_synthesis_matches_slices_of_the_twin(M, O):
    """The power-law generator on sharded handles (partition-only here: one GPU): every rank's by-user
    block and by-item block equal the slices of oracle/synth.py's matrix."""
    from myrrix_recommender_b200 import sharding as S
    from oracle import synth
    U, I, k, world = 3001, 457, 32, 3
    ptr, idx, val = synth.synth_rows_powerlaw(0, U, I, 25, max_nnz=400, seed=99, neg_fraction=0.05)
    with M.NativeALS(k) as whole:
        whole.set_interactions(U, I, ptr, idx, val)
        cptr, cidx, cval = whole.get_interactions(by_column=True)
    for rank in range(world):
        with M.NativeALS(k) as als:
            als.comm_init(rank, world, None)
            als.synth_interactions_powerlaw(U, I, 25, max_nnz=400, seed=99, neg_fraction=0.05)
            p, i, v = als.get_interactions()
            ep, ei, ev = S.shard_rows(ptr, idx, val, rank, world)
            assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(v, ev)
            p, i, v = als.get_interactions(by_column=True)
            ep, ei, ev = S.shard_rows(cptr, cidx, cval, rank, world)
            assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(v, ev)


def test_build_then_fold_in(M, O):
    """The step after the path: a model build on the GPU, X'X and Y'Y from the
    resident factors (als_gramian), then online writes folded in by libmyrrix_foldin.so --
    against the fold-in restatement fed with the oracle's Gramians of the same factors."""
    from myrrix_recommender_b200 import foldin as FI
    from oracle import foldin_oracle as FO
    k = 16
    ptr, idx, val, Y0 = random_problem(600, 200, 12, k, seed=31, neg_fraction=0.05)
    with M.NativeALS(k) as als:
        als.set_interactions(ptr.size - 1, 200, ptr, idx, val)
        als.set_y(Y0)
        als.iterate(3)
        als.sync()
        X, Y = als.get_x(), als.get_y()
        GX, GY = als.gramian("x"), als.gramian("y")
    GXo, GYo = O.transpose_times_self(X), O.transpose_times_self(Y)
    for a, b in ((GX, GXo), (GY, GYo)):
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    rng = np.random.default_rng(32)
    Xo, Yo = X.copy(), Y.copy()
    with FI.FoldIn(k, GX, GY) as f:
        for _ in range(100):
            u, i = int(rng.integers(len(X))), int(rng.integers(len(Y)))
            v = float(rng.choice([-1.0, 1.0, 2.0, 5.0]))
            f.update_features(X[u], Y[i], v)
            Xo[u], Yo[i] = FO.update_features(Xo[u], Yo[i], v, GXo, GYo)
    for a, b in ((X, Xo), (Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (fro, mx)


# ---------------------------------------------------------------------------------------------
# round 2: the blocked tensor-core solve in isolation, present-but-empty rows, input validation,
# and parity at the headline scale


def _spd(rng, k, cond):
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    ev = np.logspace(0, np.log10(cond), k)
    return (q * ev) @ q.T


@pytest.mark.parametrize("k", [64, 50, 33, 32, 17, 5])
def test_blocked_solver_unit(M, k):
    """chol_blocked.cuh alone (one warp, one dense system) against numpy's fp64 solve: panel
    factorisation, 3xTF32 mma.sync trailing updates, backward substitution, padding rows."""
    import ctypes as C
    lib = M._native.load()
    fn = lib.als_debug_solve_blocked
    fn.restype = C.c_int
    fp = C.POINTER(C.c_float)
    fn.argtypes = [fp, fp, C.c_int, C.c_float, fp, C.POINTER(C.c_int)]
    rng = np.random.default_rng(100 + k)
    for cond in (3.0, 50.0):
        W = _spd(rng, k, cond) * 40.0
        b = rng.standard_normal(k) * 7.0
        W32 = np.ascontiguousarray(W, dtype=np.float32)
        b32 = np.ascontiguousarray(b, dtype=np.float32)
        x = np.zeros(k, np.float32)
        ok = C.c_int(0)
        assert fn(W32.ctypes.data_as(fp), b32.ctypes.data_as(fp), k, 1e-5, x.ctypes.data_as(fp), C.byref(ok)) == 0
        ref = np.linalg.solve(W32.astype(np.float64), b32.astype(np.float64))
        fro, mx = rel_err(x, ref)
        assert ok.value == 1 and fro <= 2e-5 and mx <= 2e-5, (k, cond, ok.value, fro, mx)
    # ill-conditioned / indefinite systems are refused (the caller re-solves them in fp64)
    for W in (_spd(rng, k, 1e6), -np.eye(k)):
        W32 = np.ascontiguousarray(W, dtype=np.float32)
        ok = C.c_int(1)
        x = np.zeros(k, np.float32)
        b32 = np.ones(k, np.float32)
        assert fn(W32.ctypes.data_as(fp), b32.ctypes.data_as(fp), k, 1e-5, x.ctypes.data_as(fp), C.byref(ok)) == 0
        assert ok.value == 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_present_but_empty_rows_solve_to_zero_like_the_reference(M, kernel):
    """A key of RbyRow / RbyColumn whose map removeSmall emptied (InputFilesReader.java:202-211)
    is still walked by addWorkers (ALS.java:391-410): W = G, b = 0 -> the zero vector after the
    FIRST half, so it stops feeding X^T X / Y^T Y.  Checked against a dense numpy solve that
    shares nothing with the oracle or the kernels."""
    rng = np.random.default_rng(9)
    U, I, k = 40, 30, 6
    R = np.zeros((U, I), np.float32)
    for u in range(U):
        cols = rng.choice(I - 2, 7, replace=False)       # items I-2, I-1 never occur
        R[u, cols] = rng.integers(1, 6, 7)
    R[5] = 0      # user 5: key with an emptied map
    R[6] = 0      # user 6: likewise
    by_row, by_col = dense_to_maps(R)
    by_row[5] = {}
    by_row[6] = {}
    by_col[I - 2] = {}                                  # item I-2: present but empty
    d = rng.standard_normal((I, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    prev = {i: Y0[i].copy() for i in range(I)}          # item I-1: stale (not a key of RbyColumn)
    als = M.AlternatingLeastSquares(by_row, by_col, k, 1e-9, 3, kernel=kernel)
    als.setPreviousY(prev)
    als.call()
    X, Y = als.getX(), als.getY()
    Xn, Yn = dense_als_numpy(R, {5, 6}, {I - 2}, Y0, 3)
    assert np.all(X[5] == 0) and np.all(X[6] == 0) and np.all(Y[I - 2] == 0)
    assert np.array_equal(Y[I - 1], Y0[I - 1])          # stale row untouched
    Xg = np.stack([X[u] for u in range(U)])
    Yg = np.stack([Y[i] for i in range(I)])
    for a, b in ((Xg, Xn), (Yg, Yn)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (fro, mx)


def test_upload_validation_rejects_malformed_csr(M):
    """als_set_interactions* must refuse a non-monotone row_ptr or an out-of-range index with
    ALS_E_ARG instead of launching gathers on it (raw C-ABI calls: the Python wrapper's own
    checks are bypassed on purpose)."""
    import ctypes as C
    N = M._native
    with M.NativeALS(8) as als:
        def call(ptr, idx, val, U=3, I=4):
            ptr = np.asarray(ptr, np.int64); idx = np.asarray(idx, np.int32); val = np.asarray(val, np.float32)
            return als.lib.als_set_interactions(als.h, U, I, ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                                val.ctypes.data_as(C.POINTER(C.c_float)))
        assert call([0, 2, 1, 3], [0, 1, 2], [1, 1, 1]) == N.ALS_E_ARG      # not monotone
        assert call([0, 1, 2, 3], [0, 4, 2], [1, 1, 1]) == N.ALS_E_ARG      # index == n_items
        assert call([0, 1, 2, 3], [0, -1, 2], [1, 1, 1]) == N.ALS_E_ARG     # negative index
        assert call([0, 1, 2, 3], [0, 3, 2], [1, 1, 1]) == N.ALS_OK
        # the handle is still usable after the rejections
        als.n_users, als.n_items = 3, 4
        als.set_y(np.eye(4, 8, dtype=np.float32))
        als.iterate(1)
        als.sync()
        assert np.isfinite(als.get_x()).all()


def _sampled_rows(als, rows, by_column):
    ptrs, idxs, vals = [0], [], []
    for r in rows:
        p, i, v = als.get_interaction_rows(int(r), 1, by_column=by_column, capacity=1 << 16)
        idxs.append(i); vals.append(v); ptrs.append(ptrs[-1] + i.size)
    return np.array(ptrs, np.int64), np.concatenate(idxs), np.concatenate(vals)


def test_headline_config_sampled_row_parity(M, O):
    """BASELINE.json configs[2] at full size (10M x 1M, 100 entries/user, k = 64): after 1.5
    device iterations re-solve 200 sampled user rows and 50 item rows (1000 entries each: past
    the 1024-entry accumulation segment for some) on the CPU oracle from the device's own
    opposite factor and Gramian inputs; each must match to 1e-4."""
    k, U, I, nnz = 64, 10_000_000, 1_000_000, 100
    rng = np.random.default_rng(0)
    with M.NativeALS(k) as als:
        als.synth_interactions(U, I, nnz, seed=1234567890, neg_fraction=0.05)
        als.synth_y0(seed=1234567890)
        als.iterate(1)
        als.half_x()
        als.sync()
        assert als.info().kernel == 2
        urows = np.sort(rng.choice(U, 200, replace=False))
        irows = np.sort(rng.choice(I, 50, replace=False))
        up, ui, uv = _sampled_rows(als, urows, False)
        ip, ii, iv = _sampled_rows(als, irows, True)
        Y = als.get_y()
        X = als.get_x()
        als.half_y()
        als.sync()
        Y2 = als.get_y()
        retried = als.timings().fp64_retry_rows
    assert np.isfinite(X).all() and np.isfinite(Y2).all()
    out = np.zeros((urows.size, k), np.float32)
    O.als_half(up, ui, uv, Y, O.transpose_times_self(Y), out)
    fro, mx = rel_err(X[urows], out)
    assert fro <= TOL and mx <= TOL, ("X", fro, mx)
    out = np.zeros((irows.size, k), np.float32)
    O.als_half(ip, ii, iv, X, O.transpose_times_self(X), out, n_threads=8)
    fro, mx = rel_err(Y2[irows], out)
    assert fro <= TOL and mx <= TOL, ("Y", fro, mx)
    assert retried < 1000  # the fp32 fast path carries the workload


def test_headline_cutdown_full_parity_five_iterations(M, O):
    """100k x 20k, 50 entries/user, k = 64 (the cut-down of the headline config),
    5 % negative strengths, 5 fixed iterations, every row of X and Y against the oracle."""
    import os
    from oracle import synth
    k, U, I, nnz = 64, 100_000, 20_000, 50
    ptr, idx, val = synth.synth_rows(0, U, I, nnz, seed=1234567890, neg_fraction=0.05)
    Y0 = synth.unit_rows(I, k, seed=1234567890)
    Xo, Yo, its, _ = O.als_run(ptr, idx, val, I, Y0, max_iterations=5, convergence_threshold=1e-12,
                               n_threads=os.cpu_count() or 8)
    assert its == 5
    X, Y, used = _run_gpu(M, ptr, idx, val, I, Y0, 5)
    assert used == 2
    for name, a, b in (("X", X, Xo), ("Y", Y, Yo)):
        fro, mx = rel_err(a, b)
        assert fro <= TOL and mx <= TOL, (name, fro, mx)


@pytest.mark.gpu
def test_device_stop_rule_equals_host_stop_rule(M, goldens):
    """als_call (stop rule evaluated on the device, ALS.java:230-257) stops after the same
    iteration with the same DoubleWeightedMean as the host loop over als_probe."""
    for name, recon in (("als", False), ("als", True), ("negative_input", False)):
        g = goldens[name]
        runs = []
        for host_rule in (False, True):
            by_row, by_col = dense_to_maps(g["R"])
            prevY = {i: np.array(v, np.float32) for i, v in enumerate(g["Y0"])}
            M.properties.clear()
            if recon:
                M.properties["model.reconstructRMatrix"] = "true"
            if host_rule:
                M.properties["model.als.hostStopRule"] = "true"
            try:
                als = M.AlternatingLeastSquares(by_row, by_col, g["features"], g["threshold"],
                                                g["max_iterations"])
                als.setPreviousY(prevY)
                als.call()
            finally:
                M.properties.clear()
            runs.append((als.iterationsRun, als.lastConvergenceValue,
                         np.stack([als.getX()[u] for u in sorted(als.getX())])))
        assert runs[0][0] == runs[1][0]
        assert runs[0][1] == runs[1][1]  # bit-identical statistic
        assert np.array_equal(runs[0][2], runs[1][2])


@pytest.mark.gpu
def test_powerlaw_generator_matches_its_twin_and_factors_match_the_oracle(M, O):
    """Config-5-style data: the device generator against oracle/synth.py bit for bit, then
    two iterations on it (rows from a handful to thousands of entries, a few items that nearly every
    user has) against the oracle."""
    from oracle import synth
    U, I, k = 6000, 2500, 64
    with M.NativeALS(k) as als:
        als.synth_interactions_powerlaw(U, I, 30, max_nnz=2000, seed=77, neg_fraction=0.05)
        ptr, idx, val = als.get_interactions()
        tp, ti, tv = synth.synth_rows_powerlaw(0, U, I, 30, max_nnz=2000, seed=77, neg_fraction=0.05)
        assert np.array_equal(ptr, tp) and np.array_equal(idx, ti) and np.array_equal(val, tv)
        cp, ci, cv = als.get_interactions(by_column=True)
        assert cp[-1] == ptr[-1] and np.diff(cp).max() > 0.5 * U   # the most popular item: most users
        als.synth_y0(seed=77)
        Y0 = als.get_y()
        als.iterate(2)
        als.sync()
        X, Y = als.get_x(), als.get_y()
    Xo, Yo, _, _ = O.als_run(ptr, idx, val, I, Y0, max_iterations=2, convergence_threshold=1e-12,
                             n_threads=os.cpu_count() or 1)
    for A, B in ((X, Xo), (Y, Yo)):
        assert np.linalg.norm(A - B) <= TOL * np.linalg.norm(B)
        assert np.abs(A - B).max() <= TOL * np.abs(B).max()
