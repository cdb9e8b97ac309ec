"""Pins the CPU oracle (oracle/als_oracle.c) against every golden vector the reference's own
tests hold for the ALS path (tests/golden/make_golden.py lists file:line for each)."""
import numpy as np
import pytest

from conftest import dense_to_csr
from oracle import oracle as O


def _product(X, Y):
    # MatrixUtils.multiplyXYT (MatrixUtils.java:153-163) -> SimpleVectorMath.dot
    return np.array([[O.dot(x, y) for y in Y] for x in X])


def test_als_golden(goldens):
    """AlternatingLeastSquaresTest.testALS (:39-57): 25 entries of X*Y^T to 1e-6."""
    g = goldens["als"]
    ptr, idx, val = dense_to_csr(g["R"])
    X, Y, its, conv = O.als_run(ptr, idx, val, 5, np.array(g["Y0"], np.float32),
                                convergence_threshold=g["threshold"],
                                max_iterations=g["max_iterations"])
    assert np.abs(_product(X, Y) - np.array(g["product"])).max() < goldens["float_epsilon"]
    assert its == 28 and conv < g["threshold"]


def test_als_golden_reconstruct_r(goldens):
    """AlternatingLeastSquaresTest.testALSPredictingR (:60-78), model.reconstructRMatrix=true."""
    g = goldens["als"]
    ptr, idx, val = dense_to_csr(g["R"])
    X, Y, its, _ = O.als_run(ptr, idx, val, 5, np.array(g["Y0"], np.float32),
                             convergence_threshold=g["threshold"],
                             max_iterations=g["max_iterations"], reconstruct_r=True)
    assert np.abs(_product(X, Y) - np.array(g["product_reconstruct_r"])).max() < goldens["float_epsilon"]
    assert its == 34


def test_negative_input_golden(goldens):
    """NegativeInputTest.testALS (:37-80): negative strengths feed W but not b."""
    g = goldens["negative_input"]
    ptr, idx, val = dense_to_csr(g["R"])
    X, Y, its, _ = O.als_run(ptr, idx, val, 4, np.array(g["Y0"], np.float32),
                             convergence_threshold=g["threshold"],
                             max_iterations=g["max_iterations"])
    assert np.abs(_product(X, Y) - np.array(g["product"])).max() < goldens["float_epsilon"]
    assert its == 19


def test_transpose_times_self_golden(goldens):
    """MatrixUtilsTest.testTransposeTimesSelf (:63-73), exact to 1e-12."""
    g = goldens["transpose_times_self"]
    G = O.transpose_times_self(np.array(g["M"], np.float32))
    assert np.abs(G - np.array(g["MTM"])).max() <= goldens["double_epsilon"]


def test_simple_vector_math_golden(goldens):
    """SimpleVectorMathTest (:29-37)."""
    g = goldens["simple_vector_math"]
    assert abs(O.dot(g["vec1"], g["vec2"]) - g["dot"]) <= goldens["double_epsilon"]
    assert abs(O.norm(g["vec1"]) - g["norm1"]) <= goldens["double_epsilon"]
    assert abs(O.norm(g["vec2"]) - g["norm2"]) <= goldens["double_epsilon"]


def test_transpose_times_self_rounds_products_to_fp32():
    """MatrixUtils.java:231-233: float*float is rounded to fp32 before the fp64 add."""
    M = np.array([[1.0000001, 3.0000002]], np.float32)
    G = O.transpose_times_self(M)
    assert G[0, 1] == float(np.float32(M[0, 0] * M[0, 1]))
    assert G[0, 1] != float(M[0, 0]) * float(M[0, 1])


def test_solver_matches_numpy():
    rng = np.random.default_rng(0)
    for k in (1, 2, 7, 30, 64):
        A = rng.standard_normal((k + 5, k))
        W = A.T @ A + 0.3 * np.eye(k)
        b = rng.standard_normal(k)
        x = O.solve(W, b)
        ref = np.linalg.solve(W, b)
        assert np.abs(x - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_solver_singular_reports_rank():
    """CommonsMathLinearSystemSolver.java:43-54: |R_jj| <= 1e-5 -> SingularMatrixSolverException
    carrying getRank(0.01)."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((3, 6))
    W = A.T @ A  # rank 3, 6x6
    with pytest.raises(O.SingularMatrixError) as ei:
        O.solve(W, np.ones(6))
    assert ei.value.apparent_rank == 3


def test_half_leaves_empty_rows_untouched_and_threads_agree():
    rng = np.random.default_rng(2)
    from conftest import random_problem
    ptr, idx, val, Y0 = random_problem(300, 50, 8, 6, seed=3, empty_users=4)
    G = O.transpose_times_self(Y0)
    out1 = np.full((300, 6), 7.0, np.float32)
    out8 = np.full((300, 6), 7.0, np.float32)
    O.als_half(ptr, idx, val, Y0, G, out1, n_threads=1)
    O.als_half(ptr, idx, val, Y0, G, out8, n_threads=8)
    assert np.array_equal(out1, out8)
    assert np.all(out1[:4] == 7.0) and not np.any(out1[4:] == 7.0)


def test_half_matches_dense_normal_equations():
    """Worker.call (:438-502) restated densely in numpy fp64."""
    from conftest import random_problem
    k = 5
    ptr, idx, val, Y0 = random_problem(40, 30, 6, k, seed=4, neg_fraction=0.3)
    G = O.transpose_times_self(Y0)
    out = np.zeros((40, k), np.float32)
    O.als_half(ptr, idx, val, Y0, G, out, alpha=2.0, lam=0.05)
    Yd = Y0.astype(np.float64)
    for u in range(40):
        e = slice(ptr[u], ptr[u + 1])
        y = Yd[idx[e]]
        r = val[e].astype(np.float64)
        W = G + (y.T * (2.0 * np.abs(r))) @ y + 0.05 * 2.0 * len(r) * np.eye(k)
        b = (y.T * np.where(r > 0, 1 + 2.0 * np.abs(r), 0.0)).sum(axis=1)
        assert np.abs(out[u] - np.linalg.solve(W, b)).max() < 1e-6


def test_present_but_empty_rows_in_the_oracle_match_a_dense_solve():
    """Keys of RbyRow / RbyColumn whose maps removeSmall emptied are still solved by the
    reference (ALS.java:391-410; InputFilesReader.java:202-211): W = G, b = 0 -> zero vector.
    The oracle's `present` masks against a dense numpy restatement that shares no code with it."""
    from conftest import dense_als_numpy, dense_to_csr, rel_err
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    U, I, k = 40, 30, 6
    R = np.zeros((U, I), np.float32)
    for u in range(U):
        R[u, rng.choice(I - 2, 7, replace=False)] = rng.integers(1, 6, 7)
    R[5] = 0
    R[6] = 0
    d = rng.standard_normal((I, k))
    Y0 = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    ptr, idx, val = dense_to_csr(R)
    X, Y, its, _ = O.als_run(ptr, idx, val, I, Y0, max_iterations=3, convergence_threshold=1e-12,
                             present_users=[5, 6], present_items=[I - 2])
    Xn, Yn = dense_als_numpy(R, {5, 6}, {I - 2}, Y0, 3)
    assert its == 3 and np.all(X[5] == 0) and np.all(X[6] == 0) and np.all(Y[I - 2] == 0)
    assert np.array_equal(Y[I - 1], Y0[I - 1])  # stale item: not a key, untouched
    for a, b in ((X, Xn), (Y, Yn)):
        fro, mx = rel_err(a, b)
        assert fro <= 1e-6 and mx <= 1e-6, (fro, mx)
    # without the masks the rows keep their previous contents (absent from the map)
    X2, Y2, _, _ = O.als_run(ptr, idx, val, I, Y0, max_iterations=3, convergence_threshold=1e-12)
    assert np.array_equal(Y2[I - 2], Y0[I - 2])


def _ldlt_min_pivot(W):
    """Unpivoted right-looking LDL^T in fp64: the pivots csrc/solve_fp64.cuh judges (pivot <= threshold -> singular)."""
    W = W.copy()
    m = np.inf
    for j in range(W.shape[0]):
        d = W[j, j]
        m = min(m, d)
        if not d > 0:
            return d
        l = W[j + 1:, j] / d
        W[j + 1:, j + 1:] -= np.outer(l, W[j + 1:, j])
    return m


def test_where_the_ldlt_pivot_rule_and_the_rrqr_rule_part_ways():
    """The reference declares W singular when a diagonal entry of the pivoted QR has |R_jj| <= 1e-5
    (CommonsMathLinearSystemSolver.java:43-54); the CUDA fp64 path judges the LDL^T pivots against the same
    threshold.  For SPD matrices with one small eigenvalue both rules agree when lambda_min >= 1e-5 (solved)
    and when lambda_min <= 1e-13 (singular); in between the unpivoted pivot (~lambda_min / q^2, q the small
    eigenvector's last-eliminated component) overestimates lambda_min by more than the pivoted |R_kk| does, so
    the CUDA path still solves some systems the reference gives up on (known gaps).  Not reachable
    with lambda > 0: lambda_min(W_u) >= lambda alpha n_u."""
    rng = np.random.default_rng(0)
    between = 0
    for k in (8, 32, 64):
        for lam_min in (1e-14, 1e-13, 1e-10, 1e-8, 1e-7, 1e-6, 1e-5, 2e-5, 1e-4):
            for _ in range(8):
                Q, _r = np.linalg.qr(rng.standard_normal((k, k)))
                ev = rng.uniform(0.5, 2.0, k)
                ev[rng.integers(k)] = lam_min
                W = (Q * ev) @ Q.T
                W = (W + W.T) / 2
                try:
                    O.solve(W, np.ones(k))
                    rrqr_singular = False
                except O.SingularMatrixError:
                    rrqr_singular = True
                ldlt_singular = not (_ldlt_min_pivot(W) > 1e-5)
                if lam_min >= 1e-5:
                    assert not rrqr_singular and not ldlt_singular, (k, lam_min)
                elif lam_min <= 1e-13:
                    assert rrqr_singular and ldlt_singular, (k, lam_min)
                else:
                    between += rrqr_singular != ldlt_singular
    assert between > 0   # the band is real: keep DESIGN.md honest about it
