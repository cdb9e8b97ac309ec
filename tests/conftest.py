import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


def dense_to_csr(R):
    R = np.asarray(R, dtype=np.float32)
    ptr, idx, val = [0], [], []
    for r in R:
        nz = np.nonzero(r)[0]
        idx += list(nz)
        val += list(r[nz])
        ptr.append(len(idx))
    return np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32)


def dense_to_maps(R):
    """{user: {item: v}}, {item: {user: v}} in MatrixUtils.addTo call order
    (AlternatingLeastSquaresTest.java:92-105)."""
    by_row, by_col = {}, {}
    for u, r in enumerate(R):
        for i, v in enumerate(r):
            if v != 0:
                by_row.setdefault(u, {})[i] = float(v)
                by_col.setdefault(i, {})[u] = float(v)
    return by_row, by_col


def random_problem(n_users, n_items, nnz_per_user, k, seed, neg_fraction=0.05, empty_users=0,
                   stale_items=0):
    """Seeded sparse R (distinct items per user, strengths in +-{1..5}) and unit-row Y0."""
    rng = np.random.default_rng(seed)
    ptr = [0]
    idx, val = [], []
    live_items = n_items - stale_items
    for u in range(n_users):
        n = 0 if u < empty_users else min(nnz_per_user, live_items)
        cols = np.sort(rng.choice(live_items, size=n, replace=False))
        v = rng.integers(1, 6, size=n).astype(np.float32)
        v[rng.random(n) < neg_fraction] *= -1
        idx += list(cols)
        val += list(v)
        ptr.append(len(idx))
    d = rng.standard_normal((n_items, k))
    Y0 = (d / np.sqrt((d * d).sum(axis=1))[:, None]).astype(np.float32)
    return (np.array(ptr, np.int64), np.array(idx, np.int32), np.array(val, np.float32), Y0)


def rel_err(a, b):
    """(||a-b||_F/||b||_F, max|a-b|/max|b|) -- the two parity figures of SURVEY.md 8(d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return (float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)),
            float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))


def dense_als_numpy(R, present_u, present_i, Y0, iters, alpha=1.0, lam=0.1):
    """Independent fp64 restatement with dense numpy solves (NOT the oracle): a row is solved iff
    it is a key of the map; a key without entries gets W = G, b = 0."""
    R = np.asarray(R, np.float64)
    U, I = R.shape
    Y = np.asarray(Y0, np.float64).copy()
    X = np.zeros((U, Y.shape[1]))
    k = Y.shape[1]

    def half(Rm, Mf, out, present):
        G = (Mf.astype(np.float32).astype(np.float64)).T @ Mf.astype(np.float32).astype(np.float64)
        for u in range(Rm.shape[0]):
            nz = np.nonzero(Rm[u])[0]
            if nz.size == 0 and u not in present:
                continue
            ys = Mf[nz]
            r = Rm[u, nz]
            W = G + (ys.T * (alpha * np.abs(r))) @ ys + lam * alpha * nz.size * np.eye(k)
            b = ((1 + alpha * np.abs(r)) * (r > 0))[None, :] @ ys
            out[u] = np.linalg.solve(W, b.ravel()).astype(np.float32)
    for _ in range(iters):
        half(R, Y.astype(np.float32).astype(np.float64), X, present_u)
        half(R.T, X.astype(np.float32).astype(np.float64), Y, present_i)
    return X, Y
