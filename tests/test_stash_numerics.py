"""Why stash mode is accurate enough, checked in numpy without a GPU.

A row the fp32 solve refuses because G = Y'Y is ill-conditioned is re-solved as
(G_fp64 + D_tc + lambda alpha n I) x = b, where D_tc is the data term as the tensor-core path produces it:
every gathered row scaled by sqrt(alpha |r|), split into bf16 hi + bf16 lo, products summed in fp32.
The result must stay within the parity bar (1e-4) of the all-fp64 solve even when max diag / min pivot of
W is ~10^3, i.e. well behind the fast path's conditioning gate of 256.
"""
import numpy as np

TOL = 1e-4


def _bf16_rn(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _d_tensor_core(Yu, w):
    """sum_i w_i y_i y_i' the way the tensor-core path forms it: v = y sqrt(w) in fp32, v ~ hi + lo (bf16 each),
    (hi + lo)(hi + lo)' accumulated in fp32."""
    v = (Yu.astype(np.float32) * np.sqrt(w.astype(np.float32))[:, None]).astype(np.float32)
    hi = _bf16_rn(v)
    lo = _bf16_rn((v - hi).astype(np.float32))
    acc = np.zeros((Yu.shape[1], Yu.shape[1]), np.float32)
    for a, b in ((hi, hi), (hi, lo), (lo, hi), (lo, lo)):
        acc = (acc + (a.T.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)).astype(np.float32)
    return acc


def test_fp64_resolve_from_the_tensor_core_data_term_stays_within_the_parity_bar():
    rng = np.random.default_rng(5)
    k, n_items, lam_alpha = 64, 1200, 0.1
    d = 0.03 * rng.standard_normal((n_items, k))
    d[:, 0] += 1.0
    Y = (d / np.sqrt((d * d).sum(1))[:, None]).astype(np.float32)
    G = Y.astype(np.float64).T @ Y.astype(np.float64)
    worst, ratios = 0.0, []
    for n in (1, 2, 5, 9, 17, 40, 129, 600):
        for _ in range(6):
            idx = rng.choice(n_items, size=n, replace=False)
            r = rng.integers(1, 6, size=n).astype(np.float32)
            Yu = Y[idx]
            b = ((1.0 + r.astype(np.float64))[:, None] * Yu.astype(np.float64)).sum(0)
            D64 = (Yu.astype(np.float64) * r.astype(np.float64)[:, None]).T @ Yu.astype(np.float64)
            W64 = G + D64 + lam_alpha * n * np.eye(k)
            x_ref = np.linalg.solve(W64, b)
            L = np.linalg.cholesky(W64)
            ratios.append(W64.diagonal().max() / (L.diagonal() ** 2).min())
            # stash mode: fp64 Gramian + tensor-core data term, fp64 solve, (float) cast of the result
            W = G + _d_tensor_core(Yu, r).astype(np.float64) + lam_alpha * n * np.eye(k)
            x = np.linalg.solve(W, b).astype(np.float32)
            worst = max(worst, float(np.abs(x - x_ref).max() / np.abs(x_ref).max()))
    assert np.median(ratios) > 256          # these rows do sit behind the gate
    assert worst <= TOL / 4, worst           # and the fp64 re-solve from D_tc is well inside the bar
