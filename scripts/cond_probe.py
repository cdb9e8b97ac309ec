"""How ill-conditioned are the per-row systems on the headline workload?  Runs a few device
iterations on a BASELINE config, pulls G = Y^T Y / X^T X and a sample of rows back, and compares
an fp32 Cholesky solve with the fp64 solve on the host (numpy) for each sampled row."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myrrix_recommender_b200 as M
import scipy.linalg as sla

cfgs = {"c2": (1000000, 100000, 50, 32), "c3": (10000000, 1000000, 100, 64)}
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
U, I, nnz, k = cfgs[name]
with M.NativeALS(k) as als:
    als.synth_interactions(U, I, nnz, seed=1234567890)
    als.synth_y0(seed=1234567890)
    for it in range(iters):
        t0 = time.time(); als.iterate(1); als.sync()
        GY = als.gramian("y"); GX = als.gramian("x")
        ey = np.linalg.eigvalsh(GY); ex = np.linalg.eigvalsh(GX)
        print("iter %d (%.2fs): eig(G_Y) min %.4g max %.4g cond %.3g | eig(G_X) min %.4g max %.4g cond %.3g"
              % (it + 1, time.time() - t0, ey[0], ey[-1], ey[-1] / ey[0], ex[0], ex[-1], ex[-1] / ex[0]))
    Y = als.get_y()
    ns = 300
    ptr, idx, val = als.get_interaction_rows(12345, ns, by_column=False, capacity=ns * nnz)
    X = als.get_x()
Yd = Y.astype(np.float64)
worst32 = 0; errs = []; conds = []
for u in range(ns):
    e = slice(ptr[u], ptr[u + 1]); y = Yd[idx[e]]; r = val[e].astype(np.float64)
    A = (y.T * np.abs(r)) @ y
    W = GY + A + 0.1 * len(r) * np.eye(k)
    b = (y.T * np.where(r > 0, 1 + np.abs(r), 0)).sum(axis=1)
    x64 = np.linalg.solve(W, b)
    W32 = W.astype(np.float32); b32 = b.astype(np.float32)
    c, low = sla.cho_factor(W32, lower=True)
    x32 = sla.cho_solve((c, low), b32)
    # one refinement step with fp64 residual
    r1 = b - W @ x32.astype(np.float64)
    x32r = x32.astype(np.float64) + sla.cho_solve((c, low), r1.astype(np.float32)).astype(np.float64)
    errs.append((np.abs(x32 - x64).max() / np.abs(x64).max(), np.abs(x32r - x64).max() / np.abs(x64).max()))
    conds.append(np.linalg.cond(W))
errs = np.array(errs)
print("rows sampled %d: cond(W) median %.3g max %.3g" % (ns, np.median(conds), np.max(conds)))
print("fp32 Cholesky rel err: median %.3g max %.3g | after 1 refinement step: median %.3g max %.3g"
      % (np.median(errs[:, 0]), errs[:, 0].max(), np.median(errs[:, 1]), errs[:, 1].max()))
print("device X vs fp64 solve on same rows: max rel %.3g" % max(
    np.abs(X[12345 + u] - np.linalg.solve(GY + ((Yd[idx[ptr[u]:ptr[u+1]]].T * np.abs(val[ptr[u]:ptr[u+1]].astype(np.float64))) @ Yd[idx[ptr[u]:ptr[u+1]]]) + 0.1 * (ptr[u+1]-ptr[u]) * np.eye(k),
        (Yd[idx[ptr[u]:ptr[u+1]]].T * np.where(val[ptr[u]:ptr[u+1]] > 0, 1 + np.abs(val[ptr[u]:ptr[u+1]].astype(np.float64)), 0)).sum(axis=1))).max() for u in range(0, 20)))
