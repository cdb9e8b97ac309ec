#!/usr/bin/env python
"""Development diagnostic for the second-generation tensor-core kernel (GPU box only).

  1. the blocked solver alone (als_debug_solve_blocked) on random SPD systems;
  2. one X-half of a small k = 64 problem against the oracle;
  3. with a -DALS_DEBUG_SLOT build (MYRRIX_ALS_LIB=scripts/_var/dbg.so): the slot N = -W_u and the
     rhs of a few rows as the drain / producers left them, against numpy -- prints which 16 x 16
     blocks differ and how (transposed? permuted columns?), so a wrong layout assumption can be
     read off one run.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import myrrix_recommender_b200 as M  # noqa: E402
from conftest import random_problem, rel_err  # noqa: E402
from oracle import oracle as O  # noqa: E402

PS = 16


def panel_off(J, KS=64):
    return PS * (J * KS - 8 * J * (J - 1))


def swz(lr):
    return (((lr >> 1) & 1) << 1) | ((lr >> 2) & 1)


def slot_to_dense(slot, KS=64):
    """WPanels::at (chol_blocked.cuh): 64-byte panel rows, 16-byte chunks XOR-swizzled by the row."""
    W = np.full((KS, KS), np.nan)
    for c in range(KS):
        J = c // 16
        for i in range(16 * J, KS):
            lr = i - 16 * J
            W[i, c] = -slot[panel_off(J, KS) + lr * PS + ((c % 16) ^ (swz(lr) << 2))]
    return W


def main():
    lib = M._native.load()
    fp = C.POINTER(C.c_float)
    fn = lib.als_debug_solve_blocked
    fn.restype = C.c_int
    fn.argtypes = [fp, fp, C.c_int, C.c_float, fp, C.POINTER(C.c_int)]
    rng = np.random.default_rng(1)
    print("== 1. blocked solver alone")
    for k in (64, 48, 33, 32, 16, 5):
        q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        W = ((q * np.logspace(0, 1, k)) @ q.T * 30).astype(np.float32)
        b = rng.standard_normal(k).astype(np.float32)
        x = np.zeros(k, np.float32)
        ok = C.c_int(0)
        rc = fn(W.ctypes.data_as(fp), b.ctypes.data_as(fp), k, 1e-5, x.ctypes.data_as(fp), C.byref(ok))
        ref = np.linalg.solve(W.astype(np.float64), b.astype(np.float64))
        rc = fn(W.ctypes.data_as(fp), b.ctypes.data_as(fp), k, 1e-5, x.ctypes.data_as(fp), C.byref(ok))  # warm I-cache
        print("  k=%d rc=%d ok=%d err=%.2e  cycles (idle SM, 2nd run) %d" % (
            k, rc, ok.value, rel_err(x, ref)[0], lib.als_debug_last_solve_cycles()))
        if hasattr(lib, "als_debug_solve_prof"):
            buf = (C.c_longlong * 8)()
            lib.als_debug_solve_prof(buf)  # two solves since the last read
            names = ["panel load", "16 column steps", "write-back + fragments", "trailing HMMA blocks",
                     "backward: rows below + butterfly", "backward: in-block solve"]
            print("     per solve:", ", ".join("%s %d" % (n, buf[i] // 2) for i, n in enumerate(names)))
        if rel_err(x, ref)[0] > 1e-4 and k == 64:
            # which entries are wrong?
            bad = np.nonzero(np.abs(x - ref) > 1e-3 * np.abs(ref).max())[0]
            print("    wrong entries:", bad[:64])

    print("== 2. one X-half, k = 64, 300 users x 400 items x 100")
    k, U, I, nnz = 64, 300, 400, 100
    ptr, idx, val, Y0 = random_problem(U, I, nnz, k, seed=5)
    G = O.transpose_times_self(Y0)
    Xo = np.zeros((U, k), np.float32)
    O.als_half(ptr, idx, val, Y0, G, Xo)
    dbg = hasattr(lib, "als_debug_set_row")
    rows = [0, 1, 7, 150]
    for mix in ("8", "4"):
        os.environ["MYRRIX_ALS_MIX"] = mix
        for r in (rows if dbg else [-1]):
            if dbg:
                lib.als_debug_set_row.argtypes = [C.c_longlong]
                lib.als_debug_set_row(r)
            with M.NativeALS(k, kernel=2) as als:
                als.set_interactions(U, I, ptr, idx, val)
                als.set_y(Y0)
                als.half_x()
                try:
                    als.sync()
                except Exception as e:  # noqa: BLE001
                    print("  sync failed:", e)
                X = als.get_x()
                retried = als.timings().fp64_retry_rows
            err = np.abs(X - Xo).max(axis=1) / np.abs(Xo).max()
            print("  mix %s: X err fro %.2e, worst row %d (%.2e), rows > 1e-4: %d, fp64 retries %d"
                  % (mix, rel_err(X, Xo)[0], int(err.argmax()), err.max(), int((err > 1e-4).sum()), retried))
            if not dbg:
                continue
            buf = np.zeros(2560 + 64, np.float32)
            lib.als_debug_get_slot.argtypes = [fp, C.c_int]
            lib.als_debug_get_slot(buf.ctypes.data_as(fp), buf.size)
            Wd = slot_to_dense(buf[:2560])
            e0, e1 = ptr[r], ptr[r + 1]
            ys = Y0[idx[e0:e1]].astype(np.float64)
            rv = val[e0:e1].astype(np.float64)
            Wn = G + (ys.T * np.abs(rv)) @ ys + 0.1 * (e1 - e0) * np.eye(k)
            bn = (((1 + np.abs(rv)) * (rv > 0))[None, :] @ ys).ravel()
            print("   row %d: rhs err %.2e" % (r, rel_err(buf[2560:], bn)[0]))
            for bi in range(4):
                for bj in range(bi + 1):
                    a = Wd[16 * bi:16 * bi + 16, 16 * bj:16 * bj + 16]
                    b = Wn[16 * bi:16 * bi + 16, 16 * bj:16 * bj + 16]
                    if bi == bj:
                        a, b = np.tril(a), np.tril(b)
                    e = np.abs(a - b).max() / np.abs(Wn).max()
                    note = ""
                    if e > 1e-4:
                        if np.abs(a - b.T).max() / np.abs(Wn).max() < 1e-4:
                            note = " (transposed)"
                        else:
                            note = " first bad (i,j): %s" % (np.argwhere(np.abs(a - b) > 1e-3 * np.abs(Wn).max())[:6].tolist(),)
                    print("     block (%d,%d): err %.2e%s" % (bi, bj, e, note))


if __name__ == "__main__":
    main()
