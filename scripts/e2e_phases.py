#!/usr/bin/env python
"""Where does one end-to-end factorizer call (host buffers in, host buffers out) spend its time?
Same call sequence as bench.py's e2e leg on C3, each C-ABI call timed with a device sync behind it."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myrrix_recommender_b200 as M  # noqa: E402

U, I, nnz_u, k = 10_000_000, 1_000_000, 100, 64
if len(sys.argv) > 1 and sys.argv[1] == "c2":
    U, I, nnz_u, k = 1_000_000, 100_000, 50, 32
steps = 5
lib = M._native.load()
als = M.NativeALS(k)
als.synth_interactions(U, I, nnz_u, seed=1234567890)
als.synth_y0(seed=1234567890)
nnz = U * nnz_u
h_ptr = torch.empty(U + 1, dtype=torch.int64, pin_memory=True)
h_idx = torch.empty(nnz, dtype=torch.int32, pin_memory=True)
h_val = torch.empty(nnz, dtype=torch.float32, pin_memory=True)
h_x = torch.empty((U, k), dtype=torch.float32, pin_memory=True)
h_y = torch.empty((I, k), dtype=torch.float32, pin_memory=True)
h_y0 = torch.empty((I, k), dtype=torch.float32, pin_memory=True)
P = lambda t, ty: C.cast(t.data_ptr(), C.POINTER(ty))
als.check(lib.als_get_interactions(als.h, P(h_ptr, C.c_int64), P(h_idx, C.c_int32), P(h_val, C.c_float)))
als.check(lib.als_get_y(als.h, P(h_y0, C.c_float)))
als.close()


def run(trace):
    t = [time.perf_counter()]
    names = []

    def mark(name):
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        names.append(name)
    a = M.NativeALS(k)
    mark("als_create")
    a.check(lib.als_set_interactions(a.h, U, I, P(h_ptr, C.c_int64), P(h_idx, C.c_int32), P(h_val, C.c_float)))
    a.n_users, a.n_items = U, I
    mark("als_set_interactions (upload + by-item build)")
    a.check(lib.als_set_y(a.h, P(h_y0, C.c_float)))
    mark("als_set_y")
    a.iterate(steps)
    a.sync()
    mark("als_iterate(%d) + als_sync" % steps)
    a.check(lib.als_get_x(a.h, P(h_x, C.c_float)))
    mark("als_get_x")
    a.check(lib.als_get_y(a.h, P(h_y, C.c_float)))
    mark("als_get_y")
    a.close()
    mark("als_destroy")
    if trace:
        for n, a0, a1 in zip(names, t[:-1], t[1:]):
            print("%-48s %8.1f ms" % (n, (a1 - a0) * 1e3))
        print("%-48s %8.1f ms" % ("total", (t[-1] - t[0]) * 1e3))


run(False)
run(True)
