#!/usr/bin/env python
"""Top-N scoring on a resident model: latency of one recommend call and
throughput of the batch call against the HBM roofline (every pass streams Y once:
algorithmic bytes per pass = items * 4 * features), with the CPU oracle timed beside it.

    python scripts/topn_bench.py [items] [features] [how_many] > profiles/rNN_topn.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myrrix_recommender_b200 as M  # noqa: E402

I = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
U, nnz = 200_000, 100
peak = 6513.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

with M.NativeALS(k, device=0) as als:
    als.synth_interactions(U, I, nnz, seed=1234567890)
    als.synth_y0(seed=1234567890)
    als.half_x()
    als.sync()
    users = np.random.default_rng(1).integers(0, U, 4096).astype(np.int32)
    out = {"workload": "%d items x k=%d resident, howMany=%d, known items filtered" % (I, k, N)}
    for _ in range(3):
        als.recommend([int(users[0])], N)
    t0 = time.perf_counter()
    reps = 50
    for j in range(reps):
        als.recommend([int(users[j])], N)
    t1 = (time.perf_counter() - t0) / reps
    bytes_pass = I * 4 * als.info().padded_features
    out["single_query"] = {"seconds": t1, "gbs": bytes_pass / t1 / 1e9, "frac_of_hbm": bytes_pass / t1 / 1e9 / peak,
                           "note": "host call to host result (filter bitmap, score + merge kernels, one packed read-back)"}
    als.recommend_batch(users[:64], N)
    t0 = time.perf_counter()
    items, values, counts = als.recommend_batch(users, N)
    tb = time.perf_counter() - t0
    passes = (users.size + 3) // 4
    out["batch"] = {"queries": int(users.size), "seconds": tb, "queries_per_s": users.size / tb,
                    "passes": passes, "gbs": passes * bytes_pass / tb / 1e9,
                    "frac_of_hbm": passes * bytes_pass / tb / 1e9 / peak}
    # the checker, timed on the host cores for one query (numpy restatement of RecommendIterator + TopN)
    from oracle import topn_oracle as T  # noqa: E402
    X, Y = als.get_rows("x", users[:1]), als.get_y()
    ptr, idx, _ = als.get_interaction_rows(int(users[0]), 1)
    t0 = time.perf_counter()
    s = T.scores(Y, X)
    oi, ov = T.top_n_sorted(s, N, excluded=set(int(i) for i in idx))
    out["cpu_oracle_single_query_seconds"] = time.perf_counter() - t0
    out["parity"] = bool(np.array_equal(oi, items[0]) and np.array_equal(ov, values[0]))
print(json.dumps(out))
