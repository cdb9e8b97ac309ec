#!/usr/bin/env python
"""CSV files -> factor matrices, end to end on one B200: what DelegateGenerationManager's
refresh does with the CUDA factorizer plugged in (DelegateGenerationManager.java:333-343).

    python scripts/build_model.py INPUT_DIR --features 64 --iterations 10 --out model.npz

Reads every *.csv / *.csv.gz / *.csv.zip of INPUT_DIR (libmyrrix_ingest.so), runs a fixed number
of ALS iterations (libmyrrix_als.so) from random unit-norm item vectors, and writes
user_ids, item_ids (int64), X, Y (float32 [n, features]) to an .npz file -- or, when --out ends
in .gz, the reference's own model.bin.gz (libmyrrix_model_io.so).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("input_dir")
    ap.add_argument("--features", type=int, default=30)        # model.features default
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--alpha", type=float, default=1.0)        # model.als.alpha
    ap.add_argument("--lam", type=float, default=0.1)          # model.als.lambda
    ap.add_argument("--seed", type=int, default=1234567890)
    ap.add_argument("--out", default="model.npz")
    a = ap.parse_args()
    import myrrix_recommender_b200 as M
    t0 = time.time()
    r = M.ingest.read_input_files(a.input_dir)
    t1 = time.time()
    n_users, n_items = len(r.user_ids), len(r.item_ids)
    print("read %d lines (%d bad): %d users x %d items, %d entries in %.2f s"
          % (r.lines, r.bad_lines, n_users, n_items, r.col_idx.size, t1 - t0))
    if n_users == 0 or n_items == 0:
        sys.exit("no input")
    rng = np.random.default_rng(a.seed)
    y0 = rng.standard_normal((n_items, a.features))
    y0 = (y0 / np.sqrt((y0 * y0).sum(axis=1))[:, None]).astype(np.float32)  # RandomUtils.randomUnitVector
    with M.NativeALS(a.features, alpha=a.alpha, lam=a.lam) as als:
        als.set_interactions(n_users, n_items, r.row_ptr, r.col_idx, r.val)
        als.set_y(y0)
        als.iterate(a.iterations)
        als.sync()
        X, Y = als.get_x(), als.get_y()
    t2 = time.time()
    if a.out.endswith(".gz"):   # the reference's model.bin.gz (GenerationSerializer)
        item_of = r.item_ids
        g = M.model_io.Generation(r.user_ids, X, r.item_ids, Y, r.user_ids, r.known_ptr,
                                  item_of[r.known_idx], r.item_tag_ids, r.user_tag_ids)
        M.model_io.write_generation(g, a.out)
    else:
        np.savez(a.out, user_ids=r.user_ids, item_ids=r.item_ids, X=X, Y=Y)
    print("%d iterations in %.2f s -> %s" % (a.iterations, t2 - t1, a.out))


if __name__ == "__main__":
    main()
