#!/usr/bin/env python
"""Small cases for compute-sanitizer (memcheck / racecheck): one iteration of config-1-like shapes through
the tensor-core kernel (k = 64 and k = 32, both role mixes), the CUDA-core kernel, the top-N kernels and
the device fold-in."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrrix_recommender_b200 as M  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
for k, kernel in ((64, 0), (32, 0), (16, 1)):
    if which not in ("all", str(k)):
        continue
    with M.NativeALS(k, kernel=kernel) as als:
        als.synth_interactions(600, 150, 20, seed=1234567890, neg_fraction=0.05)
        als.synth_y0(seed=1234567890)
        als.iterate(1)
        als.sync()
        X = als.get_x()
        assert np.isfinite(X).all()
        items, values = als.recommend([3], 10)
        als.recompute_state()
        als.fold_in([1, 2, 1], [5, 5, 6], [1.0, 2.0, -1.0])
        print("k=%d kernel=%d ok: |X|=%.4f top item %d" % (k, als.info().kernel, float(np.abs(X).sum()), items[0]))
