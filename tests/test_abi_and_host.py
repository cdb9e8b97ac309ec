"""CPU-side checks: the C-ABI library loads and exports every symbol include/myrrix_als.h
declares (no compute calls without a GPU), the product fails loudly without a GPU, and the
host-side logic (stop rule pieces, flattening, preconditions, synthetic twin)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, dense_to_maps

import myrrix_recommender_b200 as M
from myrrix_recommender_b200 import _native as N
from myrrix_recommender_b200.factorizer import DoubleWeightedMean, ExecutionException


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "myrrix_als.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(als_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    bound = {s[0] for s in N.SYMBOLS}
    assert set(declared) == bound, set(declared) ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.als_abi_version() == 1
    assert lib.als_comm_unique_id_size() == 128


def test_config_defaults_are_the_reference_defaults():
    lib = N.load()
    cfg = N.AlsConfig()
    assert lib.als_config_default(C.byref(cfg)) == N.ALS_OK
    assert cfg.struct_size == C.sizeof(N.AlsConfig)
    # MatrixFactorizer.java:34, AlternatingLeastSquares.java:71-73, LinearSystemSolver.java:33-34
    assert (cfg.features, cfg.alpha, cfg.lambda_, cfg.singularity_threshold) == (30, 1.0, 0.1, 1e-5)
    assert cfg.reconstruct_r == 0 and cfg.loss_ignores_unspecified == 0


def test_null_and_bad_arguments_are_rejected_without_a_gpu():
    lib = N.load()
    assert lib.als_create(None, None) == N.ALS_E_ARG
    cfg = N.AlsConfig()
    lib.als_config_default(C.byref(cfg))
    h = C.c_void_p()
    cfg.features = 0
    assert lib.als_create(C.byref(cfg), C.byref(h)) == N.ALS_E_ARG
    cfg.features = 10**6
    assert lib.als_create(C.byref(cfg), C.byref(h)) == N.ALS_E_ARG
    assert lib.als_half_x(None) == N.ALS_E_ARG
    assert lib.als_last_error(None) == b"null handle"
    assert lib.als_destroy(None) == N.ALS_OK


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ExecutionException):
        M.NativeALS(16)
    by_row, by_col = dense_to_maps([[1, 0], [0, 2]])
    als = M.AlternatingLeastSquares(by_row, by_col, 2, 0.001, 3)
    with pytest.raises(ExecutionException):
        als.call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "myrrix-recommender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "als_oracle" not in src and "oracle/" not in src.replace("oracle/.", ""), f


def test_constructor_preconditions():
    """AlternatingLeastSquares.java:137-141."""
    by_row, by_col = dense_to_maps([[1, 0], [0, 2]])
    with pytest.raises(TypeError):
        M.AlternatingLeastSquares(None, by_col)
    with pytest.raises(ValueError):
        M.AlternatingLeastSquares(by_row, by_col, 0)
    for thr in (0.0, 1.0, -0.5):
        with pytest.raises(ValueError):
            M.AlternatingLeastSquares(by_row, by_col, 2, thr, 3)


def test_double_weighted_mean_semantics():
    """DoubleWeightedMean.java:73-81: first datum wins even with weight 0."""
    m = DoubleWeightedMean()
    assert np.isnan(m.getResult())
    m.increment(5.0, 0.0)
    assert m.getResult() == 5.0
    m.increment(3.0, 0.0)      # total weight still <= 0 -> mean = datum
    assert m.getResult() == 3.0
    m.increment(1.0, 2.0)      # old weight 0 -> mean = datum
    assert m.getResult() == 1.0
    m.increment(4.0, 2.0)
    assert m.getResult() == pytest.approx(2.5)


def test_flatten_maps_to_csr():
    R = [[0, 2, 3], [0, 0, 0], [1, 0, 5]]
    by_row, by_col = dense_to_maps(R)
    by_row[1] = {}  # present but empty (InputFilesReader.removeSmall can leave these)
    uidx = {0: 0, 1: 1, 2: 2}
    iidx = {0: 0, 1: 1, 2: 2}
    ptr, idx, val = M.AlternatingLeastSquares._flatten(by_row, uidx, iidx)
    assert list(ptr) == [0, 2, 2, 4]
    assert list(idx) == [1, 2, 0, 2] and list(val) == [2, 3, 1, 5]


def test_initial_y_adopts_previous_and_fills_missing_with_unit_vectors():
    """constructInitialY, ALS.java:264-335."""
    by_row, by_col = dense_to_maps([[1, 1, 0], [0, 1, 1]])
    prev = {0: np.array([0.6, 0.8], np.float32), 7: np.array([1.0, 0.0], np.float32)}
    als = M.AlternatingLeastSquares(by_row, by_col, 2, 0.001, 3)
    Y = als._construct_initial_y(prev)
    assert Y is prev                         # same feature count: adopted in place
    assert set(Y) == {0, 7, 1, 2}            # stale row 7 kept, items 1,2 added
    for i in (1, 2):
        assert abs(np.linalg.norm(Y[i]) - 1) < 1e-6
    als3 = M.AlternatingLeastSquares(by_row, by_col, 3, 0.001, 3)
    Y3 = als3._construct_initial_y({0: np.array([0.6, 0.8], np.float32)})
    assert all(len(v) == 3 and abs(np.linalg.norm(v) - 1) < 1e-6 for v in Y3.values())
    als1 = M.AlternatingLeastSquares(by_row, by_col, 1, 0.001, 3)
    Y1 = als1._construct_initial_y({0: np.array([0.6, 0.8], np.float32)})
    assert Y1[0][0] == pytest.approx(1.0)


def test_synth_twin_properties():
    from oracle import synth
    ptr, idx, val = synth.synth_rows(10, 200, 1000, 20, seed=1234567890, neg_fraction=0.05)
    assert ptr[-1] == 4000 and idx.min() >= 0 and idx.max() < 1000
    rows = idx.reshape(200, 20)
    assert np.all(np.diff(rows, axis=1) > 0)          # distinct and ascending per user
    assert set(np.abs(val)) == {1, 2, 3, 4, 5}
    assert 0.02 < (val < 0).mean() < 0.09
    p2, i2, v2 = synth.synth_rows(0, 400, 1000, 20, seed=1234567890, neg_fraction=0.05)
    assert np.array_equal(i2[10 * 20:210 * 20], idx)  # counter-based: shard == slice of whole
    assert np.array_equal(v2[10 * 20:210 * 20], val)


def test_powerlaw_twin_properties():
    """The power-law workload of config 5: skewed row lengths, Zipf-like item popularity,
    no item twice per user, counter-based (a shard is a slice of the whole)."""
    from oracle import synth
    ptr, idx, val = synth.synth_rows_powerlaw(0, 20000, 5000, 40, max_nnz=2000, neg_fraction=0.05)
    n = np.diff(ptr)
    assert n.min() >= 1 and n.max() <= 2000 and 30 < n.mean() < 50
    assert n.max() > 20 * np.median(n)                     # a heavy tail of long rows
    assert idx.min() >= 0 and idx.max() < 5000
    for u in range(0, 20000, 61):
        row = idx[ptr[u]:ptr[u + 1]]
        assert len(set(row.tolist())) == row.size          # distinct items per user
    pop = np.sort(np.bincount(idx, minlength=5000))[::-1]
    assert pop[0] > 100 * np.median(pop)                   # a few very popular items
    assert set(np.abs(val)) == {1, 2, 3, 4, 5} and 0.02 < (val < 0).mean() < 0.09
    p2, i2, v2 = synth.synth_rows_powerlaw(100, 300, 5000, 40, max_nnz=2000, neg_fraction=0.05)
    assert np.array_equal(i2, idx[ptr[100]:ptr[400]]) and np.array_equal(v2, val[ptr[100]:ptr[400]])
    full = synth.synth_rows_powerlaw(0, 50, 30, 40, max_nnz=2000)  # rows as long as the item set
    assert np.diff(full[0]).max() == 30
    for u in range(50):
        assert len(set(full[1][full[0][u]:full[0][u + 1]].tolist())) == full[0][u + 1] - full[0][u]
