"""Top-N scoring: the oracle against the reference's own vectors (CPU), the CUDA
path against the oracle, bit for bit (GPU)."""
import numpy as np
import pytest

from oracle import topn_oracle as T


# ---- the oracle against the reference's tests (CPU) ----------------------------------------
def _candidates(n):  # TopNTest.makeNCandidates (common/test/net/myrrix/common/TopNTest.java:76-82)
    return [(i, float(i)) for i in range(1, n + 1)]


def test_topn_empty():  # TopNTest.java:30-35
    assert T.select_top_n(iter([]), 2) == []


@pytest.mark.parametrize("n_cand,first,last", [(3, 3, 1), (4, 4, 2), (20, 20, 18)])
def test_topn_reference_vectors(n_cand, first, last):  # TopNTest.java:37-74
    top3 = T.select_top_n(iter(_candidates(n_cand)), 3)
    assert len(top3) == 3
    assert top3[0] == (first, np.float32(first))
    assert top3[2] == (last, np.float32(last))


def test_by_value_asc_comparator_reference_vectors():  # ByValueAscComparatorTest.java:28-58
    a, b = (1, 2.0), (5, 1.0)
    assert T.by_value_asc_compare(a, b) > 0 and T.by_value_asc_compare(b, a) < 0
    assert T.by_value_asc_compare(a, a) == 0
    a, b = (1, 2.0), (5, 2.0)  # same value: the larger ID sorts first (ascending order)
    assert T.by_value_asc_compare(a, b) > 0 and T.by_value_asc_compare(b, a) < 0


def test_queue_and_closed_form_agree_on_ascending_streams():
    rng = np.random.default_rng(5)
    for n_items, n in [(50, 10), (7, 10), (300, 1), (300, 64)]:
        vals = rng.integers(0, 12, n_items).astype(np.float32)  # many ties, also at the cut
        lit = T.select_top_n(((i, v) for i, v in enumerate(vals)), n)
        ids, vv = T.top_n_sorted(vals, n)
        assert [i for i, _ in lit] == list(ids)
        assert [v for _, v in lit] == list(vv)
    assert T.select_top_n(iter([None, (3, 1.0), None]), 2) == [(3, np.float32(1.0))]


def test_scores_are_fp32_products_summed_in_fp64():
    # a case where a plain fp32 dot, or fp64 products, round differently
    y = np.array([[1.0000001, 3.0000002, -2.9999998, 1e-3]], np.float32)
    x = np.array([[0.3333333, 0.1111111, 0.1111112, 7.0]], np.float32)
    want = np.float32(sum(float(np.float32(a * b)) for a, b in zip(y[0], x[0])))
    assert T.scores(y, x)[0] == want
    two = T.scores(y, np.vstack([x, 2 * x]))[0]
    d1 = sum(float(np.float32(a * b)) for a, b in zip(y[0], x[0]))
    d2 = sum(float(np.float32(a * b)) for a, b in zip(y[0], (2 * x)[0]))
    assert two == np.float32((d1 + d2) / 2.0)


def test_known_items_are_the_intersection_of_the_non_empty_sets():
    ptr = np.array([0, 3, 3, 5, 8], np.int64)
    idx = np.array([1, 2, 3, 2, 9, 2, 3, 7], np.int32)
    assert T.known_items(ptr, idx, [0]) == {1, 2, 3}
    assert T.known_items(ptr, idx, [1]) is None
    assert T.known_items(ptr, idx, [0, 1, 3]) == {2, 3}
    assert T.known_items(ptr, idx, [0, 2, 3]) == {2}


# ---- CUDA path against the oracle (GPU) -------------------------------------------------------
@pytest.fixture(scope="module")
def M():
    import myrrix_recommender_b200 as M
    return M


def _model(M, k, U, I, nnz, seed=7):
    als = M.NativeALS(k, device=0)
    als.synth_interactions(U, I, nnz, seed=seed, neg_fraction=0.05)
    als.synth_y0(seed=seed)
    als.iterate(2)
    als.sync()
    return als


@pytest.mark.gpu
@pytest.mark.parametrize("k,U,I,nnz", [(64, 3000, 5000, 30), (30, 2000, 777, 20), (128, 500, 1500, 10),
                                       (16, 800, 129, 8), (3, 300, 4000, 5), (32, 1000, 20000, 12)])
def test_recommend_is_bit_identical_to_the_oracle(M, k, U, I, nnz):
    with _model(M, k, U, I, nnz) as als:
        X, Y = als.get_x(), als.get_y()
        ptr, idx, _ = als.get_interactions()
        rng = np.random.default_rng(k)
        for user in rng.integers(0, U, 6):
            for how_many, known in [(10, False), (1, True), (100, False), (128, True)]:
                items, values = als.recommend([user], how_many, consider_known_items=known)
                oi, ov = T.recommend(Y, X, ptr, idx, [user], how_many, consider_known_items=known)
                assert np.array_equal(items, oi), (user, how_many, known)
                assert np.array_equal(values, ov)
        # several users per query: mean score, intersection of known items, extra exclusions
        for n_u in (2, 3, 8, 11, 16):
            users = rng.integers(0, U, n_u)
            ex = rng.integers(0, I, 40)
            items, values = als.recommend(users, 25, consider_known_items=False, exclude=ex)
            oi, ov = T.recommend(Y, X, ptr, idx, users, 25, consider_known_items=False, extra_excluded=ex)
            assert np.array_equal(items, oi), n_u
            assert np.array_equal(values, ov)


@pytest.mark.gpu
def test_recommend_batch_matches_single_queries_and_the_oracle(M):
    with _model(M, 64, 4000, 6000, 25) as als:
        X, Y = als.get_x(), als.get_y()
        ptr, idx, _ = als.get_interactions()
        users = np.random.default_rng(3).integers(0, 4000, 203)
        for known in (False, True):
            items, values, counts = als.recommend_batch(users, 10, consider_known_items=known)
            assert np.all(counts == 10)
            for j in range(0, users.size, 17):
                oi, ov = T.recommend(Y, X, ptr, idx, [users[j]], 10, consider_known_items=known)
                assert np.array_equal(items[j], oi) and np.array_equal(values[j], ov)
            si, sv = als.recommend([users[5]], 10, consider_known_items=known)
            assert np.array_equal(si, items[5]) and np.array_equal(sv, values[5])


@pytest.mark.gpu
def test_ties_short_lists_and_anonymous_vectors(M):
    k, I = 8, 1000
    with M.NativeALS(k, device=0) as als:
        ptr = np.array([0, 2, 2], np.int64)
        als.set_interactions(2, I, ptr, np.array([3, 5], np.int32), np.array([1, 1], np.float32))
        Y = np.zeros((I, k), np.float32)
        Y[:, 0] = np.repeat(np.arange(I // 4, dtype=np.float32), 4)  # every score four times
        Y[7] = 0
        Y[7, 1] = -0.0
        als.set_y(Y)
        X = np.zeros((2, k), np.float32)
        X[0, 0] = 1.0
        als.set_x(X)
        items, values = als.recommend([0], 10, consider_known_items=True)
        oi, ov = T.top_n_sorted(T.scores(Y, X[:1]), 10)
        assert np.array_equal(items, oi) and np.array_equal(values, ov)  # equal scores: ascending item
        items, values = als.recommend([0], 128, consider_known_items=False, exclude=np.arange(10, I))
        assert list(items) == [8, 9, 4, 6, 0, 1, 2, 7]  # items 3 and 5 are known; equal scores by ascending item
        # caller-supplied vectors (recommendToAnonymous after the caller's fold-in), and user rows
        f = np.random.default_rng(1).standard_normal((3, k)).astype(np.float32)
        ids, vals = als.top_n("y", f, 20, exclude=[1, 2])
        oi, ov = T.top_n_sorted(T.scores(Y, f), 20, excluded=[1, 2])
        assert np.array_equal(ids, oi) and np.array_equal(vals, ov)
        ids, vals = als.top_n("x", f[:1], 5)
        oi, ov = T.top_n_sorted(T.scores(X, f[:1]), 5)
        assert np.array_equal(ids, oi) and np.array_equal(vals, ov)
        with pytest.raises(Exception):
            als.recommend([0], 129)
        with pytest.raises(ValueError):
            als.recommend([5], 10)


@pytest.mark.gpu
def test_recommender_mirror_with_long_ids(M):
    from myrrix_recommender_b200.recommender import NoSuchUserException, Recommender
    with _model(M, 16, 500, 300, 10) as als:
        X, Y = als.get_x(), als.get_y()
        ptr, idx, _ = als.get_interactions()
        uid = 10_000_000_000 + 7 * np.arange(500, dtype=np.int64)
        iid = -5 * np.arange(300, dtype=np.int64) - 3
        rec = Recommender(als, uid, iid)
        got = rec.recommend(int(uid[42]), 5)
        oi, ov = T.recommend(Y, X, ptr, idx, [42], 5)
        assert got == [(int(iid[i]), float(v)) for i, v in zip(oi, ov)]
        got = rec.recommendToMany([int(uid[1]), 123, int(uid[2])], 7, considerKnownItems=True)
        oi, ov = T.recommend(Y, X, ptr, idx, [1, 2], 7, consider_known_items=True)
        assert got == [(int(iid[i]), float(v)) for i, v in zip(oi, ov)]
        with pytest.raises(NoSuchUserException):
            rec.recommend(123, 5)


@pytest.mark.gpu
def test_denormal_products_and_huge_scores_keep_the_reference_arithmetic(M):
    """Products below 1.2e-38 (denormal floats) and sums near the float range: the widening of the
    fp32 products to fp64 must stay exact (the kernel leaves its integer fast path for these)."""
    k, I = 8, 700
    rng = np.random.default_rng(9)
    with M.NativeALS(k, device=0) as als:
        als.set_interactions(1, I, np.array([0, 1], np.int64), np.array([0], np.int32), np.array([1], np.float32))
        for scale_y, scale_x in ((1e-20, 1e-19), (1e-30, 1e-9), (1e18, 1e19), (1.0, 1.0)):
            Y = (rng.standard_normal((I, k)) * scale_y).astype(np.float32)
            Y[::7] = 0.0
            f = (rng.standard_normal((2, k)) * scale_x).astype(np.float32)
            als.set_y(Y)
            ids, vals = als.top_n("y", f, 50)
            oi, ov = T.top_n_sorted(T.scores(Y, f), 50)
            assert np.array_equal(ids, oi), (scale_y, scale_x)
            assert np.array_equal(vals, ov)
