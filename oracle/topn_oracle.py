"""CPU restatement of the reference's top-N scoring -- TEST INFRASTRUCTURE ONLY.

Never imported by the product; it is the checker for csrc/topn.cuh.

Follows, line by line:
  * RecommendIterator.next (online/src/net/myrrix/online/RecommendIterator.java:68-110): for every
    item of Y that is not filtered, `sum += SimpleVectorMath.dot(itemFeatures, oneUserFeatures)`
    over the query's user vectors, `result = (float) (sum / count)`;
  * SimpleVectorMath.dot (common/.../math/SimpleVectorMath.java:34-41): float * float rounded to
    fp32, accumulated in fp64 in feature order;
  * TopN.selectTopNIntoQueue / selectTopNFromQueue (common/src/net/myrrix/common/TopN.java:55-75,
    122-131): a java.util.PriorityQueue of at most n + 1 entries ordered by ByValueAscComparator
    (common/.../ByValueAscComparator.java:37-55: value ascending, ties by item ID DEscending), strict
    `>` against the least entry once more than n are held, the surplus polled, the rest sorted in
    reverse comparator order (value descending, ties by item ID ascending);
  * the known-items filter of ServerRecommender.recommendToMany (online/.../ServerRecommender.java
    :396-425): the intersection of the known-item sets of the query's users that have one.

Pinned against the reference's own vectors: TopNTest.java:30-83 (empty, exactly n, n + 1, many) and
ByValueAscComparatorTest.java:28-58 (tests/test_topn.py).

The reference streams Y in hash-slot order, so which of several EQUAL scores at the cut survives
is unspecified there; the dense-index path streams in ascending item order, and for such a stream
the queue keeps the lower IDs -- i.e. the result is the first n of (value descending, ID
ascending).  `select_top_n` is the literal queue; `top_n_sorted` is that closed form.
"""
import heapq

import numpy as np


def by_value_asc_compare(a, b):
    """ByValueAscComparator.compare on (item, value) pairs."""
    (ai, av), (bi, bv) = a, b
    if av < bv:
        return -1
    if av > bv:
        return 1
    if ai > bi:
        return -1
    if ai < bi:
        return 1
    return 0


class _Entry:
    __slots__ = ("item", "value")

    def __init__(self, item, value):
        self.item, self.value = item, value

    def __lt__(self, other):
        return by_value_asc_compare((self.item, self.value), (other.item, other.value)) < 0


def select_top_n(stream, n):
    """TopN.selectTopN (TopN.java:138-142) over an iterable of (item, value) / None."""
    q = []
    for v in stream:
        if v is None:
            continue
        item, value = v
        value = np.float32(value)
        if len(q) > n:
            if value > q[0].value:
                heapq.heapreplace(q, _Entry(item, value))
        else:
            heapq.heappush(q, _Entry(item, value))
    if not q:
        return []
    while len(q) > n:
        heapq.heappop(q)
    out = sorted(q, reverse=True)
    return [(e.item, e.value) for e in out]


def scores(Y, features):
    """(float) (sum_v dot(y_i, x_v) / count) for every row of Y; features: [count, k] fp32."""
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    F = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, Y.shape[1])
    total = np.zeros(Y.shape[0], dtype=np.float64)
    for x in F:
        prod = Y * x[None, :]            # float * float -> fp32 (SimpleVectorMath.java:38)
        dot = np.zeros(Y.shape[0], dtype=np.float64)
        for f in range(Y.shape[1]):      # accumulated in fp64 in feature order
            dot += prod[:, f].astype(np.float64)
        total += dot
    return (total / float(F.shape[0])).astype(np.float32)


def top_n_sorted(values, n, excluded=None):
    """First n of (value descending, ID ascending) among the non-excluded rows."""
    values = np.asarray(values, dtype=np.float32)
    ids = np.arange(values.size, dtype=np.int64)
    keep = np.ones(values.size, dtype=bool)
    if excluded is not None and len(excluded):
        keep[np.asarray(list(excluded), dtype=np.int64)] = False
    ids, vals = ids[keep], values[keep]
    order = np.lexsort((ids, -vals.astype(np.float64)))[:n]
    return ids[order].astype(np.int32), vals[order]


def known_items(ptr, idx, users):
    """Intersection of the known-item sets of the users that have one
    (ServerRecommender.java:402-421); None when no user has one."""
    known = None
    for u in users:
        row = set(int(i) for i in idx[ptr[u]:ptr[u + 1]])
        if not row:
            continue
        known = row if known is None else (known & row)
        if not known:
            break
    return known


def recommend(Y, X, ptr, idx, users, how_many, consider_known_items=False, extra_excluded=()):
    """ServerRecommender.recommendToMany on dense indices: (items, values)."""
    s = scores(Y, X[np.asarray(users, dtype=np.int64)])
    excl = set(int(i) for i in extra_excluded)
    if not consider_known_items:
        k = known_items(ptr, idx, users)
        if k:
            excl |= k
    return top_n_sorted(s, how_many, excl)
