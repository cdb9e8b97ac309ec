"""numpy twin of the device workload generator (csrc/aux_kernels.cuh: synth_rows_kernel,
synth_hash, mix64) -- TEST INFRASTRUCTURE ONLY.  Used to check the device generator
bit-for-bit and to build the same workload for the CPU baseline without a GPU."""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_C_ROW = np.uint64(0x9E3779B97F4A7C15)
_C_J = np.uint64(0xD1B54A32D192ED03)


def mix64(x):
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= _M1
    x ^= x >> np.uint64(27)
    x *= _M2
    x ^= x >> np.uint64(31)
    return x


def synth_hash(seed, rows, js):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) + rows.astype(np.uint64) * _C_ROW +
                     (js.astype(np.uint64) + np.uint64(1)) * _C_J)


def synth_rows(row_begin, n_rows, n_items, nnz_per_user, seed=1234567890, neg_fraction=0.0):
    """CSR (row_ptr int64, col_idx int32, val float32) of users [row_begin, row_begin+n_rows)."""
    rows = np.repeat(np.arange(row_begin, row_begin + n_rows, dtype=np.uint64), nnz_per_user)
    js = np.tile(np.arange(nnz_per_user, dtype=np.uint64), n_rows)
    h = synth_hash(seed, rows, js)
    j64 = js.astype(np.int64)
    lo = (j64 * n_items) // nnz_per_user
    hi = ((j64 + 1) * n_items) // nnz_per_user
    col = lo + ((h >> np.uint64(32)) % (hi - lo).astype(np.uint64)).astype(np.int64)
    s = (1 + ((h & np.uint64(0xffff)) % np.uint64(5)).astype(np.int64)).astype(np.float32)
    thr = np.uint64(int(neg_fraction * 16777216.0))
    neg = ((h >> np.uint64(8)) & np.uint64(0xffffff)) < thr
    s[neg] = -s[neg]
    row_ptr = np.arange(n_rows + 1, dtype=np.int64) * nnz_per_user
    return row_ptr, col.astype(np.int32), s


def _unit(h):
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def powerlaw_params(n_items, mean_nnz, max_nnz, seed):
    """(scale, perm_mul, perm_add) exactly as als_synth_interactions_powerlaw derives them."""
    import math
    ex = math.log(float(max_nnz)) * float(max_nnz) / (float(max_nnz) - 1.0)
    scale = mean_nnz / ex
    mul = (int(mix64(np.array([seed ^ 0x2545f4914f6cdd1d], dtype=np.uint64))[0]) % n_items) | 1
    while math.gcd(mul, n_items) != 1:
        mul += 2
    add = int(mix64(np.array([seed ^ 0x9e3779b97f4a7c15], dtype=np.uint64))[0]) % n_items
    return scale, mul, add


def powerlaw_counts(row_begin, n_rows, n_items, mean_nnz, max_nnz=20000, seed=1234567890):
    scale, _, _ = powerlaw_params(n_items, mean_nnz, max_nnz, seed)
    users = np.arange(row_begin, row_begin + n_rows, dtype=np.uint64)
    u = _unit(synth_hash(seed ^ 0x7c3a9f1d5b2e8461, users, np.zeros(n_rows, dtype=np.uint64)))
    x = 1.0 / (1.0 - u * (1.0 - 1.0 / float(max_nnz)))
    n = np.floor(scale * x + 0.5).astype(np.int64)
    return np.clip(n, 1, min(max_nnz, n_items))


def synth_rows_powerlaw(row_begin, n_rows, n_items, mean_nnz, max_nnz=20000, seed=1234567890, neg_fraction=0.0):
    """numpy twin of powerlaw_counts_kernel + powerlaw_rows_kernel: CSR of users [row_begin, +n_rows)."""
    _, mul, add = powerlaw_params(n_items, mean_nnz, max_nnz, seed)
    counts = powerlaw_counts(row_begin, n_rows, n_items, mean_nnz, max_nnz, seed)
    ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    rows = np.repeat(np.arange(row_begin, row_begin + n_rows, dtype=np.uint64), counts)
    js = np.arange(ptr[-1], dtype=np.int64) - np.repeat(ptr[:-1], counts)
    nn = np.repeat(counts, counts)
    h = synth_hash(seed, rows, js.astype(np.uint64))
    v = (js.astype(np.float64) + _unit(h)) / nn.astype(np.float64)
    rank = np.floor(np.expm1(v * np.log1p(float(n_items)))).astype(np.int64)
    # strictly increasing per user: rank_j = j + cummax(rank_j - j), then kept below n_items - (n - j)
    d = rank - js
    big = np.int64(4) * np.int64(n_items)
    seg = np.repeat(np.arange(n_rows, dtype=np.int64), counts) * big   # restarts the running maximum
    d = np.maximum.accumulate(d + seg) - seg
    rank = np.minimum(d + js, n_items - (nn - js))
    col = ((np.uint64(mul) * rank.astype(np.uint64) + np.uint64(add)) % np.uint64(n_items)).astype(np.int32)
    s = (1 + ((h & np.uint64(0xffff)) % np.uint64(5)).astype(np.int64)).astype(np.float32)
    thr = np.uint64(int(neg_fraction * 16777216.0))
    neg = ((h >> np.uint64(8)) & np.uint64(0xffffff)) < thr
    s[neg] = -s[neg]
    return ptr, col, s


def unit_rows(n_rows, k, seed=1234567890):
    """Y0: rows of k i.i.d. N(0,1) normalised to unit L2 (RandomUtils.java:88-100 distribution).
    numpy's RNG -- statistically, not bitwise, the device generator's rows."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n_rows, k))
    v = d.astype(np.float32)
    v /= np.sqrt((d * d).sum(axis=1)).astype(np.float32)[:, None]
    return v
