// aux_kernels.cuh -- everything around the two hot kernels: CSR transposition
// (RbyRow -> RbyColumn), the convergence probe, the synthetic workload generator
// and the rare-path apparent-rank estimate.
#pragma once
#include <math.h>

#include "common.cuh"

namespace als {

// ---------------------------------------------------------------------------
// CSR -> (key=column, value=row<<32|valbits) expansion for the stable radix sort.
// row_offset: global index of local row 0 (stored in the packed value); col_offset is
// subtracted from the column so the keys are local to the target shard.
__global__ void expand_rows_kernel(const long long* __restrict__ row_ptr, long long n_rows,
                                   long long row_offset, int col_offset,
                                   const int* __restrict__ col_idx, const float* __restrict__ val,
                                   int* __restrict__ keys, unsigned long long* __restrict__ packed) {
  // one warp per row: coalesced over the row's entries
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) / kWarp;
  for (long long r = warp; r < n_rows; r += n_warps) {
    const long long e0 = row_ptr[r], e1 = row_ptr[r + 1];
    for (long long e = e0 + lane; e < e1; e += kWarp) {
      keys[e] = col_idx[e] - col_offset;
      packed[e] = ((unsigned long long)(unsigned int)(r + row_offset) << 32) |
                  (unsigned long long)__float_as_uint(val[e]);
    }
  }
}

// Sorted keys -> col_ptr (col_ptr[c] = first position with key >= c), and unpack.
__global__ void build_ptr_unpack_kernel(const int* __restrict__ keys_sorted,
                                        const unsigned long long* __restrict__ packed_sorted,
                                        long long nnz, long long n_cols,
                                        long long* __restrict__ col_ptr, int* __restrict__ row_idx,
                                        float* __restrict__ val) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int kcur = keys_sorted[e];
    const int kprev = (e == 0) ? -1 : keys_sorted[e - 1];
    for (int c = kprev + 1; c <= kcur; c++) col_ptr[c] = e;
    if (e == nnz - 1)
      for (long long c = (long long)kcur + 1; c <= n_cols; c++) col_ptr[c] = nnz;
    const unsigned long long pk = packed_sorted[e];
    row_idx[e] = (int)(pk >> 32);
    val[e] = __uint_as_float((unsigned int)(pk & 0xffffffffULL));
  }
}

__global__ void accumulate_count_kernel(const int* count, long long* total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *total += (long long)*count;
}

// total[0] += records, total[1] += records: rows re-solved in fp64 from the stash (records = min(count, cap))
__global__ void accumulate_resolve_kernel(const int* count, int cap, long long* total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const long long n = *count < cap ? *count : cap;
    total[0] += n;
    total[1] += n;
  }
}

__global__ void fill_ptr_zero_kernel(long long* ptr, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ptr[i] = 0;
}

// ---------------------------------------------------------------------------
// Convergence probe: out[i*ni+j] = dot(X[users[i]], Y[items[j]]), fp32-rounded
// products summed in fp64 in index order (SimpleVectorMath.java:34-41).
__global__ void probe_kernel(const float* __restrict__ X, const float* __restrict__ Y, int ks, int k,
                             const int* __restrict__ users, int nu, const int* __restrict__ items,
                             int ni, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nu * ni) return;
  const float* x = X + (long long)users[t / ni] * ks;
  const float* y = Y + (long long)items[t % ni] * ks;
  double dot = 0.0;
  for (int f = 0; f < k; f++) dot += (double)__fmul_rn(x[f], y[f]);
  out[t] = dot;
}

// The stop rule's statistic (AlternatingLeastSquares.java:232-240) over the probe's fresh
// estimates, on the device: DoubleWeightedMean.increment (common/.../stats/DoubleWeightedMean.java
// :73-81) is an order-dependent fp64 recurrence, so ONE thread walks the (user, item) pairs in the
// reference's order (at most ~1e4 pairs: about 0.2 ms per iteration).  estimates[] <- the new values.
// out[0] = mean (getResult()), out[1] = total weight.
__global__ void stop_rule_kernel(const double* __restrict__ fresh, double* __restrict__ estimates, int n,
                                 double* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double total_weight = 0.0, mean = 0.0;
  for (int t = 0; t < n; t++) {
    const double nv = fresh[t], ov = estimates[t];
    estimates[t] = nv;
    const double datum = fabs(nv - ov);       // FastMath.abs(newValue - oldValue)
    const double weight = fmax(0.0, nv);      // FastMath.max(0.0, newValue): NaN stays NaN like Java's max
    const double weight_j = (nv != nv) ? nv : weight;
    const double old_total = total_weight;
    total_weight = __dadd_rn(total_weight, weight_j);
    if (old_total <= 0.0) {
      mean = datum;
    } else {
      mean = __dadd_rn(__ddiv_rn(__dmul_rn(mean, old_total), total_weight),
                       __ddiv_rn(__dmul_rn(datum, weight_j), total_weight));
    }
  }
  out[0] = mean;
  out[1] = total_weight;
}

// ---- long rows split into chunks (virtual rows; see RowUpdateParams) -------------------------------
// nch[r] = chunks of row r (1 unless it has more than `limit` entries), is_long[r] = 1 for split rows
__global__ void split_counts_kernel(const long long* __restrict__ ptr, long long n_rows, long long limit,
                                    long long* __restrict__ nch, int* __restrict__ is_long) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += stride) {
    long long c = 0;
    int l = 0;
    if (r < n_rows) {
      const long long n = ptr[r + 1] - ptr[r];
      c = n > limit ? (n + limit - 1) / limit : 1;
      l = c > 1;
    }
    nch[r] = c;
    is_long[r] = l;
  }
}
// vfirst / accfirst: exclusive prefix sums of nch / is_long.  Chunks of a row are equal (rounded up to a
// multiple of 32 entries so that all but the last end on a stage boundary).
__global__ void split_fill_kernel(const long long* __restrict__ ptr, long long n_rows, long long limit,
                                  const long long* __restrict__ vfirst, const int* __restrict__ accfirst,
                                  long long* __restrict__ vptr, int* __restrict__ vrow, int* __restrict__ vacc,
                                  int* __restrict__ acc_chunks) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    const long long e0 = ptr[r], n = ptr[r + 1] - e0;
    const long long c = n > limit ? (n + limit - 1) / limit : 1;
    const long long v0 = vfirst[r];
    if (c == 1) {
      vptr[v0] = e0;
      vrow[v0] = (int)r;
      vacc[v0] = -1;
    } else {
      const long long len = ((n + c - 1) / c + 31) / 32 * 32;
      const int a = accfirst[r];
      long long used = 0;
      for (long long j = 0; j < c; j++) {  // (a trailing chunk can come out empty: it is skipped like an empty row)
        vptr[v0 + j] = e0 + (j * len < n ? j * len : n);
        vrow[v0 + j] = (int)r;
        vacc[v0 + j] = a;
        if (j * len < n) used++;
      }
      acc_chunks[a] = (int)used;
    }
    if (r == n_rows - 1) vptr[vfirst[n_rows]] = ptr[n_rows];
  }
}

// Entries of a CSR whose column lies in [lo, hi): first call (counts != nullptr) counts them per row,
// second call (optr = exclusive sum of the counts) copies them, order kept.
__global__ void filter_columns_kernel(const long long* __restrict__ ptr, long long n_rows, const int* __restrict__ idx,
                                      const float* __restrict__ val, int lo, int hi, long long* __restrict__ counts,
                                      const long long* __restrict__ optr, int* __restrict__ oidx,
                                      float* __restrict__ oval) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += stride) {
    if (r == n_rows) {
      if (counts) counts[r] = 0;
      continue;
    }
    long long o = counts ? 0 : optr[r];
    for (long long e = ptr[r]; e < ptr[r + 1]; e++) {
      const int c = idx[e];
      if (c < lo || c >= hi) continue;
      if (!counts) {
        oidx[o] = c;
        oval[o] = val[e];
      }
      o++;
    }
    if (counts) counts[r] = o;
  }
}

// keys[r] = entries of row r, vals[r] = r: sorted by key, descending, they give the visiting order of
// the row updates (longest rows first)
__global__ void row_length_keys_kernel(const long long* __restrict__ ptr, long long n_rows, unsigned* __restrict__ keys,
                                       int* __restrict__ vals) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    keys[r] = (unsigned)(ptr[r + 1] - ptr[r]);
    vals[r] = (int)r;
  }
}

// out[i][0..k) = F[rows[i]][0..k): selected factor rows, unpadded (als_get_rows)
__global__ void gather_rows_kernel(const float* __restrict__ F, int ks, int k, const int* __restrict__ rows,
                                   int n, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * k) return;
  out[t] = F[(long long)rows[t / k] * ks + (t % k)];
}

// ---------------------------------------------------------------------------
// Synthetic workload. Counter-based: value = f(seed,row,j), so any
// shard is reproducible without host materialisation. tests/synth_ref.py holds the
// numpy twin of these hashes.
__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}
__host__ __device__ inline unsigned long long synth_hash(unsigned long long seed,
                                                         unsigned long long row,
                                                         unsigned long long j) {
  return mix64(seed + row * 0x9E3779B97F4A7C15ULL + (j + 1ULL) * 0xD1B54A32D192ED03ULL);
}

// Each user draws nnz_per_user DISTINCT items: one uniformly from each of
// nnz_per_user equal strata of [0,n_items) (distinct and ascending by construction);
// strength uniform in {1..5}; negated with probability neg_fraction.
__global__ void synth_rows_kernel(long long row_begin, long long n_local_rows, long long n_items,
                                  int nnz_per_user, unsigned long long seed,
                                  unsigned int neg_threshold_24, long long* __restrict__ row_ptr,
                                  int* __restrict__ col_idx, float* __restrict__ val) {
  const long long total = n_local_rows * nnz_per_user;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long r = e / nnz_per_user;
    const int j = (int)(e % nnz_per_user);
    const unsigned long long h = synth_hash(seed, (unsigned long long)(row_begin + r),
                                            (unsigned long long)j);
    const long long lo = ((long long)j * n_items) / nnz_per_user;
    const long long hi = ((long long)(j + 1) * n_items) / nnz_per_user;
    col_idx[e] = (int)(lo + (long long)((h >> 32) % (unsigned long long)(hi - lo)));
    float s = (float)(1 + (int)((h & 0xffffULL) % 5ULL));
    if (((h >> 8) & 0xffffffULL) < neg_threshold_24) s = -s;
    val[e] = s;
    if (j == 0) row_ptr[r] = r * (long long)nnz_per_user;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) row_ptr[n_local_rows] = total;
}

// Sharded by-item orientation: every rank regenerates ALL users' draws (counter-based, no
// communication) but keeps only the entries whose item falls in its own item block.
// Pass 1 (counts != nullptr): counts[u] = kept entries of user u. Pass 2: write them at
// row_ptr[u].. in ascending item order (the generator's j order).
__global__ void synth_item_block_kernel(long long n_users, long long n_items, int nnz_per_user,
                                        unsigned long long seed, unsigned int neg_threshold_24,
                                        long long item_begin, long long item_end,
                                        long long* __restrict__ counts,
                                        const long long* __restrict__ row_ptr,
                                        int* __restrict__ col_idx, float* __restrict__ val) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  // strata that can intersect [item_begin, item_end)
  const int j_lo = (int)((item_begin * nnz_per_user) / n_items);
  int j_hi = (int)(((item_end - 1) * nnz_per_user) / n_items) + 2;  // +margin: filtered exactly below
  if (j_hi > nnz_per_user) j_hi = nnz_per_user;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += stride) {
    long long n = 0;
    const long long base = counts ? 0 : row_ptr[u];
    for (int j = (j_lo > 0 ? j_lo - 1 : 0); j < j_hi; j++) {
      const unsigned long long h = synth_hash(seed, (unsigned long long)u, (unsigned long long)j);
      const long long lo = ((long long)j * n_items) / nnz_per_user;
      const long long hi = ((long long)(j + 1) * n_items) / nnz_per_user;
      const long long item = lo + (long long)((h >> 32) % (unsigned long long)(hi - lo));
      if (item < item_begin || item >= item_end) continue;
      if (!counts) {
        float sgn = (float)(1 + (int)((h & 0xffffULL) % 5ULL));
        if (((h >> 8) & 0xffffffULL) < neg_threshold_24) sgn = -sgn;
        col_idx[base + n] = (int)item;
        val[base + n] = sgn;
      }
      n++;
    }
    if (counts) counts[u] = n;
  }
}

// ---- power-law workload ---------------------------------------------
// Per-user entry counts from a truncated power law (density ~ x^-2 on [1, n_max], scaled so the mean
// is ~mean_nnz), item popularity Zipf(s = 1) over a pseudo-random permutation of the items, no item
// twice per user.  Counter-based like the uniform generator: everything is a function of
// (seed, user, j); a numpy twin in the test infrastructure mirrors it.
__host__ __device__ inline double synth_unit(unsigned long long h) {  // (0, 1), 53 bits
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
// entries of user u: clamp(round(scale / (1 - U (1 - 1/n_max))), 1, min(n_max, n_items))
__host__ __device__ inline long long powerlaw_count(unsigned long long seed, unsigned long long user, double scale,
                                                    int n_max, long long n_items) {
  const double u = synth_unit(synth_hash(seed ^ 0x7c3a9f1d5b2e8461ULL, user, 0ULL));
  const double x = 1.0 / (1.0 - u * (1.0 - 1.0 / (double)n_max));  // inverse CDF, x in [1, n_max)
  long long n = (long long)(scale * x + 0.5);
  if (n < 1) n = 1;
  if (n > n_max) n = n_max;
  if (n > n_items) n = n_items;
  return n;
}
__global__ void powerlaw_counts_kernel(long long row_begin, long long n_local_rows, long long n_items, double scale,
                                       int n_max, unsigned long long seed, long long* __restrict__ counts) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= n_local_rows; r += stride)
    counts[r] = r < n_local_rows ? powerlaw_count(seed, (unsigned long long)(row_begin + r), scale, n_max, n_items) : 0;
}
// Entries of each user (one thread per user: the distinctness fix-up is a running maximum).
// Draw j of n comes from stratum j of the Zipf CDF: rank = floor((n_items + 1)^((j + U_j) / n)) - 1,
// made strictly increasing (rank_j = max(rank_j, rank_{j-1} + 1)) and kept below n_items - (n - 1 - j);
// item = (mul * rank + add) mod n_items with gcd(mul, n_items) = 1.
__global__ void powerlaw_rows_kernel(long long row_begin, long long n_local_rows, long long n_items,
                                     unsigned long long seed, unsigned int neg_threshold_24,
                                     unsigned long long perm_mul, unsigned long long perm_add,
                                     const long long* __restrict__ row_ptr, int* __restrict__ col_idx,
                                     float* __restrict__ val) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double log_n = log1p((double)n_items);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_local_rows; r += stride) {
    const long long e0 = row_ptr[r], n = row_ptr[r + 1] - e0;
    const unsigned long long user = (unsigned long long)(row_begin + r);
    long long prev = -1;
    for (long long j = 0; j < n; j++) {
      const unsigned long long h = synth_hash(seed, user, (unsigned long long)j);
      const double v = ((double)j + synth_unit(h)) / (double)n;
      long long rank = (long long)floor(expm1(v * log_n));
      if (rank <= prev) rank = prev + 1;
      const long long cap = n_items - (n - j);  // leaves room for the entries still to come
      if (rank > cap) rank = cap;
      prev = rank;
      col_idx[e0 + j] = (int)((perm_mul * (unsigned long long)rank + perm_add) % (unsigned long long)n_items);
      float s = (float)(1 + (int)((h & 0xffffULL) % 5ULL));
      if (((h >> 8) & 0xffffffULL) < neg_threshold_24) s = -s;
      val[e0 + j] = s;
    }
  }
}

// Y0 rows: k i.i.d. N(0,1) (Box-Muller on hashed uniforms) normalised to unit L2.
__global__ void synth_y0_kernel(float* __restrict__ Y, long long n_rows, int ks, int k,
                                unsigned long long seed) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  float* y = Y + r * ks;
  double total = 0.0;
  for (int f = 0; f < k; f += 2) {
    const unsigned long long h = synth_hash(seed ^ 0x5bf03635f0a5b2d1ULL, (unsigned long long)r,
                                            (unsigned long long)f);
    const float u1 = ((float)((h >> 40) & 0xffffffULL) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = (float)((h >> 8) & 0xffffffULL) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    const float g0 = rad * c, g1 = rad * s;
    y[f] = g0;
    total += (double)g0 * g0;
    if (f + 1 < k) {
      y[f + 1] = g1;
      total += (double)g1 * g1;
    }
  }
  const float norm = (float)sqrt(total);
  for (int f = 0; f < k; f++) y[f] /= norm;  // SimpleVectorMath.normalize, :78-83
  for (int f = k; f < ks; f++) y[f] = 0.f;
}

// ---------------------------------------------------------------------------
// Rare path: apparent rank of the W_u that failed, for SingularMatrixSolverException
// (CommonsMathLinearSystemSolver.java:46-54: new RRQRDecomposition(W,1e-5).getRank(0.01)).
// One thread, fp64, W_u rebuilt exactly as Worker.call does. scratch: 2*k*k + 2k + 1 doubles.
__global__ void singular_rank_kernel(const long long* row_ptr, const int* col_idx, const float* val,
                                     long long local_row, const float* M, const double* G, int ks,
                                     int k, double alpha, double lambda_alpha, int reconstruct_r,
                                     int loss_ignores_unspecified, double* scratch, int* rank_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double* W = scratch;            // row-major k x k
  double* Q = scratch + k * k;    // qrt[col][row]
  double* rDiag = Q + k * k;
  const long long e0 = row_ptr[local_row], e1 = row_ptr[local_row + 1];
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) W[i * k + j] = loss_ignores_unspecified ? 0.0 : G[i * ks + j];
  for (long long e = e0; e < e1; e++) {
    const float* y = M + (long long)col_idx[e] * ks;
    const double w = (reconstruct_r ? 0.0 : alpha * fabs((double)val[e])) +
                     (loss_ignores_unspecified ? 1.0 : 0.0);
    for (int i = 0; i < k; i++)
      for (int j = 0; j < k; j++) W[i * k + j] += ((double)y[i] * w) * (double)y[j];
  }
  for (int i = 0; i < k; i++) W[i * k + i] += lambda_alpha * (double)(e1 - e0);
  // Householder QR with column pivoting on the transposed copy (commons-math3 3.2 RRQR).
  for (int c = 0; c < k; c++)
    for (int r = 0; r < k; r++) Q[c * k + r] = W[r * k + c];
  for (int minor = 0; minor < k; minor++) {
    double best = 0.0;
    int besti = minor;
    for (int i = minor; i < k; i++) {
      double n2 = 0.0;
      for (int j = 0; j < k; j++) n2 += Q[i * k + j] * Q[i * k + j];
      if (n2 > best) { best = n2; besti = i; }
    }
    if (besti != minor)
      for (int j = 0; j < k; j++) {
        const double t = Q[minor * k + j];
        Q[minor * k + j] = Q[besti * k + j];
        Q[besti * k + j] = t;
      }
    double* qm = Q + minor * k;
    double xNormSqr = 0.0;
    for (int row = minor; row < k; row++) xNormSqr += qm[row] * qm[row];
    const double a = (qm[minor] > 0) ? -sqrt(xNormSqr) : sqrt(xNormSqr);
    rDiag[minor] = a;
    if (a != 0.0) {
      qm[minor] -= a;
      for (int col = minor + 1; col < k; col++) {
        double* qc = Q + col * k;
        double al = 0.0;
        for (int row = minor; row < k; row++) al -= qc[row] * qm[row];
        al /= a * qm[minor];
        for (int row = minor; row < k; row++) qc[row] -= al * qm[row];
      }
    }
  }
  // getRank(0.01): Frobenius norms of trailing blocks of R.
  double* sq = rDiag + k;  // sq[s] = ||R[s:,s:]||_F^2, k+1 entries
  sq[k] = 0.0;
  for (int s = k - 1; s >= 0; s--) {
    double acc = rDiag[s] * rDiag[s];
    for (int col = s + 1; col < k; col++) acc += Q[col * k + s] * Q[col * k + s];
    sq[s] = sq[s + 1] + acc;
  }
  int rank = 1;
  double lastNorm = sqrt(sq[0]);
  const double rNorm = lastNorm;
  while (rank < k) {
    const double thisNorm = sqrt(sq[rank]);
    if (thisNorm == 0 || (thisNorm / lastNorm) * rNorm < 0.01) break;
    lastNorm = thisNorm;
    rank++;
  }
  *rank_out = rank;
}

}  // namespace als
