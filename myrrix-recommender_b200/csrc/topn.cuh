// topn.cuh -- top-N scoring on the resident factors.
//
// What the reference does per query (ServerRecommender.multithreadedTopN,
// online/src/net/myrrix/online/ServerRecommender.java:443-509): walk every item vector of Y, skip
// filtered IDs, score = (float) (sum over the query's user vectors of SimpleVectorMath.dot / count)
// (RecommendIterator.java:68-110; the dot is fp32 products summed in fp64 in feature order,
// SimpleVectorMath.java:34-41), keep the best `howMany` in a priority queue ordered by
// ByValueAscComparator (TopN.java:55-75, 122-131; value, ties by item ID).
//
// Here: one pass over Y per group of up to kMaxQ queries.  HBM-bound by design -- the item rows
// are streamed exactly once per pass, as 128-row tiles moved by TMA (cp.async.bulk.tensor.2d,
// SWIZZLE_128B: thread t then reads its own row's 16-byte chunks conflict-free) through a
// two-stage mbarrier ring; every thread scores one item of the tile against all vectors of the pass
// with the reference's exact arithmetic (so scores are bit-identical), and the selection is a
// per-CTA candidate buffer per query with a running threshold:
//   key = (order-preserving bits of the score) << 32 | ~item   (unique; larger = better; equal
//         scores rank by ascending item ID, the order TopN.selectTopNFromQueue returns them in)
//   a key enters the buffer when it beats the CTA's threshold (the n-th best key of any subset
//   of the items is a lower bound of the final n-th best, so nothing that belongs to the result
//   is ever dropped); when a buffer is nearly full the CTA sorts it (bitonic, in shared memory),
//   keeps the best n and raises its threshold -- and publishes it with atomicMax so the other
//   CTAs stop collecting keys that can no longer make it.  The expected number of insertions
//   per CTA is n (1 + ln(items per CTA / n)): a handful of sorts per pass.
// A second small kernel merges the per-CTA lists of a query the same way and writes the result.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is looked up through the runtime)

#include "common.cuh"
#include "umma_common.cuh"

namespace als {
namespace topn {

constexpr int kTile = 128;     // items per tile = threads per CTA
constexpr int kThreads = 128;
constexpr int kMaxQ = 4;       // queries per pass over Y
constexpr int kMaxVec = 16;    // feature vectors per pass (all its queries together)
constexpr int kCap = 256;      // candidate keys per query per CTA (power of two)
constexpr int kMaxN = 128;     // largest howMany of the fused path (kCap - kTile >= kMaxN)
// Three stages: two tile loads (32 KB each at k = 64) in flight per CTA while a third is scored --
// with two stages the pass was bound by the latency of the one load in flight (85 us per query over
// 1 M x 64 items; HBM time 40 us).  k = 64: 96 KB ring + 12 KB, two CTAs per SM.
constexpr int kStages = 3;
static_assert(kCap - kTile >= kMaxN, "a tile's insertions must fit after a prune");

struct Params {
  const float* qbase;        // feature rows (X, or a staging buffer), row stride KS
  const int* qrow;           // [n_vec] row of each vector in qbase
  const int* vec_query;      // [n_vec] query slot of each vector, ascending; null: vector v = query v
  int n_vec, n_q;
  const unsigned* excl;      // [n_q][excl_words] bit i set: item i is filtered; null: nothing is
  long long excl_words;
  long long n_items;
  int how_many;
  unsigned long long* gthr;  // [n_q] thresholds shared by the CTAs (zero before the launch)
  unsigned long long* cand;  // [n_q][gridDim.x][how_many] per-CTA results (0 = empty)
  int* nonfinite;            // set when a score is not finite (RecommendIterator.java:99)
};

__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ unsigned long long make_key(float score, unsigned item) {
  if (score == 0.f) score = 0.f;  // -0 and +0 compare equal in the reference
  return ((unsigned long long)ordered_bits(score) << 32) | (unsigned long long)(0xffffffffu - item);
}

// descending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* a, int n, int tid, int nthreads) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += nthreads) {
        const int l = 2 * t - (t & (j - 1));
        const int r = l + j;
        const unsigned long long x = a[l], y = a[r];
        const bool desc = (l & k) == 0;
        if (desc ? (x < y) : (x > y)) {
          a[l] = y;
          a[r] = x;
        }
      }
      __syncthreads();
    }
  }
}

// Keep the best `n` of the `*cnt` keys in buf (capacity kCap); raise *thr to the n-th best when
// there are that many.  Called by the whole CTA; ends with a barrier.
__device__ __forceinline__ void prune(unsigned long long* buf, int* cnt, unsigned long long* thr, int n,
                                      int tid, unsigned long long* gthr, int nthreads = kThreads, int cap = kCap) {
  const int c = *cnt;
  __syncthreads();
  for (int i = c + tid; i < cap; i += nthreads) buf[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(buf, cap, tid, nthreads);
  if (tid == 0) {
    const int kept = c < n ? c : n;
    *cnt = kept;
    if (kept == n) {
      const unsigned long long t = buf[n - 1];
      if (t > *thr) *thr = t;
      if (gthr) atomicMax(gthr, t);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}

template <int KS>
struct Shape {
  static constexpr int kBoxCols = KS < 32 ? KS : 32;  // one box row <= 128 bytes (the swizzle span)
  static constexpr int kBoxes = KS / kBoxCols;
  static constexpr int kBoxBytes = kTile * kBoxCols * 4;
  static constexpr int kStageBytes = kBoxes * kBoxBytes;
  static constexpr bool kSwizzle = KS >= 32;
  static constexpr size_t off_vec = (size_t)kStages * kStageBytes;
  static constexpr size_t off_buf = off_vec + sizeof(float) * kMaxVec * KS;
  static constexpr size_t off_misc = off_buf + sizeof(unsigned long long) * kMaxQ * kCap;
  static constexpr size_t kTotal = off_misc + 256 + 1024;  // + slack to align the ring to 1 KB
};

// G = feature vectors scored side by side (independent fp64 chains per thread)
template <int KS, int G>
__global__ void __launch_bounds__(kThreads) topn_score_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const Params p) {
  using S = Shape<KS>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* vec = reinterpret_cast<float*>(smem + S::off_vec);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem + S::off_buf);
  unsigned long long* thr = reinterpret_cast<unsigned long long*>(smem + S::off_misc);  // [kMaxQ]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::off_misc + 64);               // [kStages]
  int* cnt = reinterpret_cast<int*>(smem + S::off_misc + 96);                          // [kMaxQ]
  int* nvq = reinterpret_cast<int*>(smem + S::off_misc + 128);                         // [kMaxQ] vectors per query
  int* vq = reinterpret_cast<int*>(smem + S::off_misc + 160);                          // [kMaxVec (+pad)] query of vector (as bytes below)
  const int tid = threadIdx.x;
  const uint32_t ring = umma::smem_u32(smem);

  // ---- setup: vectors of the pass, per-query state, barriers --------------------------------
  for (int e = tid; e < kMaxVec * KS; e += kThreads) {
    const int v = e / KS, f = e % KS;
    vec[e] = (v < p.n_vec) ? p.qbase[(long long)p.qrow[v] * KS + f] : 0.f;
  }
  if (tid < kMaxQ) {
    thr[tid] = 0ull;
    cnt[tid] = 0;
    nvq[tid] = 0;
  }
  __syncthreads();
  if (tid < kMaxVec) {
    const int q = (tid < p.n_vec) ? (p.vec_query ? p.vec_query[tid] : tid) : -1;
    reinterpret_cast<signed char*>(vq)[tid] = (signed char)q;
    if (q >= 0) atomicAdd(&nvq[q], 1);
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) umma::mbar_init(&full[s], 1);
    umma::mbar_init_fence();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const signed char* vqs = reinterpret_cast<const signed char*>(vq);

  const long long n_tiles = (p.n_items + kTile - 1) / kTile;
  auto issue = [&](int s, long long tile) {
    const uint32_t bar = umma::smem_u32(&full[s]);
    mbar_expect_tx(bar, (uint32_t)S::kStageBytes);
#pragma unroll
    for (int b = 0; b < S::kBoxes; b++)
      tma_load_2d(ring + (uint32_t)(s * S::kStageBytes + b * S::kBoxBytes), &tmap, b * S::kBoxCols,
                  (int)(tile * kTile), bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) {
      const long long tile = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (tile < n_tiles) issue(s, tile);
    }
  }

  int it = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int s = it % kStages;
    umma::mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
    const long long item = tile * kTile + tid;
    const bool valid = item < p.n_items;
    const unsigned char* st = smem + s * S::kStageBytes;
    // thresholds: this CTA's own or the best any CTA has published, whichever is higher (read per
    // thread, with the filter words below: no barrier, and the latency hides behind the scoring)
    unsigned long long gth[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; q++)
      gth[q] = q < p.n_q ? *reinterpret_cast<volatile unsigned long long*>(p.gthr + q) : 0ull;

    // filter words of my item, one per query, fetched before the scoring so their latency hides
    unsigned exw[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; q++)
      exw[q] = (p.excl && valid && q < p.n_q) ? __ldg(p.excl + (long long)q * p.excl_words + (item >> 5)) : 0u;
    double qsum = 0.0;
    int cur_q = vqs[0];
    auto finish_query = [&](int q, double sum) {
      if (!valid) return;
      unsigned w = 0u;
#pragma unroll
      for (int j = 0; j < kMaxQ; j++)
        if (j == q) w = exw[j];
      if ((w >> (item & 31)) & 1u) return;
      const int nv = nvq[q];
      const float r = (float)(nv == 1 ? sum : sum / (double)nv);  // RecommendIterator.java:98 (x / 1.0 == x)
      if (!isfinite(r)) {
        *p.nonfinite = 1;
        return;
      }
      const unsigned long long key = make_key(r, (unsigned)item);
      unsigned long long th = thr[q];
#pragma unroll
      for (int j = 0; j < kMaxQ; j++)
        if (j == q && gth[j] > th) th = gth[j];
      if (key > th) {
        const int pos = atomicAdd(&cnt[q], 1);
        buf[q * kCap + pos] = key;  // pos < kCap: at most kTile insertions since the last prune check
      }
    };
    for (int v0 = 0; v0 < p.n_vec; v0 += G) {
      double acc[G];
#pragma unroll
      for (int j = 0; j < G; j++) acc[j] = 0.0;
#pragma unroll 4
      for (int c = 0; c < KS / 4; c++) {
        // chunk c = features 4c..4c+3 of my row: box (4c) / kBoxCols, 16-byte chunk cc of its row
        const int b = (4 * c) / S::kBoxCols, cc = c % (S::kBoxCols / 4);
        const int phys = S::kSwizzle ? (cc ^ (tid & 7)) : cc;
        const float4 y = *reinterpret_cast<const float4*>(st + b * S::kBoxBytes + tid * (S::kBoxCols * 4) + phys * 16);
#pragma unroll
        for (int j = 0; j < G; j++) {
          const float4 x = *reinterpret_cast<const float4*>(vec + (v0 + j) * KS + 4 * c);  // broadcast
          acc[j] = __dadd_rn(acc[j], (double)__fmul_rn(y.x, x.x));  // SimpleVectorMath.java:38
          acc[j] = __dadd_rn(acc[j], (double)__fmul_rn(y.y, x.y));
          acc[j] = __dadd_rn(acc[j], (double)__fmul_rn(y.z, x.z));
          acc[j] = __dadd_rn(acc[j], (double)__fmul_rn(y.w, x.w));
        }
      }
#pragma unroll
      for (int j = 0; j < G; j++) {
        const int v = v0 + j;
        if (v < p.n_vec) {
          const int q = vqs[v];
          if (q != cur_q) {
            finish_query(cur_q, qsum);
            qsum = 0.0;
            cur_q = q;
          }
          qsum = __dadd_rn(qsum, acc[j]);  // RecommendIterator.java:86-89
        }
      }
    }
    finish_query(cur_q, qsum);
    __syncthreads();  // every thread has read the stage and made its insertions
    if (tid == 0) {
      const long long next = tile + (long long)kStages * gridDim.x;
      if (next < n_tiles) issue(s, next);
    }
    for (int q = 0; q < p.n_q; q++)
      if (cnt[q] > kCap - kTile) prune(buf + q * kCap, &cnt[q], &thr[q], p.how_many, tid, p.gthr + q);
  }
  // ---- this CTA's best keys of every query ---------------------------------------------------
  for (int q = 0; q < p.n_q; q++) {
    prune(buf + q * kCap, &cnt[q], &thr[q], p.how_many, tid, p.gthr + q);
    unsigned long long* dst = p.cand + ((long long)q * gridDim.x + blockIdx.x) * p.how_many;
    for (int i = tid; i < p.how_many; i += kThreads) dst[i] = (i < cnt[q]) ? buf[q * kCap + i] : 0ull;
    __syncthreads();
  }
}

// One CTA per query: the best how_many of the per-CTA lists, in result order.  gthr[q] is by now the
// largest of the CTAs' own how_many-th keys -- a lower bound of the result's last key -- so only the
// few keys at or above it are collected and sorted; every thread keeps kMergeLoads independent
// loads of the (L2-resident) lists in flight.
constexpr int kMergeThreads = 256;
constexpr int kMergeLoads = 4;
constexpr int kMergeCap = 2048;  // power of two
static_assert(kMergeThreads * kMergeLoads <= kMergeCap - kMaxN, "a round's insertions must fit after a prune");
__global__ void __launch_bounds__(kMergeThreads) topn_merge_kernel(const unsigned long long* cand, int lists,
                                                                   const unsigned long long* gthr, int how_many,
                                                                   int* out_items, float* out_values, int* out_count) {
  __shared__ unsigned long long buf[kMergeCap];
  __shared__ unsigned long long thr;
  __shared__ int cnt;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    const unsigned long long g = gthr[q];
    thr = g ? g - 1 : 0ull;  // key > thr  <=>  key >= g
    cnt = 0;
  }
  __syncthreads();
  const unsigned long long* src = cand + (long long)q * lists * how_many;
  const long long total = (long long)lists * how_many;
  constexpr int kRound = kMergeThreads * kMergeLoads;
  for (long long base = 0; base < total; base += kRound) {
    unsigned long long key[kMergeLoads];
#pragma unroll
    for (int j = 0; j < kMergeLoads; j++) {
      const long long i = base + tid + (long long)j * kMergeThreads;
      key[j] = i < total ? src[i] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < kMergeLoads; j++)
      if (key[j] > thr) buf[atomicAdd(&cnt, 1)] = key[j];
    __syncthreads();
    if (cnt > kMergeCap - kRound) prune(buf, &cnt, &thr, how_many, tid, nullptr, kMergeThreads, kMergeCap);
  }
  const int c = cnt;
  const int cap = c <= 256 ? 256 : (c <= 512 ? 512 : (c <= 1024 ? 1024 : kMergeCap));  // sort no more than needed
  prune(buf, &cnt, &thr, how_many, tid, nullptr, kMergeThreads, cap);
  for (int i = tid; i < how_many; i += kMergeThreads) {
    const bool on = i < cnt;
    out_items[(long long)q * how_many + i] = on ? (int)(0xffffffffu - (unsigned)(buf[i] & 0xffffffffull)) : -1;
    out_values[(long long)q * how_many + i] = on ? from_ordered_bits((unsigned)(buf[i] >> 32)) : 0.f;
  }
  if (tid == 0) out_count[q] = cnt;
}

// ---- filtered-item bitmaps ------------------------------------------------------------------
// OR the entries of CSR row `row` (local index) into bitmap `dst`
__global__ void mark_row_kernel(const long long* ptr, const int* idx, long long row, unsigned* dst) {
  const long long e0 = ptr[row], e1 = ptr[row + 1];
  for (long long e = e0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < e1;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = idx[e];
    atomicOr(&dst[i >> 5], 1u << (i & 31));
  }
}
// one single-user query per slot: OR row rows[q] into bitmap q (blockIdx.y = q); rows[q] < 0: none
__global__ void mark_rows_kernel(const long long* ptr, const int* idx, const int* users, long long row_begin,
                                 long long local_rows, unsigned* dst, long long words) {
  const int q = blockIdx.y;
  const long long row = (long long)users[q] - row_begin;
  if (row < 0 || row >= local_rows) return;
  const long long e0 = ptr[row], e1 = ptr[row + 1];
  for (long long e = e0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < e1;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = idx[e];
    atomicOr(&dst[q * words + (i >> 5)], 1u << (i & 31));
  }
}
__global__ void mark_list_kernel(const int* items, int n, unsigned* dst, long long words, int n_q) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = items[t];
  for (int q = 0; q < n_q; q++) atomicOr(&dst[q * words + (i >> 5)], 1u << (i & 31));
}
__global__ void and_clear_kernel(unsigned* dst, unsigned* tmp, long long words) {
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w < words) {
    dst[w] &= tmp[w];
    tmp[w] = 0u;
  }
}

}  // namespace topn
}  // namespace als
