// gramian.cuh -- G = M^T M in fp64 for a tall-skinny fp32 factor matrix.
//
// Replaces MatrixUtils.transposeTimesSelf (common/.../math/MatrixUtils.java:219-239),
// which the reference runs single-threaded before every half-iteration
// (AlternatingLeastSquares.java:342, 369).  The reference rounds each product to fp32
// and accumulates in fp64; here each element is widened once and the product is an
// exact fp64 FMA -- the two differ by <= 2^-24 relative per term, far inside the
// 1e-4 parity bar, and the sum itself is fp64 in both.
//
// HBM-bound by design: one coalesced pass over M (4*KS bytes per row), partial
// sums per CTA written once, then a fixed-order second pass so the result is
// bit-reproducible run to run.
#pragma once
#include "common.cuh"

namespace als {

constexpr int kGramThreads = 288;
constexpr int kGramChunk = 32;  // rows staged in shared memory per step

// G is symmetric: only the 4x4 tiles on or below the diagonal are accumulated (T (T + 1) / 2 of T^2:
// 136 of 256 at k = 64 -- the kernel is DFMA-bound, so this is nearly half its time) and the second
// pass mirrors them.  The tiles are dealt to the threads in GROUPS row-interleaved groups.
template <int KS>
struct GramShape {
  static constexpr int T = KS / 4;                 // 4x4 tiles per dimension
  static constexpr int NT = T * (T + 1) / 2;       // tiles on or below the diagonal
  static constexpr int TPT = NT >= kGramThreads ? (NT + kGramThreads - 1) / kGramThreads : 1;  // tiles per thread
  static constexpr int GROUPS = NT >= kGramThreads ? 1 : kGramThreads / NT;
  static constexpr int kActive = NT >= kGramThreads ? kGramThreads : GROUPS * NT;
  static constexpr bool kFold = GROUPS > 2;              // groups folded inside the CTA (k <= 32)
  static constexpr int kPartialsPerCta = kFold ? 1 : GROUPS;
};

// lower-triangular tile number -> (ti, tj), ti >= tj
__device__ __forceinline__ void gram_tile(int tile, int& ti, int& tj) {
  int i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
  while ((i + 1) * (i + 2) / 2 <= tile) i++;
  while (i * (i + 1) / 2 > tile) i--;
  ti = i;
  tj = tile - i * (i + 1) / 2;
}

// partial: [gridDim.x * GROUPS][KS*KS] doubles (entries of the tiles on or below the diagonal).
template <int KS>
__global__ void __launch_bounds__(kGramThreads)
gramian_partial_kernel(const float* __restrict__ M, long long n_rows, double* __restrict__ partial) {
  using S = GramShape<KS>;
  __shared__ double sm[kGramChunk][KS];
  const int tid = threadIdx.x;
  const bool active = tid < S::kActive;
  const int group = (S::GROUPS > 1) ? tid / S::NT : 0;
  const int tile0 = (S::GROUPS > 1) ? tid % S::NT : tid;
  int ti[S::TPT], tj[S::TPT];
#pragma unroll
  for (int t = 0; t < S::TPT; t++) {
    const int tile = tile0 + t * kGramThreads;
    ti[t] = -1;
    tj[t] = 0;
    if (active && tile < S::NT) gram_tile(tile, ti[t], tj[t]);
  }

  double acc[S::TPT][16];
#pragma unroll
  for (int t = 0; t < S::TPT; t++)
#pragma unroll
    for (int e = 0; e < 16; e++) acc[t][e] = 0.0;

  constexpr int V = KS / 4;  // float4 per row
  for (long long base = (long long)blockIdx.x * kGramChunk; base < n_rows;
       base += (long long)gridDim.x * kGramChunk) {
    const int rows = (int)min((long long)kGramChunk, n_rows - base);
    __syncthreads();
    for (int i = tid; i < kGramChunk * V; i += kGramThreads) {
      const int r = i / V, q = i % V;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = ldg_f4(M + (base + r) * KS + 4 * q);
      sm[r][4 * q + 0] = (double)v.x;
      sm[r][4 * q + 1] = (double)v.y;
      sm[r][4 * q + 2] = (double)v.z;
      sm[r][4 * q + 3] = (double)v.w;
    }
    __syncthreads();
    if (active) {
#pragma unroll 2
      for (int r = group; r < kGramChunk; r += S::GROUPS) {
#pragma unroll
        for (int t = 0; t < S::TPT; t++) {
          if (ti[t] < 0) continue;
          double a[4], b[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            a[e] = sm[r][4 * ti[t] + e];
            b[e] = sm[r][4 * tj[t] + e];
          }
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[t][4 * i + j] = fma(a[i], b[j], acc[t][4 * i + j]);
        }
      }
    }
  }
  if constexpr (S::kFold) {
    // many small groups (k <= 32): fold them inside the CTA, in group order, so that the second pass
    // sees one partial per CTA
    __syncthreads();
    double* fold = &sm[0][0];  // (KS * KS doubles fit: KS <= 32)
    for (int e = tid; e < KS * KS; e += kGramThreads) fold[e] = 0.0;
    for (int g = 0; g < S::GROUPS; g++) {
      __syncthreads();
      if (active && group == g) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) fold[(4 * ti[0] + i) * KS + 4 * tj[0] + j] += acc[0][4 * i + j];
      }
    }
    __syncthreads();
    double* out = partial + (size_t)blockIdx.x * (size_t)(KS * KS);
    for (int e = tid; e < KS * KS; e += kGramThreads) out[e] = fold[e];
    return;
  }
  if (!active) return;
  double* out = partial + ((size_t)blockIdx.x * S::GROUPS + group) * (size_t)(KS * KS);
#pragma unroll
  for (int t = 0; t < S::TPT; t++) {
    if (ti[t] < 0) continue;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) out[(4 * ti[t] + i) * KS + 4 * tj[t] + j] = acc[t][4 * i + j];
  }
}

// G[i][j] = sum_p partial[p][i][j] in a fixed order (deterministic) for the tiles on or below the diagonal;
// the tiles above it are their mirror images (a[i] * b[j] and b[j] * a[i] are the same product).
// 64 entries x 4 interleaved slices of the partials per CTA: the loads of a slice are independent, the
// slices are combined in slice order.
constexpr int kGramReduceSlices = 4;
__global__ void __launch_bounds__(64 * kGramReduceSlices)
gramian_reduce_kernel(const double* __restrict__ partial, int n_partials, int ks, double* __restrict__ G) {
  __shared__ double part[kGramReduceSlices][64];
  const int le = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + le;
  const int kk = ks * ks;
  const int i = e < kk ? e / ks : 0, j = e < kk ? e % ks : 0;
  const bool mine = e < kk && (i >> 2) >= (j >> 2);  // (the rest is written by the thread of (j, i))
  double s = 0.0;
  if (mine) {
#pragma unroll 4
    for (int p = slice; p < n_partials; p += kGramReduceSlices) s += partial[(size_t)p * kk + e];
  }
  part[slice][le] = s;
  __syncthreads();
  if (slice == 0 && mine) {
    double t = part[0][le];
#pragma unroll
    for (int q = 1; q < kGramReduceSlices; q++) t += part[q][le];
    G[e] = t;
    if ((i >> 2) > (j >> 2)) G[j * ks + i] = t;
  }
}

}  // namespace als
