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

constexpr int kGramThreads = 256;
constexpr int kGramChunk = 32;  // rows staged in shared memory per step

template <int KS>
struct GramShape {
  static constexpr int T = KS / 4;                 // 4x4 tiles per dimension
  static constexpr int NT = T * T;                 // tiles in the full matrix
  static constexpr int TPT = NT >= kGramThreads ? NT / kGramThreads : 1;  // tiles per thread
  static constexpr int GROUPS = NT >= kGramThreads ? 1 : kGramThreads / NT;
};

// partial: [gridDim.x * GROUPS][KS*KS] doubles.
template <int KS>
__global__ void __launch_bounds__(kGramThreads)
gramian_partial_kernel(const float* __restrict__ M, long long n_rows, double* __restrict__ partial) {
  using S = GramShape<KS>;
  __shared__ double sm[kGramChunk][KS];
  const int tid = threadIdx.x;
  const int group = (S::GROUPS > 1) ? tid / S::NT : 0;
  const int tile0 = (S::GROUPS > 1) ? tid % S::NT : tid;

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
#pragma unroll 2
    for (int r = group; r < kGramChunk; r += S::GROUPS) {
#pragma unroll
      for (int t = 0; t < S::TPT; t++) {
        const int tile = tile0 + t * kGramThreads;
        const int ti = tile / S::T, tj = tile % S::T;
        double a[4], b[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          a[e] = sm[r][4 * ti + e];
          b[e] = sm[r][4 * tj + e];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[t][4 * i + j] = fma(a[i], b[j], acc[t][4 * i + j]);
      }
    }
  }
  double* out = partial + ((size_t)blockIdx.x * S::GROUPS + group) * (size_t)(KS * KS);
#pragma unroll
  for (int t = 0; t < S::TPT; t++) {
    const int tile = tile0 + t * kGramThreads;
    const int ti = tile / S::T, tj = tile % S::T;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) out[(4 * ti + i) * KS + 4 * tj + j] = acc[t][4 * i + j];
  }
}

// G[e] = sum_p partial[p][e] in fixed order (deterministic).
__global__ void gramian_reduce_kernel(const double* __restrict__ partial, int n_partials, int kk,
                                      double* __restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kk) return;
  double s = 0.0;
  for (int p = 0; p < n_partials; p++) s += partial[(size_t)p * kk + e];
  G[e] = s;
}

}  // namespace als
