// row_update_simt.cuh -- CUDA-core row update: any feature count up to 128.
//
// One CTA per row of R (persistent grid, atomic row ticket).  Replaces the body of
// Worker.call (AlternatingLeastSquares.java:438-502):
//     W_u = G + sum_i alpha*|r_ui| * y_i y_i^T + lambda*alpha*n_u * I
//     b_u = sum_{r_ui > 0} (1 + alpha*|r_ui|) * y_i
//     x_u = W_u^{-1} b_u
// The reference solves with a pivoted Householder QR in fp64
// (CommonsMathLinearSystemSolver.java:41-45).  W_u is symmetric positive definite
// whenever lambda*alpha*n_u > 0, so an LDL^T factorisation gives the same x_u; a
// pivot <= the singularity threshold (or non-finite) reports ALS_E_SINGULAR instead
// of emitting NaN/Inf (LinearSystemSolver.java:33-34 semantics).  The rank-1 sums are
// fp32 (exact products, <= ~1e-6 relative after summation); W_u itself, the
// factorisation and the two triangular solves are fp64 like the reference's, so
// ill-conditioned rows (few entries, large G) stay inside the 1e-4 parity bar.
//
// Data movement per row: n_u gathered factor rows of 4*KS bytes (coalesced float4,
// staged in shared memory), n_u (index,value) pairs streamed, one 4*KS-byte row
// written.  Arithmetic: lower-triangular 4x4 register tiles, packed fp32x2 FMAs.
#pragma once
#include "common.cuh"
#include "solve_fp64.cuh"

namespace als {

constexpr int kMaxPeers = 15;  // other ranks of one NVSwitch domain

struct RowUpdateParams {
  const long long* row_ptr;  // CSR of this orientation, local rows [0, n_rows]
  const int* col_idx;
  const float* val;
  long long n_rows;      // local rows to process
  long long row_offset;  // global index of local row 0 in `out`
  const float* M;        // opposite factor, [n_other][KS]
  const double* G;       // M^T M, [KS*KS] fp64
  float* out;            // this factor, [n][KS]; padding columns stay 0
  int k;                 // true feature count (<= KS)
  float alpha;
  double lambda_alpha;   // lambda * alpha  (AlternatingLeastSquares.java:435)
  int reconstruct_r;
  int loss_ignores_unspecified;
  float threshold;
  int which;  // 0 = X half, 1 = Y half (for error reports)
  DeviceStatus* status;
  unsigned long long* ticket;  // zeroed before launch
  // Row-list mode (CUDA-core kernel): process only row_list[0 .. *row_list_count).
  const int* row_list;
  const int* row_list_count;
  // Visiting order of the rows (a permutation of [0, n_rows), longest rows first; nullptr: as stored):
  // with skewed row lengths (power-law data) the long rows must start first, or the CTA that meets
  // one last decides the kernel's tail.
  const int* row_order;
  // Long rows split into chunks (tensor-core kernel only; nullptr: none).  The kernel then walks
  // VIRTUAL rows (row_ptr / n_rows describe them): vrow[v] = the row a virtual row belongs to,
  // vacc[v] = -1 for a whole row, else the index of the row's accumulation record
  // (gacc[vacc * (slot floats + KS)]: the chunks' partial -D and rhs are added there; gcount counts the
  // chunks that arrived; the last one assembles W_u with G and lambda alpha n_u -- n_u from real_ptr -- and
  // solves).  acc_chunks[a] = chunks of record a.
  const int* vrow;
  const int* vacc;
  const int* acc_chunks;
  const long long* real_ptr;
  float* gacc;
  int* gcount;
  // Stash mode (tensor-core kernel only; nullptr: off).  For workloads where many rows fail the fp32
  // solve's conditioning gate because G itself is ill-conditioned (power-law data): the drain hands over
  // the data term -D alone, the solving warp keeps a copy of it (and of the rhs) in its own record of
  // `stash` ([CTA][solving warp][slot floats + KS]) before it adds G and lambda alpha n_u in fp32, and a
  // row that fails is appended to resolve_buf / resolve_rows (up to resolve_cap records; beyond that:
  // the retry list) -- resolve_fp64_kernel then solves (G + D + lambda alpha n_u I) x = b in fp64 from
  // the fp64 Gramian and the tensor-core D, without gathering the row again.
  float* stash;
  float* resolve_buf;
  int* resolve_rows;
  int* resolve_count;
  int resolve_cap;
  // Solve rows without entries too (W_u = G, b_u = 0): set for the list of rows that are keys
  // of the reference's map but whose entries were all pruned (InputFilesReader.java:202-211).
  int solve_empty;
  // Filled by the tensor-core kernel: rows it refused (singular / ill-conditioned in fp32).
  int* retry_rows;
  int* retry_count;
  // Multi-GPU: the other ranks' replicas of `out` (peer memory over NVLink).  Every finished row
  // is stored into all of them from the solve epilogue, so the factor exchange overlaps the
  // kernel instead of following it.  n_peers = 0: single GPU / exchange by NCCL.
  int n_peers;
  float* peer_out[kMaxPeers];
};

// One thread's element of a finished row -> every peer replica.
__device__ __forceinline__ void push_to_peers(const RowUpdateParams& p, long long elem_offset, float v) {
  for (int r = 0; r < p.n_peers; r++) p.peer_out[r][elem_offset] = v;
}

constexpr int kSimtThreads = 128;
constexpr int kSimtChunk = 32;  // gathered rows staged per step
constexpr int kFlushChunks = 8; // fp32 partial sums are folded into fp64 every 256 entries

template <int KS>
struct SimtShape {
  static constexpr int T = KS / 4;
  static constexpr int NTL = T * (T + 1) / 2;  // lower-triangular 4x4 tiles
  static constexpr int TPT = (NTL + kSimtThreads - 1) / kSimtThreads;
  static constexpr int LDW = KS + 1;  // padded leading dimension of W in smem
  static constexpr size_t smem_bytes() {
    return sizeof(double) * ((size_t)KS * LDW + 2 * KS) +
           sizeof(float) * ((size_t)kSimtChunk * KS + 2 * kSimtChunk) + sizeof(int) * kSimtChunk + 16;
  }
};

template <int KS>
__global__ void __launch_bounds__(kSimtThreads)
row_update_simt_kernel(const RowUpdateParams p) {
  using S = SimtShape<KS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* W = reinterpret_cast<double*>(smem_raw);  // [KS][LDW]
  double* bvec = W + KS * S::LDW;                   // [KS]
  double* invd = bvec + KS;                         // [KS]
  float* Ys = reinterpret_cast<float*>(invd + KS);  // [chunk][KS], 16B aligned
  float* wgt = Ys + kSimtChunk * KS;                // [chunk] SYRK weight
  float* cb = wgt + kSimtChunk;                     // [chunk] rhs weight
  int* idx = reinterpret_cast<int*>(cb + kSimtChunk);  // [chunk]
  __shared__ long long s_row;

  const int tid = threadIdx.x;
  const int k = p.k;

  // tile coordinates of this thread's lower-triangular tiles
  int ti[S::TPT], tj[S::TPT];
#pragma unroll
  for (int t = 0; t < S::TPT; t++) {
    const int tile = tid + t * kSimtThreads;
    int i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= tile) i++;
    while (i * (i + 1) / 2 > tile) i--;
    ti[t] = i;
    tj[t] = tile - i * (i + 1) / 2;
    if (tile >= S::NTL) { ti[t] = -1; tj[t] = 0; }
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_row = (long long)atomicAdd(p.ticket, 1ULL);
    }
    __syncthreads();
    long long row = s_row;
    if (p.row_list) {
      if (row >= (long long)*p.row_list_count) break;
      row = p.row_list[row];
    } else if (row >= p.n_rows) {
      break;
    } else if (p.row_order) {
      row = p.row_order[row];
    }
    const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
    const long long nu = e1 - e0;
    if (nu == 0 && !p.solve_empty) continue;  // not in the reference's map: leave the factor row untouched

    float2 acc[S::TPT][8];
#pragma unroll
    for (int t = 0; t < S::TPT; t++)
#pragma unroll
      for (int e = 0; e < 8; e++) acc[t][e] = make_float2(0.f, 0.f);
    float bacc = 0.f;
    double bsum = 0.0;

    // W starts as G + lambda*alpha*n_u*I in fp64 (lower triangle; ALS.java:447-450, 488-492).
    // The fp32 rank-1 partial sums are folded into it every kFlushChunks chunks, so a row
    // with thousands of entries never carries more than a few hundred terms in fp32.
    const bool add_g = !p.loss_ignores_unspecified;
    for (int e = tid; e < KS * KS; e += kSimtThreads) {
      const int r = e / KS, c = e % KS;
      if (c <= r) {
        double v = 0.0;  // padding rows/columns (k < KS) stay zero and are never factorised
        if (r < k && add_g) v = p.G[r * KS + c];
        if (r == c && r < k) v += p.lambda_alpha * (double)nu;
        W[r * S::LDW + c] = v;
      }
    }
    __syncthreads();
    auto flush_tiles = [&]() {
#pragma unroll
      for (int t = 0; t < S::TPT; t++) {
        if (ti[t] < 0) continue;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int r = 4 * ti[t] + i, c = 4 * tj[t] + j;
            const float2 v2 = acc[t][2 * i + j / 2];
            if (c <= r) W[r * S::LDW + c] += (double)((j & 1) ? v2.y : v2.x);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; e++) acc[t][e] = make_float2(0.f, 0.f);
      }
      bsum += (double)bacc;
      bacc = 0.f;
    };
    int chunks_since_flush = 0;

    for (long long c0 = e0; c0 < e1; c0 += kSimtChunk) {
      const int n = (int)min((long long)kSimtChunk, e1 - c0);
      __syncthreads();  // previous chunk fully consumed
      if (tid < n) {
        const int ci = ld_stream_i32(p.col_idx + c0 + tid);
        const float r = ld_stream_f32(p.val + c0 + tid);
        idx[tid] = ci;
        const float ar = p.alpha * fabsf(r);
        // SYRK weight: (c_u - 1) = alpha*|r| (ALS.java:471-479); +1 when the loss ignores
        // unspecified entries (partialTransposeTimesSelf, :524-539); 0 when reconstructing R.
        wgt[tid] = (p.reconstruct_r ? 0.f : ar) + (p.loss_ignores_unspecified ? 1.f : 0.f);
        // rhs weight: r (ALS.java:466-469) or c_u gated on r > 0 (:480-482)
        cb[tid] = p.reconstruct_r ? r : (r > 0.f ? 1.f + ar : 0.f);
      }
      __syncthreads();
      constexpr int V = KS / 4;
#pragma unroll 4
      for (int i = tid; i < n * V; i += kSimtThreads) {
        const int e = i / V, q = i % V;
        const float4 v = ldg_f4(p.M + (long long)idx[e] * KS + 4 * q);
        *reinterpret_cast<float4*>(Ys + e * KS + 4 * q) = v;
      }
      __syncthreads();
      for (int e = 0; e < n; e++) {
        const float w = wgt[e];
#pragma unroll
        for (int t = 0; t < S::TPT; t++) {
          if (ti[t] < 0) continue;
          const float4 a = *reinterpret_cast<const float4*>(Ys + e * KS + 4 * ti[t]);
          const float4 b = *reinterpret_cast<const float4*>(Ys + e * KS + 4 * tj[t]);
          const float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
          const float a0 = a.x * w, a1 = a.y * w, a2 = a.z * w, a3 = a.w * w;
          acc[t][0] = ffma2(make_float2(a0, a0), b01, acc[t][0]);
          acc[t][1] = ffma2(make_float2(a0, a0), b23, acc[t][1]);
          acc[t][2] = ffma2(make_float2(a1, a1), b01, acc[t][2]);
          acc[t][3] = ffma2(make_float2(a1, a1), b23, acc[t][3]);
          acc[t][4] = ffma2(make_float2(a2, a2), b01, acc[t][4]);
          acc[t][5] = ffma2(make_float2(a2, a2), b23, acc[t][5]);
          acc[t][6] = ffma2(make_float2(a3, a3), b01, acc[t][6]);
          acc[t][7] = ffma2(make_float2(a3, a3), b23, acc[t][7]);
        }
        if (tid < KS) bacc = fmaf(cb[e], Ys[e * KS + tid], bacc);
      }
      if (++chunks_since_flush == kFlushChunks) {
        flush_tiles();  // each tile entry is owned by exactly one thread: no barrier needed
        chunks_since_flush = 0;
      }
    }

    flush_tiles();  // last partial sums
    if (tid < KS) bvec[tid] = bsum;

    // fp64 LDL^T + solves (first barrier inside covers the W / bvec writes above)
    ldlt_solve_fp64<KS>(W, bvec, invd, k, tid, 0, (double)p.threshold, p.status, p.which,
                        p.row_offset + row, p.out + (p.row_offset + row) * KS);
    if (p.n_peers > 0) {
      __syncthreads();  // the row as the solve left it (untouched if it failed)
      if (tid < k) push_to_peers(p, (p.row_offset + row) * KS + tid, p.out[(p.row_offset + row) * KS + tid]);
    }
  }
}

}  // namespace als
