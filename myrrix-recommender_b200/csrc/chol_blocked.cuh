// chol_blocked.cuh -- one warp factorises and solves one k x k SPD system, k = KS in {32, 64},
// as a 16-column-blocked Cholesky (L L^T) whose trailing updates run on the tensor cores.
//
// This is the fast path of MatrixUtils.getSolver(Wu).solveDToF(b) (AlternatingLeastSquares.java
// :494 -> CommonsMathLinearSystemSolver.java:41-45): rows whose pivots indicate a singular or
// ill-conditioned W_u (max diag / min pivot > cond_limit, or a pivot <= threshold / non-finite)
// are NOT solved here; the caller re-solves them in fp64 (row_update_simt.cuh), which also
// raises ALS_E_SINGULAR exactly like the reference.
//
// Storage: the drain warps hand over N = -W_u in a "panel-major lower triangle": panel J holds
// rows 16J .. KS-1 of columns 16J .. 16J+15, 64-byte rows whose 16-byte chunks are XOR-swizzled
// by the row (WPanels::swz: 16-byte row loads, ldmatrix rows and the float2 fragment accesses are
// all bank-conflict free).  The solve runs in place; finished columns hold L (positive), the diagonal holds 1 / L[j][j].
//
// Per panel P (16 columns):
//   1. panel factorisation on the CUDA cores, rows distributed one (KS = 32) or two (KS = 64)
//      per lane: 16 column steps; the pivot and the rhs entry come out of the owner lane by
//      shuffle, 1/sqrt(d) from the SFU, the column's entries inside the 16 x 16 diagonal block
//      are published through a 16-float shared-memory vector and broadcast back with 16-byte
//      loads; packed fp32x2 FMAs update the remaining panel columns.  The forward substitution
//      z = L^-1 b rides along.
//   2. the panel (L, fp32) is written back; its rows below the diagonal block are re-read as
//      mma.sync fragments (ldmatrix; every value split into tf32 hi + tf32 lo, hi*hi + hi*lo
//      + lo*hi = fp32-grade products, ~2^-21 relative) and every trailing 16 x 16 block gets
//      N[I][J] += L[I][P] L[J][P]^T  as 12 HMMA.1688.F32.TF32 on fragments loaded from / stored
//      to the slot.  (N is the negated matrix, so the update is a plain accumulate.)
// Backward substitution x = L^-T z per panel from the last: rows below the block contribute
// sum_i L[i][c] x_i (packed FMAs + a transposing butterfly over the warp), the 16 x 16
// triangular solve inside the block runs on 16 lane pairs with the block's column read back
// from the slot.
//
// About 2.3k issued instructions per 64 x 64 solve against ~6.8k for the in-register
// right-looking sweep of round 1 (chol_warp.cuh, kept for the k = 32 tensor-core kernel).
#pragma once
#include <type_traits>

#include "chol_warp.cuh"  // StaticFor
#include "common.cuh"

namespace als {

template <int KS>
struct WPanels {
  static_assert(KS % 16 == 0, "16-column panels");
  // floats per panel row: 64 bytes, no padding; the four 16-byte chunks of a row are XOR-swizzled
  // with swz(row) so that row-per-lane 16-byte accesses, ldmatrix rows and the float2 accumulator
  // fragments are all bank-conflict free (round 2 profile of the 80-byte padded rows: every
  // fragment access took twice its wavefronts)
  static constexpr int kPS = 16;
  static constexpr int kNP = KS / 16;   // panels
  __host__ __device__ static constexpr int panel_off(int J) { return kPS * (J * KS - 8 * J * (J - 1)); }
  static constexpr int kFloats = kPS * (kNP * KS - 8 * kNP * (kNP - 1));  // KS = 64: 2560 (10 KB)
  // chunk swizzle of local row lr of a panel (depends on bits 1 and 2 of lr only, so rows lr,
  // lr + 8, lr + 16, ... share it)
  __host__ __device__ static constexpr int swz(int lr) { return (((lr >> 1) & 1) << 1) | ((lr >> 2) & 1); }
  // float offset inside a panel of (local row lr, column c in 0..15)
  __host__ __device__ static constexpr int in_panel(int lr, int c) { return lr * kPS + (c ^ (swz(lr) << 2)); }
  // float offset of element (i, c), i >= 16 * (c / 16)
  __host__ __device__ static constexpr int at(int i, int c) {
    return panel_off(c >> 4) + in_panel(i - (c & ~15), c & 15);
  }
};

__device__ __forceinline__ float rsqrt_fast(float d) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t saddr, uint32_t (&d)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
               : "r"(saddr)
               : "memory");
}
// D(16x8, fp32) += A(16x8, tf32, row) * B(8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

#ifdef ALS_SOLVE_PROF
// development builds: cycles per phase of the last solve (panel factorisation, panel write-back +
// fragment loads, trailing HMMAs, backward sweep), summed over panels
__device__ long long g_solve_prof[8];
#define ALS_SP_MARK(i) do { const long long _t = clock64(); if (lane == 0) g_solve_prof[i] += _t - sp_t; sp_t = clock64(); } while (0)
#else
#define ALS_SP_MARK(i) do { } while (0)
#endif

template <int KS>
struct CholBlocked {
  static_assert(KS == 32 || KS == 64, "blocked warp Cholesky supports k = 32 or 64");
  using WP = WPanels<KS>;
  static constexpr int kPS = WP::kPS;
  static constexpr int kNP = WP::kNP;
  static constexpr int kS = KS / 32;  // rhs / solution entries per lane: row = lane + 32 * s
  // per-warp scratch (floats): two 16-float column buffers, then the rhs / solution vector
  static constexpr int kScratch = 32 + KS;
  static constexpr unsigned FULL = 0xffffffffu;

  __device__ static __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  __device__ static __forceinline__ float get(const float2 (&A)[8], int c) {
    return (c & 1) ? A[c >> 1].y : A[c >> 1].x;
  }
  __device__ static __forceinline__ void set(float2 (&A)[8], int c, float v) {
    if (c & 1) A[c >> 1].y = v; else A[c >> 1].x = v;
  }

  // largest diagonal entry of W_u over the true rows (warp-uniform); w holds N = -W_u
  __device__ static __forceinline__ float diag_max(const float* w, int lane, int k) {
    float mine = 0.f;
#pragma unroll
    for (int s = 0; s < kS; s++) {
      const int row = lane + 32 * s;
      const float d = -w[WP::at(row, row)];
      if (row < k) mine = fmaxf(mine, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmaxf(mine, __shfl_xor_sync(FULL, mine, o));
    return mine;
  }

  // The 16 column steps of one panel.  Rows are mapped to lanes PER PANEL: lane l holds local
  // rows l and (TWO) l + 32 of the panel, i.e. matrix rows c0 + l and c0 + l + 32, so the
  // diagonal block always sits in lanes 0..15 of the first row set and the same code serves
  // every panel (the panel loop is a run-time loop: ~1.3k instructions of solver code that stay
  // in the instruction cache, instead of 3.2k of straight-line code per solve).
  // Each step looks one column ahead: the next pivot (and the next column's update) come
  // straight out of the owner lane by shuffle as soon as this column's multipliers exist, so the
  // shared-memory broadcast of the column (needed for the other 14 columns) is off the critical
  // path: SHFL -> MUFU.RSQ -> FMUL -> FFMA -> SHFL per step.
  template <bool TWO>
  __device__ static __forceinline__ void panel_steps(float* pan, int c0, float* ubuf, float* zx, int lane,
                                                     int k, float dmax, float inv_t, float cond_limit,
                                                     bool& good) {
    const int nrows = KS - c0;
    const bool act0 = lane < nrows;
    const bool act1 = TWO && (lane + 32 < nrows);
    float2 a0[8], a1[8];
    const int fz = WP::swz(lane);  // chunk swizzle of my rows (lane and lane + 32 share it)
    {
      const float4* s0 = reinterpret_cast<const float4*>(pan + lane * kPS);
      const float4* s1 = reinterpret_cast<const float4*>(pan + (lane + 32) * kPS);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act0) v = s0[q ^ fz];
        if (act1) t = s1[q ^ fz];
        a0[2 * q] = make_float2(v.x, v.y);
        a0[2 * q + 1] = make_float2(v.z, v.w);
        a1[2 * q] = make_float2(t.x, t.y);
        a1[2 * q + 1] = make_float2(t.z, t.w);
      }
    }
    float b0 = act0 ? zx[c0 + lane] : 0.f;
    float b1 = act1 ? zx[c0 + lane + 32] : 0.f;
    float myrinv = 1.f;
    float dcur = -__shfl_sync(FULL, a0[0].x, 0);  // first pivot
    StaticFor<0, 16>::run([&](auto jc) {
      constexpr int jj = decltype(jc)::value;
      const float rinv = rsqrt_fast(dcur);
      const float mm0 = -get(a0, jj) * rinv;          // L[i][j] of my first row (junk at / above the pivot)
      const float m0 = (lane > jj) ? mm0 : 0.f;       // 0 for rows at or above the pivot
      const float m1 = TWO ? -get(a1, jj) * rinv : 0.f;
      float unext = 0.f;
      if constexpr (jj < 15) {
        // look ahead: the next pivot, W[j+1][j+1] - L[j+1][j]^2, exists on lane jj+1 now
        dcur = -__shfl_sync(FULL, fmaf(mm0, mm0, get(a0, jj + 1)), jj + 1);
        unext = __shfl_sync(FULL, mm0, jj + 1);       // L[j+1][j], for the next column's update
      }
      const float zj = __shfl_sync(FULL, b0, jj) * rinv;
      if (lane == jj) {
        myrinv = rinv;
        b0 = zj;
        set(a0, jj, rinv);  // the diagonal keeps 1 / L[j][j]
      } else {
        b0 = fmaf(-m0, zj, b0);
        set(a0, jj, m0);
      }
      if (TWO) {
        b1 = fmaf(-m1, zj, b1);
        set(a1, jj, m1);
      }
      if constexpr (jj < 15) {
        set(a0, jj + 1, fmaf(m0, unext, get(a0, jj + 1)));
        if (TWO) set(a1, jj + 1, fmaf(m1, unext, get(a1, jj + 1)));
      }
      if constexpr (jj < 14) {
        // columns jj+2 .. 15: the diagonal block's part of column j through shared memory
        float* buf = ubuf + (jj & 1) * 16;
        if (lane < 16) buf[lane] = m0;
        __syncwarp();
#pragma unroll
        for (int q = (jj + 2) / 4; q < 4; q++) {
          const float4 u = *reinterpret_cast<const float4*>(buf + 4 * q);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int pr = 2 * q + h;  // column pair (2 pr, 2 pr + 1)
            const float2 uu = h ? make_float2(u.z, u.w) : make_float2(u.x, u.y);
            if (2 * pr + 1 < jj + 2) continue;
            if (2 * pr + 1 == jj + 2) {  // only the pair's second column is still open
              a0[pr].y = fmaf(m0, uu.y, a0[pr].y);
              if (TWO) a1[pr].y = fmaf(m1, uu.y, a1[pr].y);
            } else {
              a0[pr] = ffma2(make_float2(m0, m0), uu, a0[pr]);
              if (TWO) a1[pr] = ffma2(make_float2(m1, m1), uu, a1[pr]);
            }
          }
        }
      }
    });
    // finished panel back to the slot (L; 1 / L[j][j] on the diagonal), rhs back to the vector
    {
      float4* d0 = reinterpret_cast<float4*>(pan + lane * kPS);
      float4* d1 = reinterpret_cast<float4*>(pan + (lane + 32) * kPS);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (act0) d0[q ^ fz] = make_float4(a0[2 * q].x, a0[2 * q].y, a0[2 * q + 1].x, a0[2 * q + 1].y);
        if (act1) d1[q ^ fz] = make_float4(a1[2 * q].x, a1[2 * q].y, a1[2 * q + 1].x, a1[2 * q + 1].y);
      }
      if (act0) zx[c0 + lane] = b0;
      if (act1) zx[c0 + lane + 32] = b1;
    }
    // Judge the pivots of the true rows through 1/sqrt(d): d > threshold <=> rinv^2 < 1/threshold,
    // dmax / d <= cond_limit <=> dmax * rinv^2 <= cond_limit.  Zero / negative / NaN pivots give
    // inf / NaN and fail.  (No early exit: the warp stays converged.)
    if (lane < 16 && c0 + lane < k) {
      const float r2 = myrinv * myrinv;
      good = good && myrinv > 0.f && r2 < inv_t && r2 * dmax <= cond_limit;
    }
    __syncwarp();
  }

  // hi/lo tf32 split of one 16-row block of the panel (two k-steps of 8 columns)
  // ldmatrix: lane i addresses row r = (i&7) + 8*((i>>3)&1) of the 16-row block and the 16-byte
  // chunk 2 ks + (i>>4) of the row; ldm_row = r * kPS, ldm_chunk = (i>>4) ^ swz(r)
  __device__ static __forceinline__ void load_frags(const float* blk_rows, int ldm_row, int ldm_chunk,
                                                    uint32_t (&hi)[2][4], uint32_t (&lo)[2][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      uint32_t raw[4];
      ldmatrix_x4(smem_addr(blk_rows + ldm_row + ((ldm_chunk ^ (2 * ks)) << 2)), raw);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        hi[ks][e] = raw[e] & 0xffffe000u;
        lo[ks][e] = __float_as_uint(__uint_as_float(raw[e]) - __uint_as_float(hi[ks][e]));
      }
    }
  }

  // w: the slot (N = -W_u on entry).  scratch: kScratch floats private to this warp.
  // b[s]: rhs entry of row lane + 32 s on entry, solution entry on return.  dmax: diag_max()
  // of the slot before the sweep.  k: true feature count; padding rows (>= k) must carry a
  // unit diagonal and zero rhs and are not judged.
  // Returns (warp-uniform) whether the system was solved to the fast path's standards.
  __device__ static __forceinline__ bool factor_solve(float* w, float* scratch, float (&b)[kS],
                                                      float dmax, float threshold, float cond_limit,
                                                      int lane, int k) {
    float* ubuf = scratch;
    float* zx = scratch + 32;  // rhs -> z = L^-1 b -> x, by matrix row
    const int g = lane >> 2, t = lane & 3;
    const int ldm_r = (lane & 7) + 8 * ((lane >> 3) & 1);
    const int ldm_row = ldm_r * kPS, ldm_chunk = (lane >> 4) ^ WP::swz(ldm_r);
    // accumulator fragment of n-tile nt: rows g / g+8 (same swizzle), columns 8 nt + 2t, +1:
    // chunk (2 nt + (t>>1)) ^ swz(g), float (t&1)*2 inside it
    const int cfr_base = g * kPS + (t & 1) * 2, cfr_chunk = (t >> 1) ^ WP::swz(g);
#pragma unroll
    for (int s = 0; s < kS; s++) zx[lane + 32 * s] = b[s];
    __syncwarp();
    bool good = dmax > threshold && isfinite(dmax);
    const float inv_t = 1.0f / threshold;
#ifdef ALS_SOLVE_PROF
    long long sp_t = clock64();
#endif

    // ---- forward: panel factorisation + trailing update --------------------------------------
#pragma unroll 1
    for (int P = 0; P < kNP; P++) {
      const int c0 = 16 * P;
      float* pan = w + WP::panel_off(P);
      if (c0 + 32 < KS) panel_steps<true>(pan, c0, ubuf, zx, lane, k, dmax, inv_t, cond_limit, good);
      else panel_steps<false>(pan, c0, ubuf, zx, lane, k, dmax, inv_t, cond_limit, good);
      ALS_SP_MARK(1);  // panel: load, 16 column steps, write-back
      // trailing update on the tensor cores: N[I][J] += L[I][P] L[J][P]^T, P < J <= I; every
      // product as hi*hi + hi*lo + lo*hi of tf32 splits, in three independent accumulators
      const int NB = kNP - 1 - P;  // block rows below the diagonal block
#pragma unroll 1
      for (int bj = 0; bj < NB; bj++) {
        uint32_t hiB[2][4], loB[2][4];
        load_frags(pan + 16 * (bj + 1) * kPS, ldm_row, ldm_chunk, hiB, loB);
        float* panJ = w + WP::panel_off(P + 1 + bj);
#pragma unroll
        for (int dd = 0; dd < kNP - 1; dd++) {  // (unrolled: the next block's loads overlap this block's HMMAs)
          const int bi = bj + dd;
          if (bi >= NB) break;
          uint32_t hiA[2][4], loA[2][4];
          if (bi == bj) {
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
#pragma unroll
              for (int e = 0; e < 4; e++) { hiA[ks][e] = hiB[ks][e]; loA[ks][e] = loB[ks][e]; }
          } else {
            load_frags(pan + 16 * (bi + 1) * kPS, ldm_row, ldm_chunk, hiA, loA);
          }
          float* blk = panJ + 16 * (bi - bj) * kPS + cfr_base;
          float chh[2][4], chl[2][4], clh[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; nt++) {
            const float2 v0 = *reinterpret_cast<const float2*>(blk + ((cfr_chunk ^ (2 * nt)) << 2));
            const float2 v1 = *reinterpret_cast<const float2*>(blk + 8 * kPS + ((cfr_chunk ^ (2 * nt)) << 2));
            chh[nt][0] = v0.x; chh[nt][1] = v0.y; chh[nt][2] = v1.x; chh[nt][3] = v1.y;
#pragma unroll
            for (int e = 0; e < 4; e++) { chl[nt][e] = 0.f; clh[nt][e] = 0.f; }
          }
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
              // B fragment of n-tile nt = A-fragment registers (nt, nt + 2) of block row bj
              mma_tf32(chh[nt], hiA[ks][0], hiA[ks][1], hiA[ks][2], hiA[ks][3], hiB[ks][nt], hiB[ks][nt + 2]);
              mma_tf32(chl[nt], hiA[ks][0], hiA[ks][1], hiA[ks][2], hiA[ks][3], loB[ks][nt], loB[ks][nt + 2]);
              mma_tf32(clh[nt], loA[ks][0], loA[ks][1], loA[ks][2], loA[ks][3], hiB[ks][nt], hiB[ks][nt + 2]);
            }
          }
#pragma unroll
          for (int nt = 0; nt < 2; nt++) {
            *reinterpret_cast<float2*>(blk + ((cfr_chunk ^ (2 * nt)) << 2)) =
                make_float2(chh[nt][0] + (chl[nt][0] + clh[nt][0]), chh[nt][1] + (chl[nt][1] + clh[nt][1]));
            *reinterpret_cast<float2*>(blk + 8 * kPS + ((cfr_chunk ^ (2 * nt)) << 2)) =
                make_float2(chh[nt][2] + (chl[nt][2] + clh[nt][2]), chh[nt][3] + (chl[nt][3] + clh[nt][3]));
          }
        }
      }
      __syncwarp();
      ALS_SP_MARK(3);  // trailing blocks
    }

    // ---- backward: x = L^-T z, panel by panel from the last -----------------------------------
    const int tt = (lane >> 1) & 15;  // column of the block this lane (pair) finishes
#pragma unroll 1
    for (int P = kNP - 1; P >= 0; P--) {
      const int c0 = 16 * P;
      const float* pan = w + WP::panel_off(P);
      const int nrows = KS - c0;
      float sum = 0.f;  // sum_{i >= c0+16} L[i][c0+tt] x_i
      if (P + 1 < kNP) {
        float2 acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kS; s++) {
          const int lr = lane + 32 * s;  // local row of the panel
          const bool on = lr >= 16 && lr < nrows;
          const float xm = on ? zx[c0 + lr] : 0.f;
          const float2 xx = make_float2(xm, xm);
          const float4* src = reinterpret_cast<const float4*>(pan + lr * kPS);
          const int fz = WP::swz(lane);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) v = src[q ^ fz];
            acc[2 * q] = ffma2(make_float2(v.x, v.y), xx, acc[2 * q]);
            acc[2 * q + 1] = ffma2(make_float2(v.z, v.w), xx, acc[2 * q + 1]);
          }
        }
        // transposing butterfly: 16 sums over 32 lanes; lane ends with the sum for column tt
        const bool b16 = (lane & 16) != 0, b8 = (lane & 8) != 0, b4 = (lane & 4) != 0, b2 = (lane & 2) != 0;
        float r8[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float lo_v = (e & 1) ? acc[e >> 1].y : acc[e >> 1].x;            // column e
          const float hi_v = (e & 1) ? acc[4 + (e >> 1)].y : acc[4 + (e >> 1)].x;  // column e + 8
          r8[e] = (b16 ? hi_v : lo_v) + __shfl_xor_sync(FULL, b16 ? lo_v : hi_v, 16);
        }
        float r4[4];
#pragma unroll
        for (int e = 0; e < 4; e++) r4[e] = (b8 ? r8[e + 4] : r8[e]) + __shfl_xor_sync(FULL, b8 ? r8[e] : r8[e + 4], 8);
        float r2[2];
#pragma unroll
        for (int e = 0; e < 2; e++) r2[e] = (b4 ? r4[e + 2] : r4[e]) + __shfl_xor_sync(FULL, b4 ? r4[e] : r4[e + 2], 4);
        float r1 = (b2 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b2 ? r2[0] : r2[1], 2);
        r1 += __shfl_xor_sync(FULL, r1, 1);
        sum = r1;
      }
      ALS_SP_MARK(4);  // rows below the block + butterfly
      // x_t = dinv_t (z_t - sum_t) - sum_{t' > t} (dinv_t L[t'][t]) x_t': column tt of the diagonal
      // block pre-scaled by 1 / L[tt][tt], so each of the 16 steps is one shuffle + one FMA
      const float dinv = pan[WP::in_panel(tt, tt)];
      float v = (zx[c0 + tt] - sum) * dinv;
      float col[16];
#pragma unroll
      for (int tp = 1; tp < 16; tp++) col[tp] = pan[WP::in_panel(tp, tt)] * dinv;
      float xmine = 0.f;
#pragma unroll
      for (int tp = 15; tp >= 0; tp--) {
        const float xt = __shfl_sync(FULL, v, 2 * tp);  // x_{c0+tp}, final on its lane pair
        if (tt == tp) xmine = xt;
        if (tp > 0) v = fmaf(-col[tp], xt, v);  // lanes tt >= tp read junk here; their x is already taken
      }
      __syncwarp();  // every lane has read z of this block
      if ((lane & 1) == 0) zx[c0 + tt] = xmine;
      __syncwarp();
      ALS_SP_MARK(5);  // in-block triangular solve
    }
    bool fin = true;
#pragma unroll
    for (int s = 0; s < kS; s++) {
      b[s] = zx[lane + 32 * s];
      fin = fin && isfinite(b[s]);
    }
    return __all_sync(FULL, good && fin);
  }
};

}  // namespace als
