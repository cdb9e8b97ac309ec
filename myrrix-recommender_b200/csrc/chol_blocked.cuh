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
// rows 16J .. KS-1 of columns 16J .. 16J+15, row stride kPS floats (80 bytes: 16-byte row loads,
// ldmatrix rows and the float2 fragment accesses are all (nearly) conflict-free).  The solve
// runs in place; finished columns hold L (positive), the diagonal holds 1 / L[j][j].
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
  static constexpr int kPS = 20;        // floats per panel row (16 + 4 pad)
  static constexpr int kNP = KS / 16;   // panels
  __host__ __device__ static constexpr int panel_off(int J) { return kPS * (J * KS - 8 * J * (J - 1)); }
  static constexpr int kFloats = kPS * (kNP * KS - 8 * kNP * (kNP - 1));  // KS = 64: 3200 (12.8 KB)
  // float offset of element (i, c), i >= 16 * (c / 16)
  __host__ __device__ static constexpr int at(int i, int c) {
    return panel_off(c >> 4) + (i - (c & ~15)) * kPS + (c & 15);
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

template <int KS>
struct CholBlocked {
  static_assert(KS == 32 || KS == 64, "blocked warp Cholesky supports k = 32 or 64");
  using WP = WPanels<KS>;
  static constexpr int kPS = WP::kPS;
  static constexpr int kNP = WP::kNP;
  static constexpr int kS = KS / 32;  // rows per lane: row = lane + 32 * s
  static constexpr int kScratch = 32; // floats of per-warp scratch: two 16-float column buffers
  static constexpr unsigned FULL = 0xffffffffu;

  __device__ static __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
  }

  // largest diagonal entry of W_u over the true rows (warp-uniform); w holds N = -W_u
  __device__ static __forceinline__ float diag_max(const float* w, int lane, int k) {
    float mine = 0.f;
#pragma unroll
    for (int s = 0; s < kS; s++) {
      const int row = lane + 32 * s;
      const float d = -w[WP::panel_off(row >> 4) + (row & 15) * (kPS + 1)];
      if (row < k) mine = fmaxf(mine, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = fmaxf(mine, __shfl_xor_sync(FULL, mine, o));
    return mine;
  }

  // w: the slot (N = -W_u on entry).  scratch: kScratch floats private to this warp.
  // b[s]: rhs entry of row lane + 32 s on entry, solution entry on return.  dmax: diag_max()
  // of the slot before the sweep.  k: true feature count; padding rows (>= k) must carry a
  // unit diagonal and zero rhs and are not judged.
  // Returns (warp-uniform) whether the system was solved to the fast path's standards.
  __device__ static __forceinline__ bool factor_solve(float* w, float* scratch, float (&b)[kS],
                                                      float dmax, float threshold, float cond_limit,
                                                      int lane, int k) {
    float rinv_mine[kS];  // 1/sqrt(pivot) of my rows
#pragma unroll
    for (int s = 0; s < kS; s++) rinv_mine[s] = 0.f;
    const int g = lane >> 2, t = lane & 3;
    // ldmatrix: lane i addresses row (i&7) + 8*((i>>3)&1) of a 16-row block, 16-byte column
    // group (i>>4) of an 8-float k-step
    const int ldm_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * kPS + 4 * (lane >> 4);
    const int cfr_off = g * kPS + 2 * t;  // accumulator fragment: rows g / g+8, columns 2t, 2t+1 (+8)

    // ---- forward: panel factorisation + trailing update --------------------------------------
    StaticFor<0, kNP>::run([&](auto pc) {
      constexpr int P = decltype(pc)::value;
      constexpr int c0 = 16 * P;
      constexpr int s_min = c0 / 32;  // first slot with active rows
      float* pan = w + WP::panel_off(P);
      float2 a[kS][8];
      bool act[kS];
#pragma unroll
      for (int s = s_min; s < kS; s++) {
        const int row = lane + 32 * s;
        act[s] = row >= c0;
        const float4* src = reinterpret_cast<const float4*>(pan + (row - c0) * kPS);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (act[s]) v = src[q];
          a[s][2 * q] = make_float2(v.x, v.y);
          a[s][2 * q + 1] = make_float2(v.z, v.w);
        }
      }
      StaticFor<0, 16>::run([&](auto jc) {
        constexpr int jj = decltype(jc)::value;
        constexpr int j = c0 + jj;
        constexpr int own = j & 31, os = j >> 5;
        float* buf = scratch + (jj & 1) * 16;
        const float njj = (jj & 1) ? a[os][jj >> 1].y : a[os][jj >> 1].x;  // -W[.][j] of my row in slot os
        const float d = -__shfl_sync(FULL, njj, own);                      // pivot
        const float rinv = rsqrt_fast(d);
        const float zj = __shfl_sync(FULL, b[os], own) * rinv;
        float m[kS];
#pragma unroll
        for (int s = s_min; s < kS; s++) {
          const float nv = (jj & 1) ? a[s][jj >> 1].y : a[s][jj >> 1].x;
          const bool below = (s > os) || (lane > own);
          m[s] = below ? -nv * rinv : 0.f;  // L[i][j]; 0 for rows at or above the pivot
          float keep = m[s];
          if (s == os && lane == own) {
            keep = rinv;  // the diagonal keeps 1 / L[j][j]
            rinv_mine[s] = rinv;
            b[s] = zj;
          } else {
            b[s] = fmaf(-m[s], zj, b[s]);
          }
          if (jj & 1) a[s][jj >> 1].y = keep; else a[s][jj >> 1].x = keep;
        }
        if constexpr (jj < 15) {
          // the diagonal block's part of column j, for everybody
          if ((lane >> 4) == ((c0 & 31) >> 4)) buf[lane & 15] = m[os];
          __syncwarp();
#pragma unroll
          for (int q = (jj + 1) / 4; q < 4; q++) {
            const float4 u = *reinterpret_cast<const float4*>(buf + 4 * q);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int pr = 2 * q + h;  // column pair (2 pr, 2 pr + 1) of the panel
              const float2 uu = h ? make_float2(u.z, u.w) : make_float2(u.x, u.y);
              if (2 * pr + 1 <= jj) continue;
#pragma unroll
              for (int s = s_min; s < kS; s++) {
                if (2 * pr == jj) a[s][pr].y = fmaf(m[s], uu.y, a[s][pr].y);  // .x is column j itself
                else a[s][pr] = ffma2(make_float2(m[s], m[s]), uu, a[s][pr]);
              }
            }
          }
        }
      });
      // write the finished panel back (L; 1/L[j][j] on the diagonal)
#pragma unroll
      for (int s = s_min; s < kS; s++) {
        const int row = lane + 32 * s;
        float4* dst = reinterpret_cast<float4*>(pan + (row - c0) * kPS);
        if (act[s]) {
#pragma unroll
          for (int q = 0; q < 4; q++)
            dst[q] = make_float4(a[s][2 * q].x, a[s][2 * q].y, a[s][2 * q + 1].x, a[s][2 * q + 1].y);
        }
      }
      __syncwarp();
      if constexpr (P + 1 < kNP) {
        // trailing update on the tensor cores: N[I][J] += L[I][P] L[J][P]^T, P < J <= I
        constexpr int NB = kNP - 1 - P;  // block rows below the diagonal block
        uint32_t hi[NB][2][4], lo[NB][2][4];
#pragma unroll
        for (int bi = 0; bi < NB; bi++) {
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {
            uint32_t raw[4];
            ldmatrix_x4(smem_addr(pan + 16 * (bi + 1) * kPS + 8 * ks + ldm_off), raw);
#pragma unroll
            for (int e = 0; e < 4; e++) {
              hi[bi][ks][e] = raw[e] & 0xffffe000u;
              lo[bi][ks][e] = __float_as_uint(__uint_as_float(raw[e]) - __uint_as_float(hi[bi][ks][e]));
            }
          }
        }
#pragma unroll
        for (int bj = 0; bj < NB; bj++) {
          float* panJ = w + WP::panel_off(P + 1 + bj);
#pragma unroll
          for (int bi = bj; bi < NB; bi++) {
            float* blk = panJ + 16 * (bi - bj) * kPS + cfr_off;
            float c[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
              const float2 v0 = *reinterpret_cast<const float2*>(blk + 8 * nt);
              const float2 v1 = *reinterpret_cast<const float2*>(blk + 8 * kPS + 8 * nt);
              c[nt][0] = v0.x; c[nt][1] = v0.y; c[nt][2] = v1.x; c[nt][3] = v1.y;
            }
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                // B fragment of n-tile nt = A-fragment registers (nt, nt + 2) of block row bj
                mma_tf32(c[nt], hi[bi][ks][0], hi[bi][ks][1], hi[bi][ks][2], hi[bi][ks][3],
                         hi[bj][ks][nt], hi[bj][ks][nt + 2]);
                mma_tf32(c[nt], hi[bi][ks][0], hi[bi][ks][1], hi[bi][ks][2], hi[bi][ks][3],
                         lo[bj][ks][nt], lo[bj][ks][nt + 2]);
                mma_tf32(c[nt], lo[bi][ks][0], lo[bi][ks][1], lo[bi][ks][2], lo[bi][ks][3],
                         hi[bj][ks][nt], hi[bj][ks][nt + 2]);
              }
            }
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
              *reinterpret_cast<float2*>(blk + 8 * nt) = make_float2(c[nt][0], c[nt][1]);
              *reinterpret_cast<float2*>(blk + 8 * kPS + 8 * nt) = make_float2(c[nt][2], c[nt][3]);
            }
          }
        }
        __syncwarp();
      }
    });

    // Judge the pivots of the true rows through 1/sqrt(d): d > threshold <=> rinv^2 < 1/threshold,
    // dmax / d <= cond_limit <=> dmax * rinv^2 <= cond_limit.  Zero / negative / NaN pivots give
    // inf / NaN and fail.  No early exit: the warp stays converged through the backward sweep.
    bool good = dmax > threshold && isfinite(dmax);
    {
      const float inv_t = 1.0f / threshold;
#pragma unroll
      for (int s = 0; s < kS; s++) {
        const float r2 = rinv_mine[s] * rinv_mine[s];
        if (lane + 32 * s < k) good = good && rinv_mine[s] > 0.f && r2 < inv_t && r2 * dmax <= cond_limit;
      }
    }

    // ---- backward: x = L^-T z, panel by panel from the last -----------------------------------
    float x[kS];
#pragma unroll
    for (int s = 0; s < kS; s++) x[s] = 0.f;
    const int tt = (lane >> 1) & 15;  // column of the block this lane (pair) finishes
    StaticFor<kNP - 1, -1, -1>::run([&](auto pc) {
      constexpr int P = decltype(pc)::value;
      constexpr int c0 = 16 * P;
      const float* pan = w + WP::panel_off(P);
      float sum = 0.f;  // sum_{i >= c0+16} L[i][c0+tt] x_i
      if constexpr (P + 1 < kNP) {
        constexpr int s_lo = (c0 + 16) / 32;
        float2 acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = s_lo; s < kS; s++) {
          const int row = lane + 32 * s;
          const bool on = row >= c0 + 16;
          const float xm = on ? x[s] : 0.f;
          const float2 xx = make_float2(xm, xm);
          const float4* src = reinterpret_cast<const float4*>(pan + (row - c0) * kPS);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) v = src[q];
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
      // z of column c0 + tt sits in the lane that owns that row
      constexpr int zs = c0 >> 5;
      float v = __shfl_sync(FULL, b[zs], (c0 & 31) + tt) - sum;
      // column tt of the diagonal block (rows tt+1..15) and 1/L[tt][tt]
      float col[16];
#pragma unroll
      for (int tp = 1; tp < 16; tp++) col[tp] = pan[tp * kPS + tt];
      const float dinv = pan[tt * kPS + tt];
      float xmine = 0.f;
#pragma unroll
      for (int tp = 15; tp >= 0; tp--) {
        const float xt = __shfl_sync(FULL, v * dinv, 2 * tp);  // x_{c0+tp}, final on its lane pair
        if (tt == tp) xmine = xt;
        if (tp > 0) v = fmaf(-col[tp], xt, v);  // lanes tt >= tp read junk here; their x is already taken
      }
      // hand the block's solution to the row owners
      const int src = 2 * ((lane - (c0 & 31)) & 15);
      const float xo = __shfl_sync(FULL, xmine, src);
      if ((lane >> 4) == ((c0 & 31) >> 4)) x[zs] = xo;
    });
    bool fin = true;
#pragma unroll
    for (int s = 0; s < kS; s++) {
      fin = fin && isfinite(x[s]);
      b[s] = x[s];
    }
    return __all_sync(FULL, good && fin);
  }
};

}  // namespace als
