// chol_warp.cuh -- one warp factorises and solves one k x k SPD system entirely in
// registers (fp32), k = KS in {32, 64}.
//
// Row distribution: lane l owns row l (and row l+32 when KS = 64); a row lives in registers
// as float2 pairs along the column index, every index static (the sweep is fully unrolled).
//
// Forward sweep (right-looking LDL^T, one column per step, KS steps):
//   every lane publishes its entries of column j (u[i] = W[i][j]) into a double-buffered
//   KS-float vector in shared memory, the pivot d_j and the rhs entry z_j are shuffled out of
//   the lane that owns row j; one __syncwarp; every lane forms L[i][j] = u[i]/d_j for its own
//   rows (kept in place of W[i][j]), folds the forward substitution b_i -= L[i][j] z_j, and
//   updates its rows right of column j with 16-byte broadcast loads of u and packed fp32x2
//   FMAs:  W[i][c] -= L[i][j] * u[c].  Rows at or above the pivot use a zero multiplier, so
//   entries above the diagonal may hold anything and are never read.
// Backward sweep (x = L^-T D^-1 z), eight columns per block from the last block up:
//   rows below the block contribute  sum_i L[i][jb+t] x_i  (packed FMAs, then a transposing
//   butterfly: 9 shuffles reduce 8 sums over 32 lanes), the 8x8 unit-triangular solve inside
//   the block runs redundantly on every lane with the block's L entries read back from a
//   small shared-memory copy written during the forward sweep (no dependent shuffles).
//
// This is the fast path of MatrixUtils.getSolver(Wu).solveDToF(b) (ALS.java:494): rows whose
// pivots indicate a singular or ill-conditioned W_u (max diag / min pivot > cond_limit, or
// a pivot <= threshold / non-finite) are NOT solved here; the caller re-solves them in fp64
// (row_update_simt.cuh), which also raises ALS_E_SINGULAR exactly as before.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace als {

__device__ __forceinline__ float fast_rcp(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r * fmaf(-d, r, 2.0f);  // one Newton step: ~1 ulp
}

// Compile-time loop: f(integral_constant<int, J>) for J = BEGIN, BEGIN+STEP, ... (END exclusive).
// (#pragma unroll gives up on bodies this large and would demote the register rows to
// local memory.)
template <int J, int END, int STEP = 1>
struct StaticFor {
  template <class F>
  __device__ static __forceinline__ void run(F&& f) {
    if constexpr ((STEP > 0) ? (J < END) : (J > END)) {
      f(std::integral_constant<int, J>{});
      StaticFor<J + STEP, END, STEP>::run(f);
    }
  }
};

template <int KS>
struct CholWarpRows {
  static_assert(KS == 32 || KS == 64, "in-register Cholesky supports k = 32 or 64");
  static constexpr int kRowsPerLane = (KS == 64) ? 2 : 1;
  // matrix row held in slot s of this lane, and whether this lane stores its solution entry
  __device__ static __forceinline__ int row_of(int lane, int s) { return lane + 32 * s; }
  __device__ static __forceinline__ bool writes(int, int) { return true; }
  static constexpr bool kTwoRows = (KS == 64);
  // W planes are "pair-packed lower triangles": for column pair P = (2P, 2P+1) the rows
  // i >= 2P are stored as float2 at float offset offP(P) + 2*(i - 2P).  (The .y slot of row
  // i = 2P is above the diagonal and holds 0.)  Consecutive lanes (rows) touch consecutive
  // 8-byte words: conflict-free for both the drain's stores and the Cholesky warp's loads.
  static constexpr int kPlane = KS * KS / 2 + KS;
  static constexpr int kBlocks = KS / 8;
  static constexpr int kP1 = kTwoRows ? 32 : 1;  // pairs in row lane+32 (dummy for KS=32)
  // per-warp scratch: 2 column buffers of KS entries + (d_j, z_j) + pad, then the 8x8 diagonal
  // blocks of L (row-major, one per block of 8 columns)
  static constexpr int kUBuf = KS + 4;
  static constexpr int kScratch = 2 * kUBuf + kBlocks * 64;  // floats, 16-byte multiple
  __host__ __device__ static constexpr int offP(int P) { return 2 * (KS * P - P * (P - 1)); }

  struct Rows {
    float2 A0[16];   // row `lane`,    columns 0..31
    float2 A1[kP1];  // row `lane+32`, columns 0..63 (KS = 64 only)
    float diag0, diag1;
  };

  // plane (shared memory): pair-packed lower triangle of W_u (G, lambda*alpha*n_u and the unit
  // diagonal of padding rows already folded in by the drain warps). Also returns this lane's
  // diagonal entries (for the conditioning check).
  __device__ static __forceinline__ void load(const float* pl, int lane, Rows& R) {
#pragma unroll
    for (int P = 0; P < 16; P++) {
      float2 v = make_float2(0.f, 0.f);
      if (lane >= 2 * P) v = *reinterpret_cast<const float2*>(pl + offP(P) + 2 * (lane - 2 * P));
      R.A0[P] = v;
    }
    if constexpr (kTwoRows) {
#pragma unroll
      for (int P = 0; P < 32; P++) {
        float2 v = make_float2(0.f, 0.f);
        if (lane + 32 >= 2 * P) v = *reinterpret_cast<const float2*>(pl + offP(P) + 2 * (lane + 32 - 2 * P));
        R.A1[P] = v;
      }
    } else {
      R.A1[0] = make_float2(0.f, 0.f);
    }
    // diagonal entries of my rows: element (r, r) lives in pair r/2, slot r&1
    R.diag0 = pl[offP(lane >> 1) + 2 * (lane - (lane & ~1)) + (lane & 1)];
    R.diag1 = kTwoRows ? pl[offP((lane + 32) >> 1) + 2 * ((lane + 32) - ((lane + 32) & ~1)) + (lane & 1)] : 0.f;
  }

  // element c of a pair-packed register row (c static after unrolling; clamped so dead
  // instantiations never index out of bounds)
  template <int N>
  __device__ static __forceinline__ float get(const float2 (&A)[N], int c) {
    const int p = (c >> 1) < N ? (c >> 1) : 0;
    return (c & 1) ? A[p].y : A[p].x;
  }
  template <int N>
  __device__ static __forceinline__ void set(float2 (&A)[N], int c, float v) {
    const int p = (c >> 1) < N ? (c >> 1) : 0;
    if (c & 1) A[p].y = v; else A[p].x = v;
  }
  // pair p of the row: W[.][2p..2p+1] -= m * (u.x, u.y) for the columns right of column j
  template <int N>
  __device__ static __forceinline__ void update_pair(float2 (&A)[N], int p, int j, float m, float2 nm,
                                                     float2 u) {
    if (p >= N || 2 * p + 1 <= j) return;             // pair finished (or outside this row)
    if (2 * p == j) A[p].y = fmaf(-m, u.y, A[p].y);   // .x is column j itself: now holds L[i][j]
    else A[p] = ffma2(nm, u, A[p]);
  }

  // scratch: kScratch floats of shared memory private to this warp. b0/b1: rhs entries of
  // this lane's rows. k: true feature count; padding rows (j >= k) must carry a unit diagonal
  // and are not judged. Returns (warp-uniform) true if the system was solved; x0/x1 then hold
  // the solution entries of rows lane / lane+32.
  // hook: called once by every lane when the forward sweep reaches column HOOK_STEP (the caller
  // uses it to hand its input slots back to the upstream roles at a chosen point of the sweep)
  template <int HOOK_STEP, class Hook>
  __device__ static __forceinline__ bool factor_solve(Rows& R, float* scratch,
                                                      float (&bx)[kRowsPerLane], float threshold,
                                                      float cond_limit, int lane, int k, Hook&& hook) {
    float x0 = 0.f, x1 = 0.f;
    const bool ok = factor_solve<HOOK_STEP>(R, scratch, bx[0], kTwoRows ? bx[kRowsPerLane - 1] : 0.f,
                                            threshold, cond_limit, lane, k, x0, x1, hook);
    bx[0] = x0;
    if (kTwoRows) bx[kRowsPerLane - 1] = x1;
    return ok;
  }
  template <int HOOK_STEP, class Hook>
  __device__ static __forceinline__ bool factor_solve(Rows& R, float* scratch, float b0, float b1,
                                                      float threshold, float cond_limit, int lane,
                                                      int k, float& x0, float& x1, Hook&& hook) {
    constexpr unsigned FULL = 0xffffffffu;
    float2 (&A0)[16] = R.A0;
    float2 (&A1)[kP1] = R.A1;
    float* Ld = scratch + 2 * kUBuf;  // [kBlocks][8][8] diagonal blocks of L

    // largest diagonal entry (for the conditioning check)
    float dmax;
    {
      float mine = (lane < k) ? R.diag0 : 0.f;
      if (kTwoRows && lane + 32 < k) mine = fmaxf(mine, R.diag1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = fmaxf(mine, __shfl_xor_sync(FULL, mine, o));
      dmax = mine;
    }
    float inv0 = 0.f, inv1 = 0.f;  // 1/d of my rows

    // ---- forward sweep ----------------------------------------------------------------------
    StaticFor<0, KS>::run([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if constexpr (j == HOOK_STEP) hook();
      float* buf = scratch + (j & 1) * kUBuf;
      constexpr bool piv0 = j < 32;  // the pivot row lives in A0 (else in A1)
      constexpr int own = j & 31;    // lane that owns row j
      const float w0 = piv0 ? get(A0, j) : 0.f;
      const float w1 = kTwoRows ? get(A1, j) : 0.f;
      if (piv0) buf[lane] = w0;
      if (kTwoRows) buf[32 + lane] = w1;
      // pivot d_j and rhs entry z_j straight from their owner lane (shorter dependence chain
      // than the round trip through shared memory the column itself takes)
      float2 dz;
      dz.x = __shfl_sync(FULL, piv0 ? w0 : w1, own);
      dz.y = __shfl_sync(FULL, piv0 ? b0 : b1, own);
      __syncwarp();
      const float inv = fast_rcp(dz.x);
      float m0 = 0.f, m1 = 0.f;  // L[i][j] of my rows; 0 for rows at or above the pivot
      if (piv0) {
        m0 = (lane > own) ? w0 * inv : 0.f;
        if (lane == own) inv0 = inv;
        if (kTwoRows) m1 = w1 * inv;
        b0 = fmaf(-m0, dz.y, b0);
        set(A0, j, m0);
      } else {
        m1 = (lane > own) ? w1 * inv : 0.f;
        if (lane == own) inv1 = inv;
      }
      if (kTwoRows) {
        b1 = fmaf(-m1, dz.y, b1);
        set(A1, j, m1);
      }
      const float2 nm0 = make_float2(-m0, -m0), nm1 = make_float2(-m1, -m1);
#pragma unroll
      for (int q = (j + 1) / 4; q < KS / 4; q++) {
        const float4 u = *reinterpret_cast<const float4*>(buf + 4 * q);
        const float2 ua = make_float2(u.x, u.y), ub = make_float2(u.z, u.w);
        if (kTwoRows) {
          update_pair(A1, 2 * q, j, m1, nm1, ua);
          update_pair(A1, 2 * q + 1, j, m1, nm1, ub);
        }
        if (piv0) {
          update_pair(A0, 2 * q, j, m0, nm0, ua);
          update_pair(A0, 2 * q + 1, j, m0, nm0, ub);
        }
      }
      if ((j & 7) == 7) {
        // keep the finished 8x8 diagonal block of L for the backward sweep: the eight lanes
        // owning rows jb..jb+7 write their in-block entries (entries right of the diagonal are
        // never read back)
        const int jb = j - 7;
        if ((lane >> 3) == ((jb & 31) >> 3)) {
          float* dst = Ld + (j >> 3) * 64 + (lane & 7) * 8;
          float4 lo, hi;
          if (piv0) {
            lo = make_float4(get(A0, jb), get(A0, jb + 1), get(A0, jb + 2), get(A0, jb + 3));
            hi = make_float4(get(A0, jb + 4), get(A0, jb + 5), get(A0, jb + 6), get(A0, jb + 7));
          } else {
            lo = make_float4(get(A1, jb), get(A1, jb + 1), get(A1, jb + 2), get(A1, jb + 3));
            hi = make_float4(get(A1, jb + 4), get(A1, jb + 5), get(A1, jb + 6), get(A1, jb + 7));
          }
          *reinterpret_cast<float4*>(dst) = lo;
          *reinterpret_cast<float4*>(dst + 4) = hi;
        }
      }
    });
    // Judge the pivots of the true rows through their reciprocals (each lane kept 1/d of its
    // own rows): d > threshold  <=>  0 < 1/d < 1/threshold, and d * cond_limit >= dmax  <=>
    // dmax / d <= cond_limit.  Zero, negative and NaN pivots fail the first test.  No early
    // exit: the backward sweep runs regardless, so the warp provably stays converged.
    bool good = dmax > threshold && isfinite(dmax);
    {
      const float inv_t = 1.0f / threshold;
      if (lane < k) good = good && inv0 > 0.f && inv0 < inv_t && inv0 * dmax <= cond_limit;
      if (kTwoRows && lane + 32 < k) good = good && inv1 > 0.f && inv1 < inv_t && inv1 * dmax <= cond_limit;
    }
    __syncwarp();  // diagonal blocks visible to every lane

    // ---- y = D^{-1} z, then x = L^{-T} y, block-wise from the last block ---------------------
    x0 = b0 * inv0;
    x1 = kTwoRows ? b1 * inv1 : 0.f;
    const bool k4 = (lane & 4) != 0, k2 = (lane & 2) != 0, k1 = (lane & 1) != 0;
    StaticFor<kBlocks - 1, -1, -1>::run([&](auto bc) {
      constexpr int b = decltype(bc)::value;
      constexpr int jb = 8 * b;
      constexpr bool in0 = jb < 32;               // the block's rows live in A0/x0 (else A1/x1)
      constexpr int grp = (jb & 31) >> 3;         // lanes 8*grp .. 8*grp+7 own rows jb..jb+7
      const bool mine = (lane >> 3) == grp;
      const float y = in0 ? x0 : x1;          // y_{jb + (lane & 7)} on the owner lanes
      float v;                                // y_j - sum_{i >= jb+8} L[i][j] x_i for j = jb + (lane & 7)
      if (jb + 8 < KS) {
        float2 acc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) acc[q] = make_float2(0.f, 0.f);
        if (kTwoRows) {
          float xm = x1;                      // rows lane+32 >= jb+8 only
          if (jb + 8 > 32) xm = (lane + 32 >= jb + 8) ? x1 : 0.f;
          const float2 xx = make_float2(xm, xm);
#pragma unroll
          for (int q = 0; q < 4; q++) acc[q] = ffma2(A1[(jb / 2 + q) < kP1 ? (jb / 2 + q) : 0], xx, acc[q]);
        }
        if (jb + 8 < 32) {
          const float xm = (lane >= jb + 8) ? x0 : 0.f;
          const float2 xx = make_float2(xm, xm);
#pragma unroll
          for (int q = 0; q < 4; q++) acc[q] = ffma2(A0[(jb / 2 + q) & 15], xx, acc[q]);
        }
        // transposing butterfly: after the xor-4/2/1 rounds a lane holds the sum for
        // t = lane & 7 over its group of 8 lanes; xor-8/16 finish it over the warp
        const float s0 = acc[0].x, s1 = acc[0].y, s2 = acc[1].x, s3 = acc[1].y;
        const float s4 = acc[2].x, s5 = acc[2].y, s6 = acc[3].x, s7 = acc[3].y;
        const float r0 = (k4 ? s4 : s0) + __shfl_xor_sync(FULL, k4 ? s0 : s4, 4);
        const float r1 = (k4 ? s5 : s1) + __shfl_xor_sync(FULL, k4 ? s1 : s5, 4);
        const float r2 = (k4 ? s6 : s2) + __shfl_xor_sync(FULL, k4 ? s2 : s6, 4);
        const float r3 = (k4 ? s7 : s3) + __shfl_xor_sync(FULL, k4 ? s3 : s7, 4);
        const float q0 = (k2 ? r2 : r0) + __shfl_xor_sync(FULL, k2 ? r0 : r2, 2);
        const float q1 = (k2 ? r3 : r1) + __shfl_xor_sync(FULL, k2 ? r1 : r3, 2);
        float p = (k1 ? q1 : q0) + __shfl_xor_sync(FULL, k1 ? q0 : q1, 1);
        p = mine ? y - p : -p;
        p += __shfl_xor_sync(FULL, p, 8);
        p += __shfl_xor_sync(FULL, p, 16);
        v = p;
      } else {
        v = y;  // last block: nothing below it
      }
      float s[8];
#pragma unroll
      for (int t = 0; t < 8; t++) s[t] = __shfl_sync(FULL, v, (jb + 8 < KS) ? t : 8 * grp + t);
      // in-block unit-upper-triangular solve, redundantly on every lane:
      // x_{jb+t} = s_t - sum_{t' > t} L[jb+t'][jb+t] x_{jb+t'}
      const float* Lb = Ld + b * 64;
#pragma unroll
      for (int tp = 7; tp >= 1; tp--) {
        const float4 la = *reinterpret_cast<const float4*>(Lb + tp * 8);
        float l[8] = {la.x, la.y, la.z, la.w, 0.f, 0.f, 0.f, 0.f};
        if (tp > 4) {
          const float4 lb = *reinterpret_cast<const float4*>(Lb + tp * 8 + 4);
          l[4] = lb.x; l[5] = lb.y; l[6] = lb.z; l[7] = lb.w;
        }
#pragma unroll
        for (int t = 0; t < tp; t++) s[t] = fmaf(-l[t], s[tp], s[t]);
      }
      const float sa = k1 ? s[1] : s[0], sb = k1 ? s[3] : s[2], sc = k1 ? s[5] : s[4], sd = k1 ? s[7] : s[6];
      const float se = k2 ? sb : sa, sf = k2 ? sd : sc;
      const float xs = k4 ? sf : se;  // s[lane & 7]
      if (mine) { if (in0) x0 = xs; else x1 = xs; }
    });
    return __all_sync(FULL, good && isfinite(x0) && isfinite(x1));
  }
};


template <int KS>
using CholWarp = CholWarpRows<KS>;

}  // namespace als
