// chol_warp.cuh -- one warp factorises and solves one k x k SPD system entirely in
// registers (fp32), k = KS in {32, 64}.
//
// Row distribution: lane l owns row l (and row l+32 when KS = 64); a row lives in
// registers as float2 pairs along the column index.  Panel-blocked right-looking LDL^T,
// 8 columns per block, block loop is a REAL loop (compact code: the four roles of the
// row-update kernel must share the instruction cache) made possible by rotating each row's
// register array by 4 pairs per block so the active panel always sits at positions 0..3:
//   1. eight cheap in-panel steps: the 8 owner lanes of rows jb..jb+7 publish their entry of
//      column j (10 floats incl. pivot and z_j), every lane scales its own entry and updates
//      only the panel columns of its rows; the forward substitution z = L^{-1} b rides along;
//   2. every lane publishes the unscaled panel entries u_c[t] of its rows (transposed:
//      Ut[t][c]), then one rank-8 update   W[i][c] -= sum_t L[i][jb+t] * u_c[t]   of all
//      columns right of the panel: 16-byte broadcast loads + packed fp32x2 FMAs, whole
//      8-column groups at a time (finished groups skipped by a warp-uniform test).
// D^{-1} is lane-local; the backward substitution x = L^{-T} y runs block-wise with warp
// all-reduces for the part below the block and an 8x8 unit-triangular solve inside it.
//
// This is the fast path of MatrixUtils.getSolver(Wu).solveDToF(b) (ALS.java:494): rows whose
// pivots indicate a singular or ill-conditioned W_u (max diag / min pivot > cond_limit, or
// a pivot <= threshold / non-finite) are NOT solved here; the caller re-solves them in fp64
// (row_update_simt.cuh), which also raises ALS_E_SINGULAR exactly as before.
#pragma once
#include "common.cuh"

namespace als {

__device__ __forceinline__ float fast_rcp(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r * fmaf(-d, r, 2.0f);  // one Newton step: ~1 ulp
}

template <int KS>
struct CholWarp {
  static_assert(KS == 32 || KS == 64, "in-register Cholesky supports k = 32 or 64");
  static constexpr bool kTwoRows = (KS == 64);
  // W planes are "pair-packed lower triangles": for column pair P = (2P, 2P+1) the rows
  // i >= 2P are stored as float2 at float offset offP(P) + 2*(i - 2P).  (The .y slot of row
  // i = 2P is above the diagonal and holds 0.)  Consecutive lanes (rows) touch consecutive
  // 8-byte words: conflict-free for both the drain's stores and the Cholesky warp's loads.
  static constexpr int kPlane = KS * KS / 2 + KS;
  static constexpr int kBlocks = KS / 8;
  static constexpr int kP1 = kTwoRows ? 32 : 4;  // pairs in row lane+32 (dummy 4 for KS=32)
  // per-warp scratch: 2 step buffers of 12 floats (8 column entries, z_j, d_j) + Ut[8][KS] + Lt[8][KS]
  static constexpr int kStepBuf = 12;
  static constexpr int kScratch = 2 * kStepBuf + 16 * KS;  // floats, 16-byte multiple
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
    if (kTwoRows) {
#pragma unroll
      for (int P = 0; P < 32; P++) {
        float2 v = make_float2(0.f, 0.f);
        if (lane + 32 >= 2 * P) v = *reinterpret_cast<const float2*>(pl + offP(P) + 2 * (lane + 32 - 2 * P));
        R.A1[P] = v;
      }
    }
    // diagonal entries of my rows: element (r, r) lives in pair r/2, slot r&1
    R.diag0 = pl[offP(lane >> 1) + 2 * (lane - (lane & ~1)) + (lane & 1)];
    R.diag1 = kTwoRows ? pl[offP((lane + 32) >> 1) + 2 * ((lane + 32) - ((lane + 32) & ~1)) + (lane & 1)] : 0.f;
  }

  template <int N>
  __device__ static __forceinline__ void rotate_left4(float2 (&A)[N]) {
    const float2 t0 = A[0], t1 = A[1], t2 = A[2], t3 = A[3];
#pragma unroll
    for (int P = 0; P + 4 < N; P++) A[P] = A[P + 4];
    A[N - 4] = t0; A[N - 3] = t1; A[N - 2] = t2; A[N - 1] = t3;
  }
  template <int N>
  __device__ static __forceinline__ void rotate_right4(float2 (&A)[N]) {
    const float2 t0 = A[N - 4], t1 = A[N - 3], t2 = A[N - 2], t3 = A[N - 1];
#pragma unroll
    for (int P = N - 1; P >= 4; P--) A[P] = A[P - 4];
    A[0] = t0; A[1] = t1; A[2] = t2; A[3] = t3;
  }

  // scratch: kScratch floats of shared memory private to this warp. b0/b1: rhs entries of
  // this lane's rows. k: true feature count; padding rows (j >= k) must carry a unit diagonal
  // and are not judged. Returns (warp-uniform) true if the system was solved; x0/x1 then hold
  // the solution entries of rows lane / lane+32.
  __device__ static __forceinline__ bool factor_solve(Rows& R, float* scratch, float b0, float b1,
                                                      float threshold, float cond_limit, int lane,
                                                      int k, float& x0, float& x1) {
    constexpr unsigned FULL = 0xffffffffu;
    float2 (&A0)[16] = R.A0;
    float2 (&A1)[kP1] = R.A1;
    float* Ut = scratch + 2 * kStepBuf;  // [8][KS] unscaled panel entries, transposed
    float* Lt = Ut + 8 * KS;             // [8][KS] L panel entries, transposed

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
    float d0 = dmax, d1 = dmax;    // pivots of my rows (judged after the sweep)

#pragma unroll
    for (int b = 0; b < kBlocks; b++) {
      const int jb = 8 * b;
      const bool a0_live = jb < 32;     // row `lane` still has unfinished columns
      const bool own_in0 = jb < 32;     // the panel's diagonal rows jb..jb+7 live in A0 (else A1)
      const int tl = lane - (jb & 31);  // 0..7 on the lanes owning rows jb..jb+7
      const bool owner = tl >= 0 && tl < 8;
      // ---- 1. eight in-panel steps, two per trip; the panel's 4 pairs rotate by one pair per
      //         trip so the active pair is always position 0 (compact code) ------------------
#pragma unroll 1
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int t = 2 * r + h;
          const int j = jb + t;
          float* sb = scratch + h * kStepBuf;
          const float w0 = h ? A0[0].y : A0[0].x;
          const float w1 = kTwoRows ? (h ? A1[0].y : A1[0].x) : 0.f;
          if (owner) {
            const float mine = own_in0 ? w0 : w1;
            // slot of panel row tl in the rotated frame; rows <= j contribute zeros
            sb[(tl - 2 * r) & 7] = (tl > t) ? mine : 0.f;
            if (tl == t) {
              sb[8] = own_in0 ? b0 : b1;  // z_j (unit-lower L)
              sb[9] = mine;               // pivot d_j
            }
          }
          __syncwarp();
          const float4 qa = *reinterpret_cast<const float4*>(sb);
          const float4 qb = *reinterpret_cast<const float4*>(sb + 4);
          const float2 zd = *reinterpret_cast<const float2*>(sb + 8);
          const float zj = zd.x, d = zd.y;
          const float inv = fast_rcp(d);
          const float m0 = a0_live ? w0 * inv : 0.f;
          const float m1 = w1 * inv;  // L[i][j]
          if (tl == t) { if (own_in0) { inv0 = inv; d0 = d; } else { inv1 = inv; d1 = d; } }
          // forward substitution: b_i -= L[i][j] * z_j for rows below j
          if (a0_live && lane > j) b0 = fmaf(-m0, zj, b0);
          if (kTwoRows && lane + 32 > j) b1 = fmaf(-m1, zj, b1);
          // update the panel columns of my rows (zeros in sb leave finished columns alone)
          const float2 c0 = make_float2(qa.x, qa.y), c1 = make_float2(qa.z, qa.w);
          const float2 c2 = make_float2(qb.x, qb.y), c3 = make_float2(qb.z, qb.w);
          if (a0_live) {
            const float2 nm0 = make_float2(-m0, -m0);
            A0[0] = ffma2(nm0, c0, A0[0]); A0[1] = ffma2(nm0, c1, A0[1]);
            A0[2] = ffma2(nm0, c2, A0[2]); A0[3] = ffma2(nm0, c3, A0[3]);
          }
          if (kTwoRows) {
            const float2 nm1 = make_float2(-m1, -m1);
            A1[0] = ffma2(nm1, c0, A1[0]); A1[1] = ffma2(nm1, c1, A1[1]);
            A1[2] = ffma2(nm1, c2, A1[2]); A1[3] = ffma2(nm1, c3, A1[3]);
          }
          // publish for the rank-8 update: the unscaled entry u_c[t] of my rows (rows inside or
          // above the panel publish 0) and L[i][j]; keep L[i][j] in place of the entry
          if (a0_live) {
            Ut[t * KS + lane] = (lane >= jb + 8) ? w0 : 0.f;
            Lt[t * KS + lane] = m0;
            if (h) A0[0].y = m0; else A0[0].x = m0;
          }
          if (kTwoRows) {
            Ut[t * KS + lane + 32] = (lane + 32 >= jb + 8) ? w1 : 0.f;
            Lt[t * KS + lane + 32] = m1;
            if (h) A1[0].y = m1; else A1[0].x = m1;
          }
        }
        // rotate the panel pairs by one
        if (a0_live) { const float2 t0 = A0[0]; A0[0] = A0[1]; A0[1] = A0[2]; A0[2] = A0[3]; A0[3] = t0; }
        if (kTwoRows) { const float2 t1 = A1[0]; A1[0] = A1[1]; A1[1] = A1[2]; A1[2] = A1[3]; A1[3] = t1; }
      }
      __syncwarp();
      // ---- 2. rank-8 update of the columns right of the panel ------------------------------
      const int groups = kBlocks - b;  // live 8-column groups incl. the panel itself (g = 0)
      if (groups > 1) {
#pragma unroll 1
        for (int t = 0; t < 8; t++) {
          const float l0 = a0_live ? Lt[t * KS + lane] : 0.f;
          const float l1 = kTwoRows ? Lt[t * KS + lane + 32] : 0.f;
          const float2 nl0 = make_float2(-l0, -l0), nl1 = make_float2(-l1, -l1);
          const float* ut = Ut + t * KS + jb;
#pragma unroll
          for (int g = 1; g < kBlocks; g++) {
            if (g < groups) {  // warp-uniform
              const float4 qa = *reinterpret_cast<const float4*>(ut + 8 * g);
              const float4 qb = *reinterpret_cast<const float4*>(ut + 8 * g + 4);
              const float2 c0 = make_float2(qa.x, qa.y), c1 = make_float2(qa.z, qa.w);
              const float2 c2 = make_float2(qb.x, qb.y), c3 = make_float2(qb.z, qb.w);
              if (kTwoRows) {
                A1[4 * g + 0] = ffma2(nl1, c0, A1[4 * g + 0]);
                A1[4 * g + 1] = ffma2(nl1, c1, A1[4 * g + 1]);
                A1[4 * g + 2] = ffma2(nl1, c2, A1[4 * g + 2]);
                A1[4 * g + 3] = ffma2(nl1, c3, A1[4 * g + 3]);
              }
              if (g < 4 && jb + 8 * g < 32) {  // row `lane` has only columns 0..31 (warp-uniform)
                A0[4 * g + 0] = ffma2(nl0, c0, A0[4 * g + 0]);
                A0[4 * g + 1] = ffma2(nl0, c1, A0[4 * g + 1]);
                A0[4 * g + 2] = ffma2(nl0, c2, A0[4 * g + 2]);
                A0[4 * g + 3] = ffma2(nl0, c3, A0[4 * g + 3]);
              }
            }
          }
        }
      }
      __syncwarp();  // Ut / Lt are rewritten by the next block
      // rotate so the next panel sits at positions 0..3 (finished L goes to the end)
      if (kTwoRows) rotate_left4(A1);
      if (a0_live) rotate_left4(A0);
    }
    // judge the pivots: smallest pivot vs threshold / largest diagonal entry
    float dmin;
    {
      float mine = dmax;
      if (lane < k) mine = fminf(mine, d0);
      if (kTwoRows && lane + 32 < k) mine = fminf(mine, d1);
      bool fin = isfinite(d0) && isfinite(d1) && isfinite(inv0) && isfinite(inv1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = fminf(mine, __shfl_xor_sync(FULL, mine, o));
      dmin = __all_sync(FULL, fin) ? mine : -1.f;
    }
    const bool bad = !(dmax > threshold) || !isfinite(dmax) || !(dmin > threshold);
    // after KS/8 rotations of A1 (32 pairs, 8 blocks) and 4 of A0 (16 pairs) both arrays are
    // back in natural column order and hold L (unit lower, strictly below the diagonal).
    if (bad || !(dmin * cond_limit >= dmax)) return false;

    // ---- y = D^{-1} z, then x = L^{-T} y, block-wise from the last block ---------------------
    x0 = b0 * inv0;
    x1 = kTwoRows ? b1 * inv1 : 0.f;
#pragma unroll
    for (int b = kBlocks - 1; b >= 0; b--) {
      const int jb = 8 * b;
      const bool in0 = jb < 32;  // the block's rows live in A0/x0 (else A1/x1)
      if (kTwoRows) rotate_right4(A1);
      if (in0) rotate_right4(A0);
      // s[t] = y_j - sum over rows i >= jb+8 of L[i][j] x_i   (j = jb+t), all-reduced
      float s[8];
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int j = jb + t;
        float part = 0.f;
        if (in0 && lane >= jb + 8) part = ((t & 1) ? A0[t >> 1].y : A0[t >> 1].x) * x0;
        if (kTwoRows && lane + 32 >= jb + 8)
          part = fmaf((t & 1) ? A1[t >> 1].y : A1[t >> 1].x, x1, part);
        part = -part;
        if (lane == (j & 31)) part += in0 ? x0 : x1;  // + y_j from its owner
        s[t] = part;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int t = 0; t < 8; t++) s[t] += __shfl_xor_sync(FULL, s[t], o);
      }
      // in-block unit-upper-triangular solve, redundantly on every lane:
      // x_j = s_j - sum_{t' > t} L[jb+t'][jb+t] x_{jb+t'}; L[jb+t'][.] sits in lane (jb+t')&31
#pragma unroll
      for (int t = 6; t >= 0; t--) {
        const float mine = in0 ? ((t & 1) ? A0[t >> 1].y : A0[t >> 1].x)
                               : (kTwoRows ? ((t & 1) ? A1[t >> 1].y : A1[t >> 1].x) : 0.f);
#pragma unroll
        for (int tp = 7; tp > t; tp--) {
          const float l = __shfl_sync(FULL, mine, (jb + tp) & 31);
          s[t] = fmaf(-l, s[tp], s[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 8; t++)
        if (lane == ((jb + t) & 31)) { if (in0) x0 = s[t]; else x1 = s[t]; }
    }
    return __all_sync(FULL, isfinite(x0) && isfinite(x1));
  }
};

}  // namespace als
