// chol_warp.cuh -- one warp factorises and solves one k x k SPD system entirely in
// registers (fp32), k = KS in {32, 64}.
//
// Row distribution: lane l owns row l (and row l+32 when KS = 64); a row lives in
// registers as float2 pairs along the column index.  Right-looking LDL^T in blocks of 8
// columns: the block loop is a REAL loop (compact code -- the instruction footprint of a
// fully unrolled 64-step sweep thrashes the instruction cache), made possible by rotating
// each row's register array by 4 pairs after every block so the active block always sits
// at register positions 0..3.  Per step j the (unscaled) column j, with zeros for rows <= j,
// is published to a 2-deep shared-memory buffer; every lane reads it back as broadcast
// 16-byte loads and applies   W[i][c] -= (w_i / d_j) * w_c   with packed fp32x2 FMAs, whole
// 8-column groups at a time (finished groups are skipped by a warp-uniform test).
// The forward substitution z = L^{-1} b is fused into the same sweep; D^{-1} is lane-local;
// the backward substitution x = L^{-T} y runs block-wise with warp all-reduces.
//
// This is the fast path of MatrixUtils.getSolver(Wu).solveDToF(b) (ALS.java:494): rows whose
// pivots indicate a singular or ill-conditioned W_u (max diag / min pivot > cond_limit, or
// a pivot <= threshold / non-finite) are NOT solved here; the caller re-solves them in fp64
// (row_update_simt.cuh), which also raises ALS_E_SINGULAR exactly as before.
#pragma once
#include "common.cuh"

namespace als {

// shared-space accessors (32-bit shared addresses: guarantees LDS/STS, no generic loads)
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32x2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
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
  static constexpr int kColBuf = KS + 8;            // column + z_j + d_j, 16-byte multiple
  static constexpr int kBlocks = KS / 8;
  static constexpr int kP1 = kTwoRows ? 32 : 4;     // pairs in row lane+32 (dummy 4 for KS=32)
  __host__ __device__ static constexpr int offP(int P) { return 2 * (KS * P - P * (P - 1)); }

  struct Rows {
    float2 A0[16];   // row `lane`,    columns 0..31
    float2 A1[kP1];  // row `lane+32`, columns 0..63 (KS = 64 only)
  };

  // planes p0 + p1 (shared addresses): pair-packed lower triangles whose sum is W_u (G and
  // lambda*alpha*n_u already folded in by the drain warps).
  __device__ static __forceinline__ void load(uint32_t p0, uint32_t p1, int lane, Rows& R) {
#pragma unroll
    for (int P = 0; P < 16; P++) {
      float2 v = make_float2(0.f, 0.f);
      if (lane >= 2 * P) {
        const uint32_t o = (uint32_t)(offP(P) + 2 * (lane - 2 * P)) * 4u;
        const float2 a = lds_f32x2(p0 + o), c = lds_f32x2(p1 + o);
        v = make_float2(a.x + c.x, a.y + c.y);
      }
      R.A0[P] = v;
    }
    if (kTwoRows) {
#pragma unroll
      for (int P = 0; P < 32; P++) {
        float2 v = make_float2(0.f, 0.f);
        if (lane + 32 >= 2 * P) {
          const uint32_t o = (uint32_t)(offP(P) + 2 * (lane + 32 - 2 * P)) * 4u;
          const float2 a = lds_f32x2(p0 + o), c = lds_f32x2(p1 + o);
          v = make_float2(a.x + c.x, a.y + c.y);
        }
        R.A1[P] = v;
      }
    }
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

  // colbuf: shared address of [2][kColBuf] floats private to this warp. b0/b1: rhs entries of
  // this lane's rows. k: true feature count; padding rows (j >= k) must carry a unit diagonal
  // and are not judged. Returns (warp-uniform) true if the system was solved; x0/x1 then hold
  // the solution entries of rows lane / lane+32.
  __device__ static __forceinline__ bool factor_solve(Rows& R, uint32_t colbuf, float b0, float b1,
                                                      float threshold, float cond_limit, int lane,
                                                      int k, float& x0, float& x1) {
    constexpr unsigned FULL = 0xffffffffu;
    float2 (&A0)[16] = R.A0;
    float2 (&A1)[kP1] = R.A1;

    // largest diagonal entry (for the conditioning check)
    float dmax;
    {
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j++)
        if (j == lane && j < k) mine = (j & 1) ? A0[j >> 1].y : A0[j >> 1].x;
      if (kTwoRows) {
#pragma unroll
        for (int j = 32; j < 64; j++)
          if (j == lane + 32 && j < k) mine = fmaxf(mine, (j & 1) ? A1[j >> 1].y : A1[j >> 1].x);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = fmaxf(mine, __shfl_xor_sync(FULL, mine, o));
      dmax = mine;
    }
    float dmin = dmax;
    bool bad = !(dmax > threshold) || !isfinite(dmax);
    float inv0 = 0.f, inv1 = 0.f;  // 1/d of my rows

    // ---- LDL^T + fused forward substitution, 8 columns per trip ---------------------------
#pragma unroll 1
    for (int b = 0; b < kBlocks; b++) {
      const int jb = 8 * b;
      const bool a0_live = jb < 32;     // row `lane` still has unfinished columns
      const int groups = kBlocks - b;   // live 8-column groups of the longest row
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int j = jb + t;
        const uint32_t cb = colbuf + (uint32_t)((t & 1) * kColBuf) * 4u;
        const float w0 = (t & 1) ? A0[t >> 1].y : A0[t >> 1].x;
        const float w1 = kTwoRows ? ((t & 1) ? A1[t >> 1].y : A1[t >> 1].x) : 0.f;
        // publish column j: rows <= j contribute zeros so finished columns are left alone
        if (a0_live) sts_f32(cb + (uint32_t)lane * 4u, (lane > j) ? w0 : 0.f);
        if (kTwoRows) sts_f32(cb + (uint32_t)(lane + 32) * 4u, (lane + 32 > j) ? w1 : 0.f);
        if (lane == (j & 31)) {
          const bool in0 = j < 32;
          sts_f32(cb + (uint32_t)KS * 4u, in0 ? b0 : b1);        // z_j (unit-lower L)
          sts_f32(cb + (uint32_t)(KS + 1) * 4u, in0 ? w0 : w1);  // pivot d_j
        }
        __syncwarp();
        const float zj = lds_f32(cb + (uint32_t)KS * 4u);
        const float d = lds_f32(cb + (uint32_t)(KS + 1) * 4u);
        if (j < k) {
          bad = bad || !(d > threshold) || !isfinite(d);
          dmin = fminf(dmin, d);
        }
        const float inv = fast_rcp(d);
        const float m0 = a0_live ? w0 * inv : 0.f;
        const float m1 = w1 * inv;  // L[i][j]
        if (lane == (j & 31)) { if (j < 32) inv0 = inv; else inv1 = inv; }
        // keep L[i][j] in place of the unscaled entry
        if (a0_live) { if (t & 1) A0[t >> 1].y = m0; else A0[t >> 1].x = m0; }
        if (kTwoRows) { if (t & 1) A1[t >> 1].y = m1; else A1[t >> 1].x = m1; }
        // forward substitution: b_i -= L[i][j] * z_j for rows below j
        if (a0_live && lane > j) b0 = fmaf(-m0, zj, b0);
        if (kTwoRows && lane + 32 > j) b1 = fmaf(-m1, zj, b1);
        // trailing update, whole 8-column groups; group g covers columns jb+8g .. jb+8g+7
        const float2 nm0 = make_float2(-m0, -m0), nm1 = make_float2(-m1, -m1);
#pragma unroll
        for (int g = 0; g < kBlocks; g++) {
          if (g < groups) {  // warp-uniform
            const float4 qa = lds_f32x4(cb + (uint32_t)(jb + 8 * g) * 4u);
            const float4 qb = lds_f32x4(cb + (uint32_t)(jb + 8 * g + 4) * 4u);
            const float2 c0 = make_float2(qa.x, qa.y), c1 = make_float2(qa.z, qa.w);
            const float2 c2 = make_float2(qb.x, qb.y), c3 = make_float2(qb.z, qb.w);
            if (kTwoRows) {
              A1[4 * g + 0] = ffma2(nm1, c0, A1[4 * g + 0]);
              A1[4 * g + 1] = ffma2(nm1, c1, A1[4 * g + 1]);
              A1[4 * g + 2] = ffma2(nm1, c2, A1[4 * g + 2]);
              A1[4 * g + 3] = ffma2(nm1, c3, A1[4 * g + 3]);
            }
            if (g < 4 && jb + 8 * g < 32) {  // row `lane` has only columns 0..31 (warp-uniform)
              A0[4 * g + 0] = ffma2(nm0, c0, A0[4 * g + 0]);
              A0[4 * g + 1] = ffma2(nm0, c1, A0[4 * g + 1]);
              A0[4 * g + 2] = ffma2(nm0, c2, A0[4 * g + 2]);
              A0[4 * g + 3] = ffma2(nm0, c3, A0[4 * g + 3]);
            }
          }
        }
      }
      // rotate so the next block's columns sit at positions 0..3 (finished L goes to the end)
      if (kTwoRows) rotate_left4(A1);
      if (a0_live) rotate_left4(A0);
    }
    // after KS/8 rotations of A1 (32 pairs, 8 blocks) and 4 of A0 (16 pairs) both arrays are
    // back in natural column order and hold L (unit lower, strictly below the diagonal).
    if (bad || !(dmin * cond_limit >= dmax)) return false;

    // ---- y = D^{-1} z, then x = L^{-T} y, block-wise from the last block ---------------------
    x0 = b0 * inv0;
    x1 = kTwoRows ? b1 * inv1 : 0.f;
#pragma unroll 1
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
