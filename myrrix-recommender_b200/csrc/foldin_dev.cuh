// foldin_dev.cuh -- online writes folded into the RESIDENT factor rows: the live
// model stays in HBM between builds, top-N queries (topn.cuh) see every write at once.
//
// ServerRecommender.updateFeatures (online/src/net/myrrix/online/ServerRecommender.java:865-907):
//   estimate = dot(x_u, y_i); w = foldInWeight(estimate, value) (:981-994); if w != 0:
//   itemFold = XTXsolver.solveFToD(x_u), userFold = YTYsolver.solveFToD(y_i) (both from the rows as
//   they were on entry), y_i += (float) (w * itemFold), x_u += (float) (w * userFold).
// The solvers are the generation's (Generation.recomputeState, online/.../generation/Generation.java
// :132-158): commons-math3's pivoted QR of X'X / Y'Y, computed ONCE per generation by
// libmyrrix_foldin.so from als_gramian's output and copied here (als_set_fold_in_state): qrt, rdiag,
// perm.  Solving applies exactly that factorisation's steps (Q' b by Householder vectors, back
// substitution against R, un-permute; csrc_host/foldin.cpp Rrqr::solve), one warp per system with
// the dot products summed across lanes -- so the fp64 intermediates can differ from the host
// library's in the last bit, the fp32 rows only where a delta sits on a rounding boundary.
// Writes are applied IN ORDER by one CTA (each sees the rows its predecessors left, like the
// reference's one-setPreference-at-a-time stream): warp 0 solves against X'X, warp 1 against Y'Y.
#pragma once
#include "common.cuh"

namespace als {

struct FoldInSolver {
  const double* qrt;    // [k][k]
  const double* rdiag;  // [k]
  const int* perm;      // [k]
};

__device__ __forceinline__ double fold_in_weight_dev(double learn_rate, double estimate, float value) {
  double w;
  if (value > 0.0f && estimate < 1.0) {
    const double multiplier = 1.0 - fmax(0.0, estimate);
    w = (1.0 - 1.0 / (1.0 + (double)value)) * multiplier;
  } else if (value < 0.0f && estimate > 0.0) {
    const double multiplier = -fmin(1.0, estimate);
    w = (1.0 - 1.0 / (1.0 - (double)value)) * multiplier;
  } else {
    w = 0.0;
  }
  return learn_rate * w;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x = (M'M)^-1 b by the stored factorisation; b: k floats in shared memory; out: k doubles in
// shared memory.  One warp; lane l holds entries l, l + 32, ... of the working vector.
template <int NS>
__device__ __forceinline__ void qr_solve_warp(const FoldInSolver& S, int k, const float* b, double* out, int lane) {
  double y[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int r = lane + 32 * s;
    y[s] = r < k ? (double)b[r] : 0.0;
  }
  for (int minor = 0; minor < k; minor++) {  // y = Q' b
    const double* col = S.qrt + (size_t)minor * k;
    double c[NS], part = 0.0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int r = lane + 32 * s;
      c[s] = (r >= minor && r < k) ? __ldg(col + r) : 0.0;
      part += y[s] * c[s];
    }
    double dot = warp_sum_f64(part);
    dot /= __ldg(S.rdiag + minor) * __ldg(col + minor);
#pragma unroll
    for (int s = 0; s < NS; s++) y[s] += dot * c[s];
  }
  for (int row = k - 1; row >= 0; row--) {  // R z = y
    double yr = 0.0;
#pragma unroll
    for (int s = 0; s < NS; s++)
      if (s == row / 32) yr = __shfl_sync(0xffffffffu, y[s], row & 31);
    yr /= __ldg(S.rdiag + row);
    if (lane == 0) out[__ldg(S.perm + row)] = yr;  // x[perm[j]] = z[j]
    const double* col = S.qrt + (size_t)row * k;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      const int r = lane + 32 * s;
      if (r < row) y[s] -= yr * __ldg(col + r);
    }
  }
}

// status: 0 ok, 2 (ALS_E_NONFINITE) a non-finite estimate or delta (Preconditions.checkState, :872, :889)
template <int NS>
__global__ void __launch_bounds__(64) fold_in_kernel(float* X, float* Y, int ks, int k, const int* users,
                                                      const int* items, const float* values, long long n,
                                                      FoldInSolver sx, FoldInSolver sy, double learn_rate,
                                                      int* status) {
  __shared__ float xu[kMaxFeatures], yi[kMaxFeatures];
  __shared__ double fold[2][kMaxFeatures];
  __shared__ double w_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long e = 0; e < n; e++) {
    float* xrow = X + (long long)users[e] * ks;
    float* yrow = Y + (long long)items[e] * ks;
    for (int f = threadIdx.x; f < k; f += 64) {
      xu[f] = xrow[f];
      yi[f] = yrow[f];
    }
    __syncthreads();
    if (warp == 0) {
      double part = 0.0;  // SimpleVectorMath.dot: fp32 products, fp64 sum
      for (int f = lane; f < k; f += 32) part += (double)__fmul_rn(xu[f], yi[f]);
      const double est = warp_sum_f64(part);
      if (lane == 0) {
        if (!isfinite(est)) {
          *status = ALS_E_NONFINITE;
          w_s = 0.0;
        } else {
          w_s = fold_in_weight_dev(learn_rate, est, values ? values[e] : 1.0f);
        }
      }
    }
    __syncthreads();
    const double w = w_s;
    if (w != 0.0) {
      if (warp == 0 && sx.qrt) qr_solve_warp<NS>(sx, k, xu, fold[0], lane);  // itemFold (:876-879)
      if (warp == 1 && sy.qrt) qr_solve_warp<NS>(sy, k, yi, fold[1], lane);  // userFold (:880-884)
      __syncthreads();
      bool bad = false;
      for (int f = threadIdx.x; f < k; f += 64) {
        if (sx.qrt) {
          const double d = w * fold[0][f];
          if (!isfinite(d)) bad = true; else yrow[f] = yi[f] + (float)d;
        }
        if (sy.qrt) {
          const double d = w * fold[1][f];
          if (!isfinite(d)) bad = true; else xrow[f] = xu[f] + (float)d;
        }
      }
      if (bad) *status = ALS_E_NONFINITE;
    }
    __syncthreads();  // (also orders this write's global stores before the next write's loads in this CTA)
  }
}

}  // namespace als
