// solve_fp64.cuh -- fp64 LDL^T factorisation + two triangular solves of one k x k SPD
// system held in shared memory, executed by a 128-thread group (a whole 128-thread CTA or
// one warpgroup of a larger CTA, synchronised with a named barrier).
//
// Stands in for MatrixUtils.getSolver(Wu).solveDToF(YTCupu) (AlternatingLeastSquares.java:494
// -> CommonsMathLinearSystemSolver.java:41-45 -> CommonsMathSolver.java:36-44): the
// reference's pivoted Householder QR and this LDL^T produce the same x for an SPD W_u; a
// pivot <= threshold (or non-finite) is reported as ALS_E_SINGULAR like |R_jj| <= 1e-5
// (LinearSystemSolver.java:33-34), a non-finite result as ALS_E_NONFINITE; nothing is
// written for a failed row.
#pragma once
#include "common.cuh"

namespace als {

__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

// W: [KS][KS+1] doubles, lower triangle valid. bvec: rhs in, solution out. invd: scratch [KS].
// tid in [0,128). All 128 threads must call. dst: where the fp32 row is written.
template <int KS>
__device__ __forceinline__ void ldlt_solve_fp64(double* W, double* bvec, double* invd, int k, int tid,
                                                int bar_id, double threshold, DeviceStatus* status,
                                                int which, long long global_row, float* dst) {
  constexpr int LDW = KS + 1;
  const int tx = tid % 16, ty = tid / 16;
  bool failed = false;
  // right-looking LDL^T, one barrier per column. Column j keeps the unscaled values
  // W[i][j] = L[i][j]*d_j; invd[j] = 1/d_j.
  for (int j = 0; j < k; j++) {
    group_barrier(bar_id);
    const double d = W[j * LDW + j];
    if (!(d > threshold) || !isfinite(d)) {  // same value seen by every thread: uniform exit
      if (tid == 0) report_error(status, ALS_E_SINGULAR, which, global_row, (float)d);
      failed = true;
      break;
    }
    const double id = 1.0 / d;
    if (tid == 0) invd[j] = id;
    for (int i = j + 1 + ty; i < k; i += 8) {
      const double lij = W[i * LDW + j] * id;
      for (int c = j + 1 + tx; c <= i; c += 16)
        W[i * LDW + c] = fma(-lij, W[c * LDW + j], W[i * LDW + c]);
    }
  }
  group_barrier(bar_id);
  if (failed || tid >= kWarp) return;
  // forward: z = L^{-1} b (column oriented), y = D^{-1} z, backward: x = L^{-T} y
  for (int j = 0; j < k; j++) {
    const double t = bvec[j] * invd[j];
    for (int i = j + 1 + tid; i < k; i += kWarp) bvec[i] = fma(-W[i * LDW + j], t, bvec[i]);
    __syncwarp();
  }
  for (int i = tid; i < k; i += kWarp) bvec[i] *= invd[i];
  __syncwarp();
  for (int j = k - 1; j >= 0; j--) {
    const double xj = bvec[j];
    for (int i = tid; i < j; i += kWarp) bvec[i] = fma(-W[j * LDW + i] * invd[i], xj, bvec[i]);
    __syncwarp();
  }
  bool bad = false;
  for (int i = tid; i < k; i += kWarp) bad |= !isfinite((float)bvec[i]);
  bad = __any_sync(0xffffffffu, bad);
  if (bad) {
    if (tid == 0) report_error(status, ALS_E_NONFINITE, which, global_row, 0.f);
    return;
  }
  for (int i = tid; i < k; i += kWarp) dst[i] = (float)bvec[i];  // solveDToF's (float) cast
}

}  // namespace als
