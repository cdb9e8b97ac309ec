// resolve_fp64.cuh -- fp64 re-solve of the rows the tensor-core kernel refused in stash mode
// (RowUpdateParams::stash): W_u = G + D_u + lambda alpha n_u I is assembled in fp64 from the fp64
// Gramian and the data term D_u the tensor cores accumulated (handed over as -D_u in the solving
// warps' panel layout, next to the rhs), then factorised and solved by the same fp64 LDL^T as the
// CUDA-core kernel (solve_fp64.cuh), with its error reporting.  The row's entries are not gathered
// again: what made the fp32 solve refuse the row is the conditioning of G, whose fp32 rounding and
// fp32 factorisation are what this path replaces; D_u's accuracy is the fast path's.
//
// Stands in for AlternatingLeastSquares.java:447-494 of such a row (Wu = YTY + ..., solveDToF).
#pragma once
#include "chol_blocked.cuh"
#include "row_update_simt.cuh"
#include "solve_fp64.cuh"

namespace als {

// p: the launch's parameters with row_ptr / n_rows of the rows as stored.  One 128-thread CTA per record.
template <int KS>
__global__ void __launch_bounds__(128) resolve_fp64_kernel(const RowUpdateParams p) {
  using WP = WPanels<KS>;
  constexpr int LDW = KS + 1;
  __shared__ double W[KS * LDW];
  __shared__ double bvec[KS];
  __shared__ double invd[KS];
  const int tid = threadIdx.x;
  const int k = p.k;
  int n_rec = *p.resolve_count;
  if (n_rec > p.resolve_cap) n_rec = p.resolve_cap;
  for (int rec = blockIdx.x; rec < n_rec; rec += gridDim.x) {
    __syncthreads();  // the previous record's solve has finished with W / bvec
    const long long row = p.resolve_rows[rec];
    const float* d = p.resolve_buf + (size_t)rec * (size_t)(WP::kFloats + KS);
    const double lam_n = p.lambda_alpha * (double)(p.row_ptr[row + 1] - p.row_ptr[row]);
    for (int e = tid; e < KS * KS; e += 128) {
      const int r = e / KS, c = e % KS;
      if (c <= r) {
        double v = 0.0;  // padding rows / columns (k < KS) are never factorised
        if (r < k) {
          v = p.G[r * KS + c] - (double)__ldcg(d + WP::at(r, c));
          if (r == c) v += lam_n;
        }
        W[r * LDW + c] = v;
      }
    }
    if (tid < KS) bvec[tid] = (double)__ldcg(d + WP::kFloats + tid);
    // (the solver's first barrier covers the writes above)
    ldlt_solve_fp64<KS>(W, bvec, invd, k, tid, 0, (double)p.threshold, p.status, p.which, p.row_offset + row,
                        p.out + (p.row_offset + row) * KS);
    if (p.n_peers > 0) {
      __syncthreads();  // the row as the solve left it (untouched if it failed)
      if (tid < k) push_to_peers(p, (p.row_offset + row) * KS + tid, p.out[(p.row_offset + row) * KS + tid]);
    }
  }
}

}  // namespace als
