// row_update_umma.cuh -- tcgen05/TMEM row-update kernel (placeholder until the
// tensor-core path lands; AUTO falls back to the CUDA-core kernel).
#pragma once
#include "common.cuh"
#include "row_update_simt.cuh"

namespace als {
inline bool umma_supported(int /*ks*/) { return false; }
inline int launch_row_update_umma(int, const RowUpdateParams&, int, cudaStream_t, char* err, size_t n) {
  snprintf(err, n, "tcgen05 kernel not built");
  return ALS_E_UNSUPPORTED;
}
}  // namespace als
