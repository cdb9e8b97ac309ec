// gather4_probe.cu -- what does cp.async.bulk.tensor.2d ... tile::gather4 do on this part?
// Builds a [rows x 64] fp32 matrix with M[r][c] = r * 1000 + c, encodes tiled tensor maps with box
// {64, 1} and {64, 4} (whichever the driver accepts), and lets ONE thread gather four rows (one of
// them out of bounds) into shared memory; prints what landed and how many bytes the mbarrier saw.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/gather4_probe scripts/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int col0,
                      uint32_t expect_bytes, float* out, int* flags) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  float* tile = reinterpret_cast<float*>(smem);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = -1.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(expect_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
        : "memory");
  }
  // wait with a timeout: did the expected byte count arrive?
  int ok = 0;
  for (int spin = 0; spin < 2000000 && !ok; spin++) {
    uint32_t p;
    asm volatile(
        "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
        : "=r"(p) : "r"(bar_a) : "memory");
    ok = (int)p;
  }
  __syncthreads();
  if (threadIdx.x == 0) flags[0] = ok;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
}

int main() {
  const int rows = 1000, cols = 64;
  std::vector<float> h((size_t)rows * cols);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) h[(size_t)r * cols + c] = r * 1000.f + c;
  float *d, *out;
  int* flags;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 2048 * 4);
  cudaMalloc(&flags, 16);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  if (!enc) { printf("no encoder\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int box_rows : {1, 4}) {
    for (int swz = 0; swz < 2; swz++) {
      const int box_cols = swz ? 32 : 64;
      CUtensorMap map;
      const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
      const cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
      const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
      const cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== box {%d, %d} swizzle %s: encode rc=%d\n", box_cols, box_rows, swz ? "128B" : "none", (int)r);
      if (r != CUDA_SUCCESS) continue;
      for (int oob = 0; oob < 2; oob++) {
        const int r3 = oob ? rows + 5 : 901;
        cudaMemset(flags, 0, 16);
        probe<<<1, 128, 16384>>>(map, 7, 500, 3, r3, 0, (uint32_t)(4 * box_cols * 4), out, flags);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(2048);
        int fl = -1;
        cudaMemcpy(o.data(), out, 2048 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&fl, flags, 4, cudaMemcpyDeviceToHost);
        printf("  rows {7, 500, 3, %d}: launch %s, barrier completed with %d bytes expected: %d\n", r3,
               cudaGetErrorString(e), 4 * box_cols * 4, fl);
        for (int j = 0; j < 4; j++) {
          printf("   smem row %d (pitch %d B): ", j, box_cols * 4);
          for (int c = 0; c < 6; c++) printf("%.0f ", o[j * box_cols + c]);
          printf("... %.0f | next: %.0f\n", o[j * box_cols + box_cols - 1], o[4 * box_cols]);
        }
        if (e != cudaSuccess) return 2;
      }
    }
  }
  return 0;
}
