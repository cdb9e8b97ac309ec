// umma_probe.cu -- standalone check of the tcgen05 building blocks used by
// csrc/row_update_umma.cuh: operand staging layout + smem descriptors + one MMA stage +
// TMEM read-back, against a CPU evaluation of the same bf16-split products.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_probe scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../myrrix-recommender_b200/csrc/umma_common.cuh"

using namespace als;
using namespace als::umma;

template <int KS>
__global__ void __launch_bounds__(256) probe_kernel(const float* __restrict__ Yt, int n_stages,
                                                    float* __restrict__ Dout) {
  using G = StageGeom<KS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  unsigned char* stage0 = smem;  // n_stages * 4096 (+2048 pad)
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  // producers: all 256 threads
  const int q = tid % G::kChunksPerRow, el = tid / G::kChunksPerRow;
  for (int s = 0; s < n_stages; s++) {
    const int e = s * G::kEntries + el;
    float4 v = *reinterpret_cast<const float4*>(Yt + (size_t)e * KS + 4 * q);
    uint2 hi, lo;
    split_bf16x2(v, hi, lo);
    uint32_t oh, ol;
    G::slots(el, q, oh, ol);
    *reinterpret_cast<uint2*>(stage0 + s * G::kBytes + oh) = hi;
    *reinterpret_cast<uint2*>(stage0 + s * G::kBytes + ol) = lo;
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16_mn(G::kM, G::kN);
    uint32_t acc = 0;
    for (int s = 0; s < n_stages; s++) {
      for (int ks = 0; ks < G::kKSteps; ks++) {
        const uint32_t a = smem_u32(stage0 + s * G::kBytes + ks * G::kKStepBytes);
        const uint64_t d = make_smem_desc(a, G::kLBO, G::kSBO);
        mma_bf16_ss(tmem_base, d, d, idesc, acc);
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  if (tid < 128) {
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c = 0; c < G::kN; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_wait_ld();
      for (int j = 0; j < 32; j++) Dout[(size_t)tid * G::kN + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

static float bf16_rn(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  r &= 0xFFFF0000u;
  float y; memcpy(&y, &r, 4); return y;
}

template <int KS>
int run(int n_stages) {
  using G = StageGeom<KS>;
  const int n = n_stages * G::kEntries;
  std::vector<float> Y((size_t)n * KS);
  srand(7);
  for (auto& v : Y) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dY, *dD;
  cudaMalloc(&dY, Y.size() * 4);
  cudaMalloc(&dD, 128 * G::kN * 4);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * G::kN * 4);
  size_t smem = (size_t)n_stages * G::kBytes + 2048;
  cudaFuncSetAttribute(probe_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<KS><<<1, 256, smem>>>(dY, n_stages, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("KS=%d: CUDA error %s\n", KS, cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * G::kN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  // CPU: stacked[m] for m in [0,2KS): hi features then lo features
  double max_err = 0, max_ref = 0, max_full = 0;
  for (int m = 0; m < 2 * KS; m++)
    for (int c = 0; c < 2 * KS; c++) {
      double ref = 0;
      for (int t = 0; t < n; t++) {
        auto part = [&](int idx) {
          float x = Y[(size_t)t * KS + idx % KS];
          float hi = bf16_rn(x);
          return idx < KS ? hi : bf16_rn(x - hi);
        };
        ref += (double)part(m) * (double)part(c);
      }
      double got = D[(size_t)m * G::kN + c];
      max_err = fmax(max_err, fabs(got - ref));
      max_ref = fmax(max_ref, fabs(ref));
    }
  // full-precision check of the recombined SYRK: W[i][j] = D[i][j]+D[i][KS+j]+D[KS+i][j]+D[KS+i][KS+j]
  for (int i = 0; i < KS; i++)
    for (int j = 0; j < KS; j++) {
      double ref = 0;
      for (int t = 0; t < n; t++) ref += (double)Y[(size_t)t * KS + i] * (double)Y[(size_t)t * KS + j];
      double got = (double)D[(size_t)i * G::kN + j] + D[(size_t)i * G::kN + KS + j] +
                   D[(size_t)(KS + i) * G::kN + j] + D[(size_t)(KS + i) * G::kN + KS + j];
      max_full = fmax(max_full, fabs(got - ref));
    }
  printf("KS=%d stages=%d entries=%d: max|D - ref| = %.3e (max|ref| %.3e); recombined SYRK abs err %.3e\n",
         KS, n_stages, n, max_err, max_ref, max_full);
  return max_err <= 1e-3 * max_ref ? 0 : 2;
}

int main() {
  int rc = 0;
  rc |= run<64>(1);
  rc |= run<64>(7);
  rc |= run<32>(1);
  rc |= run<32>(4);
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  return rc;
}
