// gather_probe.cu -- what does HBM deliver for the row update's access pattern: random 4 * K-byte rows of a
// matrix much larger than L2, indices streamed?  (The Y<-X half of C3 gathers 1e9 random 256-byte rows of a
// 2.56 GB matrix.)  Plain loads, 16 bytes per lane, eight independent rows in flight per thread group,
// grid of 148 x 8 CTAs; also cp.async.bulk-free: this is the roof any gather path has to live under.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/gather_probe scripts/gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int K>
__global__ void gather(const float* __restrict__ M, const int* __restrict__ idx, long long n, float* out) {
  constexpr int LPR = K / 4;          // lanes per row (16 bytes each)
  constexpr int RPW = 32 / LPR;       // rows per warp instruction
  const int lane = threadIdx.x & 31, q = lane % LPR, sub = lane / LPR;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long base = warp * (8 * RPW); base < n; base += nwarps * (8 * RPW)) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long e = base + u * RPW + sub;
      const int r = e < n ? __ldg(idx + e) : 0;
      v[u] = __ldg(reinterpret_cast<const float4*>(M + (long long)r * K) + q);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void fill_idx(int* idx, long long n, int rows, unsigned long long seed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    unsigned long long x = seed + e * 0x9E3779B97F4A7C15ULL;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
    idx[e] = (int)(x % (unsigned long long)rows);
  }
}

template <int K>
void run(int rows, long long n) {
  float *M, *out;
  int* idx;
  cudaMalloc(&M, (size_t)rows * K * 4);
  cudaMalloc(&idx, (size_t)n * 4);
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMemset(M, 0, (size_t)rows * K * 4);
  fill_idx<<<1184, 256>>>(idx, n, rows, 1234567890ULL);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a);
    gather<K><<<148 * 8, 256>>>(M, idx, n, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep == 2)
      printf("k=%d rows=%d (%.2f GB) gathers=%lld: %.3f ms, %.0f GB/s of row bytes (+ %.0f GB/s of indices)\n", K, rows,
             (double)rows * K * 4 / 1e9, n, ms, (double)n * K * 4 / ms / 1e6, (double)n * 4 / ms / 1e6);
  }
  cudaFree(M); cudaFree(idx); cudaFree(out);
}

int main() {
  run<64>(10000000, 250000000LL);   // the Y<-X half of C3: rows of X (2.56 GB)
  run<64>(1000000, 250000000LL);    // the X<-Y half: rows of Y (256 MB, half L2-resident)
  run<32>(1000000, 250000000LL);    // C2's Y<-X half: rows of X (128 MB)
  run<128>(10000000, 125000000LL);  // 512-byte rows
  return 0;
}
