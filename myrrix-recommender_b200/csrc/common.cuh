// common.cuh -- shared device/host helpers for the ALS core (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/myrrix_als.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "this library is written for sm_100a (B200) only"
#endif

namespace als {

constexpr int kMaxFeatures = 128;  // largest padded feature count any kernel supports
constexpr int kWarp = 32;

// Device-side deferred error record (one per handle, in HBM). Kernels never trap:
// a singular / non-finite row sets the flag and the host reports it at the next sync,
// mirroring how the reference surfaces SingularMatrixSolverException out of a worker
// Future (AlternatingLeastSquares.java:348-349).
struct DeviceStatus {
  int code;            // als_status, first error wins
  int which;           // 0 = X half, 1 = Y half
  long long row;       // offending row (dense index)
  float pivot;         // the pivot that failed
};

__device__ __forceinline__ void report_error(DeviceStatus* st, int code, int which, long long row,
                                             float pivot) {
  if (atomicCAS(&st->code, 0, code) == 0) {
    st->which = which;
    st->row = row;
    st->pivot = pivot;
  }
}

// 128-bit read-only gather of factor rows: goes through L1/L2 (rows are re-used
// across CTAs when the opposite factor fits in the 126 MB L2).
// (volatile asm: the compiler must issue the load where it is written -- all of a stage's
// gathers up front -- and not sink it into the conditional block that consumes it)
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Streaming (read-once) loads for the interaction arrays: do not pollute L1.
__device__ __forceinline__ int ld_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Ask L2 for the 128-byte line holding p (per-lane, nothing held while it is in flight).
__device__ __forceinline__ void l2_prefetch_line(const float* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}

// Packed fp32x2 FMA (sm_100+): d = a*b + c on two lanes of a 64-bit register pair.
// B200's fp32 pipe reaches full rate only with the packed form.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}

__host__ __device__ constexpr int round_up_int(int a, int b) { return (a + b - 1) / b * b; }

// Padded feature count (row stride of the device factor matrices).
inline int padded_features(int k) {
  int ks = 4;
  while (ks < k) ks <<= 1;
  return ks;
}

}  // namespace als
