// umma_common.cuh -- sm_100a tcgen05 / TMEM / mbarrier primitives (inline PTX) and the
// shared-memory operand layout used by the tensor-core row update.
//
// Operand layout (MN-major, SWIZZLE_128B, bf16): one "atom" is 8 K-rows x 128 bytes
// (= 64 bf16 along the MN dimension), 1024 bytes, 1024-byte aligned; inside an atom the
// 16-byte chunk index is XORed with the K-row index (Swizzle<3,4,3>).  Atoms along MN
// are LBO bytes apart, atoms along K are SBO bytes apart.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace als {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
#ifdef ALS_MBAR_NOHINT
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#endif
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)  // suspend-time hint (ns): sleep, don't spin
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Optional wait-time accounting (build with -DALS_PROFILE_WAITS): cycles each role spends
// blocked on each kind of barrier, summed over all warps/CTAs into g_wait_cycles[id].
#ifdef ALS_PROFILE_WAITS
__device__ unsigned long long g_wait_cycles[16];
__device__ __forceinline__ void mbar_wait_id(uint64_t* bar, uint32_t parity, int id) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  const long long dt = clock64() - t0;
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_wait_cycles[id], (unsigned long long)dt);
}
#elif defined(ALS_WATCHDOG)
// development builds: a wait that lasts longer than ~2 s reports who is stuck on what and traps
__device__ __forceinline__ void mbar_wait_id(uint64_t* bar, uint32_t parity, int id) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if ((threadIdx.x & 31) == 0)
        printf("WATCHDOG cta %d warp %d wait id %d bar +%u parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), id,
               (unsigned)(smem_u32(bar)), parity);
      __trap();
    }
  }
}
#else
// Waits of the roles that run AHEAD of the pipeline (a full ring, no free accumulator / slot:
// ids in ALS_BACKOFF_IDS) back off with an explicit sleep between polls: try_wait's own suspend
// ends at every mbarrier event of the CTA (dozens of polls per wait), and those polls take issue
// slots from the warps the waiter is waiting for.
#ifndef ALS_BACKOFF_NS
#define ALS_BACKOFF_NS 0
#endif
#ifndef ALS_BACKOFF_IDS
#define ALS_BACKOFF_IDS ((1u << 0) | (1u << 1) | (1u << 2) | (1u << 4))  // b_empty, empty, acc_empty, w_empty
#endif
__device__ __forceinline__ void mbar_wait_id(uint64_t* bar, uint32_t parity, int id) {
  if (ALS_BACKOFF_NS > 0 && ((ALS_BACKOFF_IDS >> id) & 1u)) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ALS_BACKOFF_NS);
  } else {
    mbar_wait(bar, parity);
  }
}
#endif

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- named barrier over a subset of the CTA's warps ---------------------------
__device__ __forceinline__ void bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---- register re-balancing between warp roles (whole warpgroups only) ---------------
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- TMEM --------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) columns; base address lands in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of the warp receives lane (base_lane + t),
// columns [col, col+32) in v[0..31]. taddr = tmem_base + (base_lane << 16) + col.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---- tcgen05.mma ---------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1
//   [49,52) base offset = 0 | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32,
// both operands MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4)      // c_format = F32
         | (1u << 7)    // a_format = BF16
         | (1u << 10)   // b_format = BF16
         | (1u << 15)   // a_major = MN
         | (1u << 16)   // b_major = MN
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Warp-converged variants: every lane executes the call, one elected lane issues (the same
// lane every time for a full, converged warp).
__device__ __forceinline__ void mma_bf16_ss_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A = B = the same operand tile: one descriptor (low word dlo, high word dhi).
__device__ __forceinline__ void mma_bf16_ss_same_elect(uint32_t tmem_d, uint32_t dlo, uint32_t dhi,
                                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 d;\n\t"
      "mov.b64 d, {%1, %2};\n\t"
#ifdef ALS_MMA_LANE0
      "setp.eq.b32 q, 0, 0;\n\t"
#else
      "elect.sync _|q, 0xffffffff;\n\t"
#endif
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], d, d, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(dlo), "r"(dhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_addr_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
#ifdef ALS_MMA_LANE0
      "setp.eq.b32 q, 0, 0;\n\t"
#else
      "elect.sync _|q, 0xffffffff;\n\t"
#endif
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(bar_addr)
      : "memory");
}
// Wait with a bounded poll rate: NS > 0 sleeps about NS nanoseconds between polls.  Every poll of an
// mbarrier is a shared-memory access: with the plain wait the nine consumer warps of the long-row
// mix, which wait most of the time, polled away a quarter of the SM's shared-memory bandwidth -- the
// resource that bounds that kernel.
template <int NS>
__device__ __forceinline__ void mbar_wait_paced(uint64_t* bar, uint32_t parity) {
  if constexpr (NS > 0) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(NS);
  } else {
    mbar_wait(bar, parity);
  }
}
template <int NS>
__device__ __forceinline__ void mbar_wait_addr_paced(uint32_t bar_addr, uint32_t parity) {
  if constexpr (NS > 0) {
    for (;;) {
      uint32_t ok;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(ok)
          : "r"(bar_addr), "r"(parity)
          : "memory");
      if (ok) break;
      __nanosleep(NS);
    }
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}\n" ::"r"(bar_addr), "r"(parity), "r"(1000000u)
        : "memory");
  }
}

// spin on an mbarrier given by its shared-window address
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
#ifdef ALS_MBAR_NOHINT
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#endif
      "@!p bra WAIT_%=;\n\t"
      "}\n" ::"r"(bar_addr), "r"(parity), "r"(1000000u)  // suspend-time hint (ns): sleep, don't spin
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
#ifdef ALS_MMA_LANE0
      "setp.eq.b32 q, 0, 0;\n\t"
#else
      "elect.sync _|q, 0xffffffff;\n\t"
#endif
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- operand staging -----------------------------------------------------------------
// Split v (4 consecutive features of one scaled factor row) into bf16 hi + bf16 lo
// (v ~= hi + lo to ~2^-17 relative, round-to-nearest both times) and store both halves.
// atom_row points at the 128-byte K-row inside the hi atom / lo atom; `chunk` is the
// logical 16-byte chunk, `half` selects the 8-byte half of it; krow the K-row (0..7).
__device__ __forceinline__ void split_bf16x2(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
  const __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f01 = __bfloat1622float2(h01);
  const float2 f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
  hi.x = *reinterpret_cast<const uint32_t*>(&h01);
  hi.y = *reinterpret_cast<const uint32_t*>(&h23);
  lo.x = *reinterpret_cast<const uint32_t*>(&l01);
  lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

// Geometry of one operand stage for padded feature count KS (32 or 64).
//   KS=64: one K-step (16 entries) per stage: K-atoms kb=0,1; two MN-atoms of 64 operand rows.
//          atom(kb, mb) at (kb*2 + mb)*1024.  A = B : M = N = 128, LBO 1024, SBO 2048.  Operand
//          row of feature f: hi at (f/16)*32 + f%16, lo at +16 (groups [hi16 | lo16]; the two
//          groups of the second atom are stored [lo16 | hi16] to spread the stores over banks), so
//          TMEM lane quarter g / column group g of D hold the hi and lo parts of features
//          16g..16g+15 side by side.
//   KS=32: two K-steps (2 x 16 entries) per stage; one MN-atom whose 128-byte row is
//          [hi(32) | lo(32)]; atom(ks, kb) at ks*2048 + kb*1024.  N = 64, M = 128 (rows
//          64..127 of D are don't-care), SBO 1024, LBO 1024.
template <int KS>
struct StageGeom {
  static_assert(KS == 32 || KS == 64, "tcgen05 path supports padded feature counts 32 and 64");
  static constexpr int kChunksPerRow = KS / 4;                 // float4 per factor row
  static constexpr int kEntries = 256 / kChunksPerRow;         // entries per stage (256 producers)
  static constexpr int kKSteps = kEntries / 16;                // MMA K-steps per stage
  static constexpr int kBytes = 4096;
  static constexpr int kM = 128;
  static constexpr int kN = 2 * KS;
  static constexpr uint32_t kLBO = 1024;
  static constexpr uint32_t kSBO = (KS == 64) ? 2048 : 1024;
  static constexpr uint32_t kKStepBytes = (KS == 64) ? 4096 : 2048;

  // byte offsets (within the stage) of the 8-byte slots where producer (entry slot el,
  // chunk q) stores its hi and lo halves
  __device__ static __forceinline__ void slots(int el, int q, uint32_t& off_hi, uint32_t& off_lo) {
    const int krow = el & 7;
    if (KS == 64) {
      // M-row of feature f: hi at (f/16)*32 + f%16, lo 16 rows further, so that every TMEM
      // lane quarter (32 M-rows) holds the hi AND lo halves of the same 16 features
      const int kb = (el >> 3) & 1;
      const int g16 = q >> 2, r = q & 3;            // 16-feature group, 4-feature chunk inside it
      // 16-byte chunk inside the 128-byte atom row.  The second MN-atom swaps the hi and lo
      // chunk pairs ([lo16 | hi16] groups): the 16 lanes that store one entry's hi (or lo)
      // halves then cover all 32 banks, and hi + lo is symmetric for the drain.
      const int chunk = (g16 & 1) * 4 + (r >> 1) + ((g16 >> 1) ? 2 : 0);
      const uint32_t row = (uint32_t)(kb * 2 + (g16 >> 1)) * 1024u + (uint32_t)krow * 128u;
      off_hi = row + (uint32_t)((chunk ^ krow) * 16 + (r & 1) * 8);
      off_lo = off_hi ^ 32u;                        // chunk ^ 2
    } else {
      const int ks = (el >> 4) & 1, kb = (el >> 3) & 1;
      const uint32_t row = (uint32_t)ks * 2048u + (uint32_t)kb * 1024u + (uint32_t)krow * 128u;
      off_hi = row + (uint32_t)(((q >> 1) ^ krow) * 16 + (q & 1) * 8);
      off_lo = row + (uint32_t)(((4 + (q >> 1)) ^ krow) * 16 + (q & 1) * 8);
    }
  }
};

}  // namespace umma
}  // namespace als
