// row_update_v2.cuh -- tensor-core (tcgen05 + TMEM) row update, second generation (k = 64).
//
// Same contract as row_update_simt.cuh (Worker.call, AlternatingLeastSquares.java:438-502) and
// the same four warp-specialised roles and operand layout as row_update_umma.cuh (round 1); what
// changed is the instruction budget of every role (round-1 profile: 12.3k issued warp
// instructions per solved row, issue slots and the shared-memory pipe saturated long before
// HBM):
//
//   producers   (7 or 11 warps) own every P-th 16-entry stage of the CTA's flat stage stream.
//               GATHER: asynchronous 16-byte copies (cp.async: no register staging, zero-fill
//               beyond a row's last entry) of the fp32 factor rows straight into the operand
//               ring slot of the stage, completion signalled on the slot's mbarrier
//               (cp.async.mbarrier.arrive); every producer keeps the gathers of its next kAhead
//               stages in flight while it converts the current one, so the ring is the in-flight
//               buffer: ~55-90 KB of gathers outstanding per SM with no registers tied up (round
//               1 held 7-11 warps x 4 KB in registers and was latency-bound at 1.5 / 4 TB/s).
//               CONVERT in place: read the raw fp32 rows back from the slot (16-byte loads,
//               conflict-free), scale by sqrt(alpha |r|), split into bf16 hi + bf16 lo, write the
//               swizzled MN-major operand tile into the same 4 KB.  Stage lookup is one ballot
//               over a per-batch prefix sum of the rows' stage counts (no cursor walk, no integer
//               modulo); per-entry scalars once per entry on one lane, handed to the 16 lanes of
//               the entry by shuffle; swizzled slots addressed as (lane base) ^ (per-pass
//               constant).  Each producer keeps the partial rhs of ITS stages of a row in
//               registers and writes it once, when it leaves the row.
//               Measured alternatives: a dedicated loader warp issuing the same
//               copies (one warp cannot keep enough of them in flight: 1.9 TB/s), and per-row TMA
//               bulk copies (cp.async.bulk takes its addresses from uniform registers, so 32
//               divergent rows serialise into 32 elect / R2UR round trips per warp: 0.9 TB/s).
//               TMA bulk copies are used where addresses are warp-uniform: the finished factor
//               rows (Cholesky epilogue).  -DALS_V2_TMA=0 builds the register-gather producers
//               (LDG.128 + the same conversion) for A/B runs.
//   MMA issuer  unchanged (one tcgen05.mma per stage, D += [hi;lo][hi;lo]^T); it also signals
//               "rhs complete" for a row after it has seen the row's last stage.
//   drain       reads the accumulator with tcgen05.ld.16x256b: the register layout is the
//               mma.sync accumulator fragment, the hi and lo operand rows of one matrix row
//               arrive in the SAME thread (two loads, 16 TMEM lanes apart): no shuffles.  Writes
//               N = -(G + lambda alpha n_u I + D) as 16 x 16 blocks into the panel-major slot.
//   Cholesky    chol_blocked.cuh: 16-column panels on the CUDA cores, trailing updates as
//               3xTF32 mma.sync on fragments of the slot, in place.
#pragma once
#include <cuda.h>  // CUtensorMap (types only)

#include "chol_blocked.cuh"
#include "common.cuh"
#include "row_update_simt.cuh"  // RowUpdateParams
#include "umma_common.cuh"

namespace als {
namespace v2 {

using umma::bar_sync;
using umma::fence_proxy_async_smem;
using umma::mbar_arrive;
using umma::mbar_init;
using umma::mbar_init_fence;
using umma::mbar_wait_addr;
using umma::mbar_wait_id;
using umma::smem_u32;
using umma::tc_fence_after_sync;
using umma::tc_fence_before_sync;

constexpr int kDrainWarps = 4;                 // warps 0..3 (TMEM lane quarters)
constexpr int kFirstChol = kDrainWarps;
constexpr int kAccSlots = 4;
constexpr int kSegStages = 64;
constexpr int kTmemCols = 512;
constexpr int kRegsDrain = 64;
#ifndef ALS_V2_REGS_CHOL
#define ALS_V2_REGS_CHOL 128
#endif
constexpr int kRegsChol = ALS_V2_REGS_CHOL;
constexpr float kCondLimit = 256.f;  // max diag / min pivot above which a row goes to fp64
constexpr unsigned kFull = 0xffffffffu;
// Cholesky warps enter each sweep together in groups (one instruction stream per group for the
// ~50 KB of straight-line solver code): 1 = all warps of the CTA, 2 / 4 = that many staggered
// groups (while one group solves, the drain refills the slots of the other), 0 = no grouping.
#ifndef ALS_V2_LOCKSTEP
#define ALS_V2_LOCKSTEP 2
#endif
// W slots beyond one per Cholesky warp (8-warp mix): the solve runs in place, so a slot is busy
// for the whole solve; spare slots let the drain work ahead of the oldest unfinished solve.
#ifndef ALS_V2_XSLOTS
#define ALS_V2_XSLOTS 0
#endif

// 1: the asynchronous gathers are TMA tile::gather4 loads (cp.async.bulk.tensor.2d ... tile::gather4: four
// factor rows per instruction, written into the ring slot by the TMA engine, bytes counted on the slot's
// mbarrier; indices beyond a row's last entry point past the tensor and are zero-filled);
// 0: Ampere-style cp.async (LDGSTS) copies, 16 bytes per lane
#ifndef ALS_V2_GATHER4
#define ALS_V2_GATHER4 1
#endif
constexpr bool kGather4 = ALS_V2_GATHER4 != 0;
constexpr int kOobRow = 0x40000000;  // a row index beyond any factor matrix: TMA fills zeros

// four rows {r0..r3} of the tensor (one box row each: KS floats from column 0) -> dst .. dst + 4 * KS * 4
__device__ __forceinline__ void tma_gather4(uint32_t dst_saddr, const CUtensorMap* map, int r0, int r1, int r2, int r3,
                                            uint32_t mbar_saddr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(mbar_saddr)
      : "memory");
}

#ifndef ALS_V2_PACE_CHOL_NS
#define ALS_V2_PACE_CHOL_NS 0
#endif
#ifndef ALS_V2_PACE_DRAIN_NS
#define ALS_V2_PACE_DRAIN_NS 0
#endif
#ifndef ALS_V2_PACE_MMA_NS
#define ALS_V2_PACE_MMA_NS 0
#endif
#ifndef ALS_V2_PACE_CHOL_SHORT_NS
#define ALS_V2_PACE_CHOL_SHORT_NS 0
#endif
#ifndef ALS_V2_PACE_DRAIN_SHORT_NS
#define ALS_V2_PACE_DRAIN_SHORT_NS 0
#endif
#ifndef ALS_V2_PACE_MMA_SHORT_NS
#define ALS_V2_PACE_MMA_SHORT_NS 0
#endif

// extra own-stages of look-ahead for a producer's index / value loads (see the producer loop)
#ifndef ALS_V2_FETCH_EXTRA
#define ALS_V2_FETCH_EXTRA 0
#endif
constexpr int kFetchExtra = ALS_V2_FETCH_EXTRA;

// 1: hand-offs signalled by one elected lane per warp (after the warp's fences + __syncwarp) instead
// of one mbarrier arrival per lane: 32x fewer SYNCS operations on `full`, `acc_empty`, `w_full`
#ifndef ALS_V2_ARRIVE1
#define ALS_V2_ARRIVE1 0
#endif
constexpr bool kArrive1 = ALS_V2_ARRIVE1 != 0;

// 1: asynchronous gathers into the ring by a loader warp, in-place conversion; 0: register gathers (LDG)
#ifndef ALS_V2_TMA
#define ALS_V2_TMA 1
#endif
constexpr bool kTma = ALS_V2_TMA != 0;

#ifdef ALS_DEBUG_SLOT
__device__ long long g_debug_row = -1;
__device__ float g_debug_slot[WPanels<64>::kFloats + 64];
#endif

// Role mix of one CTA: 4 drain warps | NCHOL Cholesky warps | NPROD producer warps | 1 MMA warp.
//   STAGES  operand ring depth (4 KB each); with asynchronous gathers the ring is also the
//           in-flight buffer: (AHEAD + 1) * NPROD slots are in flight or being converted
//   AHEAD   own-stages whose gathers a producer keeps in flight
//   SETREG  re-balance registers between the roles with setmaxnreg (needs whole warpgroups per
//           role); otherwise every role lives within the launch register count
//   ASYNC   producers gather with asynchronous copies into the ring and convert in place;
//           otherwise they gather into registers (LDG.128) and store the converted tile once
//           (fewer shared-memory wavefronts, but the gathers are latency-bound: measured slower on
//           both halves of the headline workload, 154 vs 143 ms and 68 vs 61 ms)
template <int NCHOL, int NPROD, int STAGES, int AHEAD, bool SETREG, bool ASYNC>
struct Mix {
  static constexpr bool kAsync = ASYNC && kTma;
  static constexpr int kCholWarps = NCHOL;
  static constexpr int kProdWarps = NPROD;
  static constexpr int kFirstProd = kFirstChol + NCHOL;
  static constexpr int kMmaWarp = kFirstProd + NPROD;
  static constexpr int kThreads = 32 * (kMmaWarp + 1);
  static constexpr int kStages = STAGES;
  static constexpr int kAhead = AHEAD;
  static constexpr int kWSlots = NCHOL;                 // the solve runs in place: one slot per Cholesky warp
  // rhs slots, row u -> slot u % kBSlots.  One per Cholesky warp: a slot's consecutive phases are
  // then waited for by the SAME warp in order (a parity wait by a different warp could run a whole
  // phase early and pass on the stale parity -- measured: deadlock with 8 slots and 10 warps).
  static constexpr int kBSlots = NCHOL;
  static constexpr bool kSetReg = SETREG;
  // poll pacing (ns between polls) of the consumer roles, which wait most of the time when the
  // producers bound the kernel (long rows); 0: plain wait
  static constexpr int kPaceChol = NCHOL == 4 ? ALS_V2_PACE_CHOL_NS : ALS_V2_PACE_CHOL_SHORT_NS;
  static constexpr int kPaceDrain = NCHOL == 4 ? ALS_V2_PACE_DRAIN_NS : ALS_V2_PACE_DRAIN_SHORT_NS;
  static constexpr int kPaceMma = NCHOL == 4 ? ALS_V2_PACE_MMA_NS : ALS_V2_PACE_MMA_SHORT_NS;
  // register budget per role (setmaxnreg): 640 threads launch at 96 registers -- drain 64, Cholesky 128,
  // producers 80 (8 + 7 mix) or 96; 768 threads (4 + 15 mix) launch at 80 -- drain 64, Cholesky 96, producers 80
  static constexpr int kLaunchRegs = kThreads == 640 ? 96 : 80;
  static constexpr int kRegsProd = (kThreads == 640 && NCHOL != 8) ? 96 : 80;
  static constexpr int kRegsCholMix = kThreads == 640 ? kRegsChol : 96;
  static_assert(kThreads <= 1024, "threads per CTA");
  static_assert(!SETREG || (NCHOL % 4 == 0 && (NPROD + 1) % 4 == 0 && (kThreads == 640 || kThreads == 768)),
                "setmaxnreg: warpgroups per role");
  static_assert(!SETREG || 128 * (kRegsDrain + (NCHOL / 4) * kRegsCholMix + ((NPROD + 1) / 4) * kRegsProd) <=
                               kThreads * kLaunchRegs,
                "register pool");
  static_assert(!kAsync || kStages >= (kAhead + 1) * kProdWarps, "ring too shallow for the gather depth");
};
// short rows (solve-bound) / long rows (gather-bound); see launch_row_update_v2
// (development overrides: -DALS_V2_S_NCHOL=.. -DALS_V2_S_NPROD=.. -DALS_V2_S_STAGES=.. -DALS_V2_S_AHEAD=..
// -DALS_V2_S_SETREG=0|1 for the short-row mix, ALS_V2_L_* for the long-row mix)
#ifndef ALS_V2_S_NCHOL
#define ALS_V2_S_NCHOL 8
#define ALS_V2_S_NPROD 7
#define ALS_V2_S_STAGES 24
#define ALS_V2_S_AHEAD 2
#define ALS_V2_S_SETREG 1
#endif
#ifndef ALS_V2_S_ASYNC
#define ALS_V2_S_ASYNC 1
#endif
#ifndef ALS_V2_L_ASYNC
#define ALS_V2_L_ASYNC 1
#endif
#ifndef ALS_V2_L_NCHOL
#define ALS_V2_L_NCHOL 4
#define ALS_V2_L_NPROD 11
#define ALS_V2_L_STAGES 34
#define ALS_V2_L_AHEAD 2
#define ALS_V2_L_SETREG 1
#endif
using MixShort = Mix<ALS_V2_S_NCHOL, ALS_V2_S_NPROD, ALS_V2_S_STAGES, ALS_V2_S_AHEAD, ALS_V2_S_SETREG != 0, ALS_V2_S_ASYNC != 0>;
using MixLong = Mix<ALS_V2_L_NCHOL, ALS_V2_L_NPROD, ALS_V2_L_STAGES, ALS_V2_L_AHEAD, ALS_V2_L_SETREG != 0, ALS_V2_L_ASYNC != 0>;
// k = 32: a solve is 4x cheaper and its slot 3 KB instead of 10 KB; the mixes are tuned separately
// (ALS_V2_S32_* / ALS_V2_L32_*)
// (short rows, measured on config 2, X-half: 8 + 7 warps with setmaxnreg 5.3 ms; 12 + 7 warps at the
// launch register count 4.9 ms; 16 + 7: 4.9 ms; 8 + 11 or 12 + 11: 5.1-5.3 ms)
#ifndef ALS_V2_S32_NCHOL
#define ALS_V2_S32_NCHOL 12
#define ALS_V2_S32_NPROD 7
#define ALS_V2_S32_STAGES 24
#define ALS_V2_S32_AHEAD 2
#define ALS_V2_S32_SETREG 0
#endif
#ifndef ALS_V2_L32_NCHOL
#define ALS_V2_L32_NCHOL ALS_V2_L_NCHOL
#define ALS_V2_L32_NPROD ALS_V2_L_NPROD
#define ALS_V2_L32_STAGES ALS_V2_L_STAGES
#define ALS_V2_L32_AHEAD ALS_V2_L_AHEAD
#define ALS_V2_L32_SETREG ALS_V2_L_SETREG
#endif
template <int KS>
struct Mixes {
  using Short = MixShort;
  using Long = MixLong;
};
template <>
struct Mixes<32> {
  using Short = Mix<ALS_V2_S32_NCHOL, ALS_V2_S32_NPROD, ALS_V2_S32_STAGES, ALS_V2_S32_AHEAD, ALS_V2_S32_SETREG != 0, true>;
  using Long = Mix<ALS_V2_L32_NCHOL, ALS_V2_L32_NPROD, ALS_V2_L32_STAGES, ALS_V2_L32_AHEAD, ALS_V2_L32_SETREG != 0, true>;
};

template <int KS, class MX>
struct Smem {
  using WP = WPanels<KS>;
  static constexpr int NCHOL = MX::kCholWarps;
  static constexpr size_t kRing = (size_t)MX::kStages * 4096 + (KS == 32 ? 2048 : 0);  // +pad: KS = 32 A-operand over-read
  static constexpr size_t kSlotBytes = sizeof(float) * WP::kFloats;
  static constexpr size_t off_slots = kRing;                                    // [NCHOL] W slots
  static constexpr size_t off_ng = off_slots + MX::kWSlots * kSlotBytes;        // -G, panel layout
  static constexpr size_t off_bpart = off_ng + kSlotBytes;                      // [kBSlots][P][KS]
  static constexpr size_t off_scratch = off_bpart + sizeof(float) * MX::kBSlots * MX::kProdWarps * KS;
  static constexpr size_t off_bars = off_scratch + sizeof(float) * NCHOL * CholBlocked<KS>::kScratch;
  static constexpr int kNumBars = 3 * MX::kStages + 2 * kAccSlots + 2 * MX::kWSlots + 2 * MX::kBSlots;
  static constexpr size_t off_misc = (off_bars + sizeof(uint64_t) * kNumBars + 15) / 16 * 16;
  static constexpr size_t kTotal = off_misc + 64;
  static_assert(kTotal <= 227 * 1024, "shared memory per CTA");
};

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ long long shfl_i64(long long v, int src) {
  int lo = __shfl_sync(kFull, (int)(v & 0xffffffffLL), src);
  int hi = __shfl_sync(kFull, (int)(v >> 32), src);
  return ((long long)hi << 32) | (unsigned int)lo;
}
// 16 TMEM lanes x 32 columns of 32-bit in the mma.sync accumulator-fragment layout: for each
// group i of 8 columns, thread (g = lane/4, t = lane%4) receives v[4i+0..1] = row g, columns
// 8i+2t, 8i+2t+1 and v[4i+2..3] = row g+8, same columns.  taddr lane field: first of the 16 lanes.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}

// TMA bulk copy global -> shared (one contiguous run of `bytes`, 16-byte aligned both sides), its
// completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t mbar_saddr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_saddr), "l"(src), "r"(bytes), "r"(mbar_saddr)
               : "memory");
}
// Ampere-style asynchronous 16-byte copy global -> shared (no register staging); src_bytes = 0
// zero-fills the destination instead of reading
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst_saddr, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all earlier cp.async of this thread are done
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_saddr, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_saddr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Operand stage geometry (one 4 KB ring slot = 256 producer chunks of 16 bytes):
//   KS = 64: 16 entries, one K-step; two MN-atoms of 64 operand rows (umma::StageGeom<64>).
//   KS = 32: 32 entries, two K-steps; one MN-atom of 64 operand rows.
// Operand row of feature f in both: hi at (f/16)*32 + f%16, lo 16 rows further, so TMEM lane
// quarter q / column group q of D hold the hi and lo parts of features 16q..16q+15 side by side.
template <int KS>
struct Geom;
template <>
struct Geom<64> : umma::StageGeom<64> {
  // producer pass ps (entries 2 ps + sub): address relative to pass 0 = ^xo, then +ko
  __device__ static constexpr uint32_t pass_xor(int ps) { return (uint32_t)(((ps & 3) << 5) | ((ps & 3) << 8)); }
  __device__ static constexpr uint32_t pass_add(int ps) { return (uint32_t)((ps >> 2) << 11); }
};
template <>
struct Geom<32> {
  static constexpr int kChunksPerRow = 8;
  static constexpr int kEntries = 32;
  static constexpr int kKSteps = 2;
  static constexpr int kBytes = 4096;
  static constexpr int kM = 128;  // (rows 64..127 of D are don't-care: the A operand over-reads into the next K-atom)
  static constexpr int kN = 64;
  static constexpr uint32_t kLBO = 1024;
  static constexpr uint32_t kSBO = 1024;
  static constexpr uint32_t kKStepBytes = 2048;
  // atom(ks, kb) at ks*2048 + kb*1024; K-row = entry & 7; the 128-byte row is
  // [hi16 | lo16 | hi16 | lo16] of features 0-15, 16-31
  __device__ static __forceinline__ void slots(int el, int q, uint32_t& off_hi, uint32_t& off_lo) {
    const int krow = el & 7, kb = (el >> 3) & 1, ks = (el >> 4) & 1;
    const int g16 = q >> 2, r = q & 3;
    const int chunk = 4 * g16 + (r >> 1);
    const uint32_t row = (uint32_t)ks * 2048u + (uint32_t)kb * 1024u + (uint32_t)krow * 128u;
    off_hi = row + (uint32_t)((chunk ^ krow) * 16 + (r & 1) * 8);
    off_lo = off_hi ^ 32u;  // chunk ^ 2
  }
  // producer pass ps (entries 4 ps + sub): K-row 4 (ps & 1) + sub of K-atom (ps >> 1) & 1, K-step ps >> 2
  __device__ static constexpr uint32_t pass_xor(int ps) { return (uint32_t)(((ps & 1) << 6) | ((ps & 1) << 9)); }
  __device__ static constexpr uint32_t pass_add(int ps) { return (uint32_t)((((ps >> 1) & 1) << 10) + ((ps >> 2) << 11)); }
};

// Row table of a batch of 32 rows (rows rb, rb + step, ...): one row per lane.
struct RowBatch {
  long long e0;     // first entry of my row
  int cnt;          // entries of my row
  uint32_t end;     // flat index one past my row's last stage (inclusive prefix sum + base)
  uint32_t ne_mask; // ballot of non-empty rows
  uint32_t batch_end;
  int nb;           // rows in this batch
};
template <int E>
__device__ __forceinline__ void load_batch(const RowUpdateParams& p, long long rb, long long row_step,
                                           int lane, uint32_t base, RowBatch& B) {
  B.e0 = 0;
  B.cnt = 0;
  const long long mypos = rb + lane * row_step;  // position in the visiting order
  if (mypos < p.n_rows) {
    const long long myrow = p.row_order ? (long long)p.row_order[mypos] : mypos;
    B.e0 = p.row_ptr[myrow];
    B.cnt = (int)(p.row_ptr[myrow + 1] - B.e0);
  }
  const long long left = (p.n_rows - rb + row_step - 1) / row_step;
  B.nb = left < 32 ? (int)left : 32;
  uint32_t incl = (uint32_t)((B.cnt + E - 1) / E);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  B.end = base + incl;
  B.batch_end = __shfl_sync(kFull, B.end, 31);
  B.ne_mask = __ballot_sync(kFull, B.cnt > 0);
}

// kExt: the variant that walks virtual rows (long rows split into chunks) and / or runs in stash mode
// (RowUpdateParams::vrow, ::stash).  The plain variant carries none of that code: its solving warps
// have no registers to spare for state that uniform workloads never use.
template <int KS, class MX, bool kExt = false>
__global__ void __launch_bounds__(MX::kThreads, 1) row_update_v2_kernel(const RowUpdateParams p,
                                                                        const __grid_constant__ CUtensorMap tmapM) {
  static_assert(KS == 64 || KS == 32, "second-generation kernel: k = 32 or 64");
  static_assert(KS == 64 || MX::kAsync, "register-gather producers are kept for k = 64 A/B builds only");
  using G = Geom<KS>;
  using S = Smem<KS, MX>;
  using WP = WPanels<KS>;
  using CB = CholBlocked<KS>;
  constexpr int kCholWarps = MX::kCholWarps, P = MX::kProdWarps, kFirstProd = MX::kFirstProd;
  constexpr int kStages = MX::kStages, kWSlots = MX::kWSlots, kBSlots = MX::kBSlots;
  constexpr int kMmaWarp = MX::kMmaWarp, kThreads = MX::kThreads;
  constexpr int kRegsProd = MX::kRegsProd;
  constexpr int E = G::kEntries;  // 16 entries per stage
  constexpr int kPS = WP::kPS;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  float* slots = reinterpret_cast<float*>(smem + S::off_slots);
  float* ng = reinterpret_cast<float*>(smem + S::off_ng);
  float* bpart = reinterpret_cast<float*>(smem + S::off_bpart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bars);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + kAccSlots;
  uint64_t* w_full = acc_empty + kAccSlots;
  uint64_t* w_empty = w_full + kWSlots;
  uint64_t* b_full = w_empty + kWSlots;
  uint64_t* b_empty = b_full + kBSlots;
  uint64_t* raw_full = b_empty + kBSlots;  // [kStages] gathered bytes of a stage have landed (kTma)
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + S::off_misc);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int k = p.k;
#ifdef ALS_PROFILE_WAITS
  const long long prof_t0 = clock64();
#endif
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // operand atoms need 1024-byte alignment

  if (tid == 0) {
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full[i], kArrive1 ? 1 : 32);  // the producer warp that owns the stage (every lane, or one elected)
      mbar_init(&empty[i], 1);
      // gathered bytes of the stage: TMA transaction bytes behind one arrival, or the copy-completion
      // arrivals of the 32 lanes that gathered it
      mbar_init(&raw_full[i], (kGather4 && MX::kAsync) ? 1 : 32);
    }
    for (int i = 0; i < kAccSlots; i++) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kArrive1 ? 4 : 128);
    }
    // W hand-off barriers are per SLOT (row u -> slot u % kWSlots): the drain sees every phase of
    // a slot's barriers in order, and phase n+1 of w_full cannot complete before the one waiter
    // of phase n (the Cholesky warp of row n * kWSlots + slot) has released the slot.
    for (int i = 0; i < kWSlots; i++) {
      mbar_init(&w_full[i], kArrive1 ? 4 : 128);
      mbar_init(&w_empty[i], 1);
    }
    // rhs hand-off barriers per rhs slot (row u -> slot u % kBSlots), same argument
    for (int i = 0; i < kBSlots; i++) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init_fence();
  }
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_base_s, kTmemCols);
  // -G in the panel layout of the slots (padding rows / columns and the unused upper triangles
  // of the diagonal blocks are 0)
  for (int e = tid; e < WP::kFloats; e += kThreads) ng[e] = 0.f;
  __syncthreads();
  for (int e = tid; e < KS * KS; e += kThreads) {
    const int i = e / KS, j = e % KS;
    if (i >= j && i < k) ng[WP::at(i, j)] = -(float)p.G[i * KS + j];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_s;
  const long long row_step = gridDim.x;

  if (warp >= kFirstProd) {
   if constexpr (MX::kSetReg && kRegsProd < MX::kLaunchRegs) umma::reg_dealloc<kRegsProd>();
   if (warp < kMmaWarp) {
    if constexpr (MX::kAsync) {
    // =========================== producers: async gather + in-place conversion ========
    // Each producer owns every P-th stage.  For a stage it (1) issues the asynchronous gather
    // of the stage kAhead own-stages ahead (16-byte cp.async copies straight into that stage's
    // ring slot, zero-fill beyond the row's last entry, completion on the slot's mbarrier),
    // (2) prefetches the indices / values of the stage after that, (3) waits for its current
    // stage's bytes, reads the raw fp32 rows from the slot and converts them in place.
    const int pw = warp - kFirstProd;
    constexpr int CPR = G::kChunksPerRow;  // lanes per factor row (16-byte chunks)
    constexpr int RPP = 32 / CPR;          // entries per pass of the warp
    static_assert(E / RPP == 8, "eight passes per stage");
    const int q = lane % CPR;    // 16-byte chunk of the factor row
    const int sub = lane / CPR;  // which of the RPP entries of a pass
    const int el0 = lane % E;    // entry whose index / value / scalars this lane holds
    constexpr int D = MX::kAhead;
    // index / value queue depth: the entries of a stage are fetched (HBM latency: ~1.5k cycles)
    // kFetchExtra + 1 own-stages before its gather is issued
    constexpr int QD = D + 1 + kFetchExtra;
    uint32_t oh0, ol0;
    G::slots(sub, q, oh0, ol0);  // pass 0; lo = hi ^ 32
    const uint32_t ring_a = smem_u32(ring);
    uint32_t f = (uint32_t)pw;              // my current stage
    uint32_t slot = (uint32_t)pw, par = 0;  // its ring slot and phase parity
    uint32_t base = 0;
    int useq_base = 0;
    float2 bacc01 = make_float2(0.f, 0.f), bacc23 = make_float2(0.f, 0.f);
    const float alpha = p.alpha;
    const bool recon = p.reconstruct_r != 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      RowBatch B;
      load_batch<E>(p, rb, row_step, lane, base, B);
      // index / value of entry el0 of flat stage ff (-1 / 0 beyond the stage or the batch)
      auto fetch = [&](uint32_t ff, int& idx_o, float& val_o) {
        idx_o = -1;
        val_o = 0.f;
        if (ff < B.batch_end) {  // warp-uniform
          const int i = __ffs(__ballot_sync(kFull, B.end > ff)) - 1;
          const int cnt = __shfl_sync(kFull, B.cnt, i);
          const uint32_t end = __shfl_sync(kFull, B.end, i);
          const long long e0 = shfl_i64(B.e0, i);
          const int st = (int)(ff - (end - (uint32_t)((cnt + E - 1) / E)));
          if (st * E + el0 < cnt) {
            idx_o = ld_stream_i32(p.col_idx + e0 + st * E + el0);
            val_o = ld_stream_f32(p.val + e0 + st * E + el0);
          }
        }
      };
      // gather of one stage into ring slot gs (phase parity gp): every lane copies chunk q of
      // entries sub, sub + RPP, ...; the slot's mbarrier gets this lane's arrival when they land
      auto issue = [&](int my_idx, uint32_t gs, uint32_t gp) {
        mbar_wait_id(&empty[gs], gp ^ 1u, 1);  // the MMA has read the slot's previous stage
        if constexpr (kGather4) {
          // lane e (< E) holds the index of entry e: every fourth lane gathers entries e .. e + 3 with one
          // TMA instruction (rows land at entry * KS * 4, exactly where the 16-byte copies put them)
          const int i0 = my_idx < 0 ? kOobRow : my_idx;
          const int i1 = __shfl_down_sync(kFull, i0, 1);
          const int i2 = __shfl_down_sync(kFull, i0, 2);
          const int i3 = __shfl_down_sync(kFull, i0, 3);
          const uint32_t bar = smem_u32(&raw_full[gs]);
          if (lane == 0) mbar_arrive_expect_tx(&raw_full[gs], (uint32_t)G::kBytes);
          if (lane < E && (lane & 3) == 0)
            tma_gather4(ring_a + gs * (uint32_t)G::kBytes + (uint32_t)(lane * (KS * 4)), &tmapM, i0, i1, i2, i3, bar);
          return;
        }
        const uint32_t dst = ring_a + gs * (uint32_t)G::kBytes + (uint32_t)(sub * (KS * 4) + q * 16);
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
          const int ci = __shfl_sync(kFull, my_idx, sub + RPP * ps);
          cp_async_16_zfill(dst + (uint32_t)(RPP * ps * (KS * 4)),
                            p.M + (long long)(ci < 0 ? 0 : ci) * KS + 4 * q, ci < 0 ? 0u : 16u);
        }
        cp_async_mbar_arrive_noinc(&raw_full[gs]);
      };
      // queue of my next D + 1 stages' (index, value): position d <-> stage f + d * P; shifted
      // down by one every step (a few register moves instead of D + 1 unrolled loop bodies)
      int q_idx[QD];
      float q_val[QD];
#pragma unroll
      for (int d = 0; d < QD; d++) fetch(f + (uint32_t)(d * P), q_idx[d], q_val[d]);
      uint32_t gslot = slot, gpar = par;  // ring position of the next stage to gather
#pragma unroll
      for (int d = 0; d < D; d++) {
        if (f + (uint32_t)(d * P) < B.batch_end) issue(q_idx[d], gslot, gpar);
        gslot += P;
        if (gslot >= (uint32_t)kStages) { gslot -= kStages; gpar ^= 1u; }
      }
#pragma unroll 1
      while (f < B.batch_end) {
        {
          const float my_val = q_val[0];
          // (1) gather of the stage D own-stages ahead
          if (f + (uint32_t)(D * P) < B.batch_end) issue(q_idx[D], gslot, gpar);
          gslot += P;
          if (gslot >= (uint32_t)kStages) { gslot -= kStages; gpar ^= 1u; }
          // (2) shift the queue; indices / values of the stage after that into its last position
#pragma unroll
          for (int d = 0; d < QD - 1; d++) { q_idx[d] = q_idx[d + 1]; q_val[d] = q_val[d + 1]; }
          fetch(f + (uint32_t)(QD * P), q_idx[QD - 1], q_val[QD - 1]);
          // (3) my current stage
          const int i = __ffs(__ballot_sync(kFull, B.end > f)) - 1;
          const int cnt = __shfl_sync(kFull, B.cnt, i);
          const uint32_t end = __shfl_sync(kFull, B.end, i);
          const int nst = (cnt + E - 1) / E;
          const int st = (int)(f - (end - (uint32_t)nst));
          // per-entry scalars, once per entry (lane q of either half):
          // SYRK weight (c_u - 1) = alpha*|r| (ALS.java:471-479), 0 when reconstructing R (:466-469);
          // rhs weight r (:466-469) or c_u gated on r > 0 (:480-482)
          const float ar = alpha * fabsf(my_val);
          const float my_s = recon ? 0.f : sqrt_approx(ar);
          const float my_cb = recon ? my_val : (my_val > 0.f ? 1.f + ar : 0.f);
          mbar_wait_id(&raw_full[slot], par, 11);
          const unsigned char* raw = ring + (size_t)slot * G::kBytes;
          float4 y[8];
#pragma unroll
          for (int ps = 0; ps < 8; ps++)  // (zero-filled beyond the stage's entries)
            y[ps] = *reinterpret_cast<const float4*>(raw + (sub + RPP * ps) * (KS * 4) + q * 16);
          __syncwarp();  // every lane holds its part of the raw rows: the tile may be overwritten
          const uint32_t st_hi = ring_a + slot * (uint32_t)G::kBytes + oh0;  // shared-window address
#pragma unroll
          for (int ps = 0; ps < 8; ps++) {
            const float sc = __shfl_sync(kFull, my_s, sub + RPP * ps);
            const float cb = __shfl_sync(kFull, my_cb, sub + RPP * ps);
            // packed fp32x2 arithmetic: scale, split, rhs (half the issue slots of the scalar form)
            const float2 y01 = make_float2(y[ps].x, y[ps].y), y23 = make_float2(y[ps].z, y[ps].w);
            const float2 sc2 = make_float2(sc, sc), cb2 = make_float2(cb, cb);
            const float2 v01 = fmul2(y01, sc2), v23 = fmul2(y23, sc2);
            // bf16 hi + bf16 lo, round-to-nearest both times
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(v01.x, v01.y);
            const __nv_bfloat162 h23 = __floats2bfloat162_rn(v23.x, v23.y);
            const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01);
            const uint32_t u23 = *reinterpret_cast<const uint32_t*>(&h23);
            const float2 r01 = fsub2(v01, make_float2(__uint_as_float(u01 << 16), __uint_as_float(u01 & 0xffff0000u)));
            const float2 r23 = fsub2(v23, make_float2(__uint_as_float(u23 << 16), __uint_as_float(u23 & 0xffff0000u)));
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(r01.x, r01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(r23.x, r23.y);
            // relative to pass 0 the slot address differs by an XOR (swizzle + K-row: below the
            // ring's 1 KB alignment, so XOR on the address is exact) and an offset (K-atom / K-step)
            const uint32_t xo = G::pass_xor(ps);
            const uint32_t ko = G::pass_add(ps);
            sts_v2((st_hi ^ xo) + ko, u01, u23);
            sts_v2((st_hi ^ (xo ^ 32u)) + ko, *reinterpret_cast<const uint32_t*>(&l01),
                   *reinterpret_cast<const uint32_t*>(&l23));
            bacc01 = ffma2(cb2, y01, bacc01);
            bacc23 = ffma2(cb2, y23, bacc23);
          }
          // my last stage of this row: publish the partial rhs of my stages (before the stage's
          // `full` arrive: the MMA warp's "rhs complete" signal then covers it)
          if (st + P >= nst) {
            const int useq = useq_base + __popc(B.ne_mask & ((1u << i) - 1u));
            const int bslot = useq % kBSlots;
            float4 v = make_float4(bacc01.x, bacc01.y, bacc23.x, bacc23.y);
#pragma unroll
            for (int off = CPR; off < 32; off <<= 1) {
              v.x += __shfl_xor_sync(kFull, v.x, off);
              v.y += __shfl_xor_sync(kFull, v.y, off);
              v.z += __shfl_xor_sync(kFull, v.z, off);
              v.w += __shfl_xor_sync(kFull, v.w, off);
            }
            mbar_wait_id(&b_empty[bslot], (uint32_t)(((useq / kBSlots) & 1) ^ 1), 0);
            if (lane < CPR) *reinterpret_cast<float4*>(bpart + (bslot * P + pw) * KS + 4 * q) = v;
            bacc01 = make_float2(0.f, 0.f);
            bacc23 = make_float2(0.f, 0.f);
          }
          fence_proxy_async_smem();
          if (kArrive1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
          } else {
            mbar_arrive(&full[slot]);
          }
          f += P;
          slot += P;
          if (slot >= (uint32_t)kStages) { slot -= kStages; par ^= 1u; }
        }
      }
      base = B.batch_end;
      useq_base += __popc(B.ne_mask);
    }
    } else {
    // =========================== producers ===========================================
    const int pw = warp - kFirstProd;
    const int q = lane & 15;    // 16-byte chunk of the factor row / entry slot this lane prepares
    const int sub = lane >> 4;  // which of the two entries of a pass
    uint32_t oh0, ol0;
    G::slots(sub, q, oh0, ol0);  // pass 0; lo = hi ^ 32
    const uint32_t ring_a = smem_u32(ring);
    uint32_t f = (uint32_t)pw;       // my next flat stage
    uint32_t slot = (uint32_t)pw, par = 0;  // its ring slot and phase parity
    uint32_t base = 0;               // flat index of the first stage of the batch
    int useq_base = 0;               // non-empty rows before the batch
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float alpha = p.alpha;
    const bool recon = p.reconstruct_r != 0;

    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      RowBatch B;
      load_batch<E>(p, rb, row_step, lane, base, B);
      // (row, stage) of flat stage ff inside this batch
      auto locate = [&](uint32_t ff, int& i, int& st, int& cnt, long long& e0) {
        i = __ffs(__ballot_sync(kFull, B.end > ff)) - 1;
        cnt = __shfl_sync(kFull, B.cnt, i);
        const uint32_t end = __shfl_sync(kFull, B.end, i);
        st = (int)(ff - (end - (uint32_t)((cnt + E - 1) / E)));
        e0 = shfl_i64(B.e0, i);
      };
      int n_idx = -1;     // prefetched index / value of entry q of my next stage
      float n_val = 0.f;
      bool pref = false;
      while (f < B.batch_end) {
        int i, st, cnt;
        long long e0;
        locate(f, i, st, cnt, e0);
        const long long es = e0 + (long long)st * E;
        const int n_here = cnt - st * E;  // >= 1; entries of this stage = min(E, n_here)
        int my_idx = n_idx;
        float my_val = n_val;
        if (!pref) {
          const bool ok = q < n_here;
          my_idx = ok ? ld_stream_i32(p.col_idx + es + q) : -1;
          my_val = ok ? ld_stream_f32(p.val + es + q) : 0.f;
        }
        // all gathers of the stage up front
        float4 y[8];
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
          const int ci = __shfl_sync(kFull, my_idx, sub + 2 * ps);
          y[ps] = (ci >= 0) ? ldg_f4(p.M + (long long)ci * KS + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // indices / values of my next stage while the rows are in flight
        {
          const uint32_t f2 = f + P;
          pref = f2 < B.batch_end;
          if (pref) {
            int i2, st2, cnt2;
            long long e02;
            locate(f2, i2, st2, cnt2, e02);
            const bool ok = q < cnt2 - st2 * E;
            const long long es2 = e02 + (long long)st2 * E;
            n_idx = ok ? ld_stream_i32(p.col_idx + es2 + q) : -1;
            n_val = ok ? ld_stream_f32(p.val + es2 + q) : 0.f;
          }
        }
        // per-entry scalars, once per entry (lane q of either half):
        // SYRK weight (c_u - 1) = alpha*|r| (ALS.java:471-479), 0 when reconstructing R (:466-469);
        // rhs weight r (:466-469) or c_u gated on r > 0 (:480-482)
        const float ar = alpha * fabsf(my_val);
        const float my_s = recon ? 0.f : sqrt_approx(ar);
        const float my_cb = recon ? my_val : (my_val > 0.f ? 1.f + ar : 0.f);

        mbar_wait_id(&empty[slot], par ^ 1u, 1);
        const uint32_t st_hi = ring_a + slot * (uint32_t)G::kBytes + oh0;  // shared-window address
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
          const float s = __shfl_sync(kFull, my_s, sub + 2 * ps);
          const float cb = __shfl_sync(kFull, my_cb, sub + 2 * ps);
          const float4 v = make_float4(y[ps].x * s, y[ps].y * s, y[ps].z * s, y[ps].w * s);
          // bf16 hi + bf16 lo, round-to-nearest both times
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
          const __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
          const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01);
          const uint32_t u23 = *reinterpret_cast<const uint32_t*>(&h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __uint_as_float(u01 << 16),
                                                           v.y - __uint_as_float(u01 & 0xffff0000u));
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __uint_as_float(u23 << 16),
                                                           v.w - __uint_as_float(u23 & 0xffff0000u));
          // pass ps fills K-row 2(ps&3)+sub of K-atom ps>>2: relative to pass 0 the slot address
          // differs by ^((ps&3) << 5) (the swizzle), ^((ps&3) << 8) (the K-row; both below the
          // ring's 1 KB alignment, so XOR on the address is exact) and + (ps>>2) * 2048
          const uint32_t xo = (uint32_t)(((ps & 3) << 5) | ((ps & 3) << 8));
          const uint32_t ko = (uint32_t)((ps >> 2) << 11);
          sts_v2((st_hi ^ xo) + ko, u01, u23);
          sts_v2((st_hi ^ (xo ^ 32u)) + ko, *reinterpret_cast<const uint32_t*>(&l01),
                 *reinterpret_cast<const uint32_t*>(&l23));
          bacc.x = fmaf(cb, y[ps].x, bacc.x);
          bacc.y = fmaf(cb, y[ps].y, bacc.y);
          bacc.z = fmaf(cb, y[ps].z, bacc.z);
          bacc.w = fmaf(cb, y[ps].w, bacc.w);
        }
        // my last stage of this row: publish the partial rhs of my stages (before the stage's
        // `full` arrive: the MMA warp's "rhs complete" signal then covers it)
        if (st + P >= (cnt + E - 1) / E) {
          const int useq = useq_base + __popc(B.ne_mask & ((1u << i) - 1u));
          const int bslot = useq % kBSlots;
          float4 v = bacc;
          v.x += __shfl_xor_sync(kFull, v.x, 16);
          v.y += __shfl_xor_sync(kFull, v.y, 16);
          v.z += __shfl_xor_sync(kFull, v.z, 16);
          v.w += __shfl_xor_sync(kFull, v.w, 16);
          mbar_wait_id(&b_empty[bslot], (uint32_t)(((useq / kBSlots) & 1) ^ 1), 0);
          if (lane < 16) *reinterpret_cast<float4*>(bpart + (bslot * P + pw) * KS + 4 * q) = v;
          bacc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
        if (kArrive1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[slot]);
        } else {
          mbar_arrive(&full[slot]);
        }
        f += P;
        slot += P;
        if (slot >= (uint32_t)kStages) { slot -= kStages; par ^= 1u; }
      }
      base = B.batch_end;
      useq_base += __popc(B.ne_mask);
    }
    }
   } else {
    // =========================== MMA issuer ==========================================
    // The whole warp runs this loop converged; only the tcgen05 instructions are issued by one
    // elected lane.  One trip per 16 entries: the serial resource of the CTA.
    const uint32_t idesc = umma::make_idesc_bf16_mn(G::kM, G::kN);
    const uint64_t desc0 = umma::make_smem_desc(smem_u32(ring), G::kLBO, G::kSBO);
    const uint32_t dhi = (uint32_t)(desc0 >> 32);
    uint32_t slot = 0, par = 0;
    uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), dlo = (uint32_t)desc0;
    uint32_t gseg = 0;
    uint32_t useq = 0;
    long long row = blockIdx.x;  // position in the visiting order
    auto count_at = [&](long long pos) {
      const long long r = p.row_order ? (long long)__ldg(p.row_order + pos) : pos;
      return (int)(__ldg(p.row_ptr + r + 1) - __ldg(p.row_ptr + r));
    };
    int cnt_next = 0;
    if (row < p.n_rows) cnt_next = count_at(row);
    for (; row < p.n_rows; row += row_step) {
      const int cnt = cnt_next;
      const long long nrow = row + row_step;
      if (nrow < p.n_rows) cnt_next = count_at(nrow);
      if (cnt == 0) continue;
      const int nst = (cnt + E - 1) / E;
      for (int st0 = 0; st0 < nst; st0 += kSegStages, gseg++) {
        const int n = (nst - st0 < kSegStages) ? nst - st0 : kSegStages;
        const uint32_t a = gseg % kAccSlots;
        mbar_wait_id(&acc_empty[a], ((gseg / kAccSlots) & 1) ^ 1, 2);
        const uint32_t d_tmem = tmem_base + a * (uint32_t)G::kN;
        for (int t = 0; t < n; t++) {
#ifdef ALS_PROFILE_WAITS
          {
            const long long t0 = clock64();
            mbar_wait_addr(full_a, par);
            if (lane == 0) atomicAdd(&umma::g_wait_cycles[3], (unsigned long long)(clock64() - t0));
          }
#else
          umma::mbar_wait_addr_paced<MX::kPaceMma>(full_a, par);
#endif
          tc_fence_after_sync();
#pragma unroll
          for (int ks = 0; ks < G::kKSteps; ks++)
            umma::mma_bf16_ss_same_elect(d_tmem, dlo + (uint32_t)(ks * (int)(G::kKStepBytes >> 4)), dhi, idesc,
                                         (t > 0 || ks > 0) ? 1u : 0u);
          umma::mma_commit_addr_elect(empty_a);  // frees the operand stage once the MMA has read it
          slot++; full_a += 8; empty_a += 8; dlo += (uint32_t)(G::kBytes >> 4);
          if (slot == kStages) {
            slot = 0; par ^= 1u;
            full_a -= 8 * kStages; empty_a -= 8 * kStages; dlo -= (uint32_t)(kStages * (G::kBytes >> 4));
          }
        }
        umma::mma_commit_elect(&acc_full[a]);  // accumulator segment complete
      }
      // every stage of the row has been seen full: the producers' rhs partials are in place
      if (lane == 0) mbar_arrive(&b_full[useq % kBSlots]);
      __syncwarp();
      useq++;
    }
   }
  } else if (warp < kDrainWarps) {
    // =========================== drain warpgroup =====================================
    // Warp qd owns TMEM lane quarter qd = matrix rows 16qd..16qd+15 (lanes +0..15: hi operand
    // rows, +16..31: lo rows); D's column group fc (32 columns) = [hi | lo] of features
    // 16fc..16fc+15.  Block (qd, fc), fc <= qd, of W = the sum of the four hi/lo quadrants.
    if constexpr (MX::kSetReg) umma::reg_dealloc<kRegsDrain>();
    const int qd = warp;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t lane_hi = (uint32_t)(32 * qd) << 16, lane_lo = (uint32_t)(32 * qd + 16) << 16;
    const bool has_diag = (t == (g >> 1));
    // accumulator-fragment addressing of the swizzled panel rows (see chol_blocked.cuh)
    const int cfr_base = g * kPS + (t & 1) * 2, cfr_chunk = (t >> 1) ^ WP::swz(g);
    uint32_t gseg = 0;
    int useq = 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      int cnt_l = 0;
      {
        const long long mypos = rb + lane * row_step;
        if (mypos < p.n_rows) {
          const long long myrow = p.row_order ? (long long)p.row_order[mypos] : mypos;
          cnt_l = (int)(p.row_ptr[myrow + 1] - p.row_ptr[myrow]);
          if constexpr (kExt) {
            if (p.vacc && p.vacc[myrow] >= 0) cnt_l = -cnt_l;  // a chunk of a split row (never empty)
          }
        }
      }
      const long long left = (p.n_rows - rb + row_step - 1) / row_step;
      const int nb = left < 32 ? (int)left : 32;
      for (int ib = 0; ib < nb; ib++) {
        const int cnt_s = __shfl_sync(kFull, cnt_l, ib);
        if (cnt_s == 0) continue;
        // chunks hand over their partial -D only: G and lambda alpha n_u are added once, by the warp
        // that assembles the row
        const bool is_chunk = kExt && cnt_s < 0;
        const bool no_g = kExt && (is_chunk || p.stash != nullptr);  // (stash mode: the solving warp adds them, see below)
        const int cnt = is_chunk ? -cnt_s : cnt_s;
        const int nst = (cnt + E - 1) / E;
        const int nseg = (nst + kSegStages - 1) / kSegStages;
        const int ws = useq % kWSlots;
        float* slot = slots + ws * WP::kFloats;
        // W = G + lambda*alpha*n_u*I + ... (ALS.java:447-450, 488-492); padding rows (>= k) get a
        // unit diagonal so the factorisation stays finite.  Rows g and g+8 of the block:
        const float lam_n = (float)(p.lambda_alpha * (double)cnt);
        const float lam0 = (16 * qd + g < k) ? lam_n : 1.f;
        const float lam1 = (16 * qd + g + 8 < k) ? lam_n : 1.f;
        mbar_wait_id(&w_empty[ws], (uint32_t)(((useq / kWSlots) & 1) ^ 1), 4);
        for (int seg = 0; seg < nseg; seg++, gseg++) {
          const int a = (int)(gseg % kAccSlots);
#if defined(ALS_PROFILE_WAITS) || defined(ALS_WATCHDOG)
          mbar_wait_id(&acc_full[a], (gseg / kAccSlots) & 1, 5);
#else
          umma::mbar_wait_paced<MX::kPaceDrain>(&acc_full[a], (gseg / kAccSlots) & 1);
#endif
          tc_fence_after_sync();
          const uint32_t tcol = tmem_base + (uint32_t)(a * G::kN);
          const float* src = (seg == 0) ? ng : slot;  // later segments add to what this thread stored
#pragma unroll
          for (int fc = 0; fc < WP::kNP; fc++) {
            if (fc > qd || qd >= WP::kNP) break;  // warp-uniform: block right of the diagonal / no such block row
            uint32_t r[16], s[16];
            tmem_ld_16x256b_x4(tcol + lane_hi + 32 * fc, r);
            tmem_ld_16x256b_x4(tcol + lane_lo + 32 * fc, s);
            umma::tmem_wait_ld();
            float wv[8];
#pragma unroll
            for (int e = 0; e < 8; e++)
              wv[e] = (__uint_as_float(r[e]) + __uint_as_float(r[8 + e])) +
                      (__uint_as_float(s[e]) + __uint_as_float(s[8 + e]));
            // wv[4i+0..1]: row g, columns 8i+2t, +1; wv[4i+2..3]: row g+8
            const int boff = WP::panel_off(fc) + 16 * (qd - fc) * kPS + cfr_base;
            float2 o[4];
            const bool from_zero = no_g && seg == 0;  // warp-uniform
#pragma unroll
            for (int i = 0; i < 2; i++) {
              float2 n0 = make_float2(0.f, 0.f), n1 = make_float2(0.f, 0.f);
              if (!from_zero) {
                n0 = *reinterpret_cast<const float2*>(src + boff + ((cfr_chunk ^ (2 * i)) << 2));
                n1 = *reinterpret_cast<const float2*>(src + boff + 8 * kPS + ((cfr_chunk ^ (2 * i)) << 2));
              }
              o[2 * i] = make_float2(n0.x - wv[4 * i], n0.y - wv[4 * i + 1]);
              o[2 * i + 1] = make_float2(n1.x - wv[4 * i + 2], n1.y - wv[4 * i + 3]);
            }
            if (fc == qd && seg == 0 && has_diag && !no_g) {
              // diagonal entries: row g at column g (i = 0), row g+8 at column g+8 (i = 1)
              if (g & 1) { o[0].y -= lam0; o[3].y -= lam1; }
              else { o[0].x -= lam0; o[3].x -= lam1; }
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
              *reinterpret_cast<float2*>(slot + boff + ((cfr_chunk ^ (2 * i)) << 2)) = o[2 * i];
              *reinterpret_cast<float2*>(slot + boff + 8 * kPS + ((cfr_chunk ^ (2 * i)) << 2)) = o[2 * i + 1];
            }
          }
          tc_fence_before_sync();
          if (kArrive1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
          } else {
            mbar_arrive(&acc_empty[a]);  // accumulator may be overwritten by the next segment
          }
        }
        if (kArrive1) {
          __syncwarp();  // (orders every lane's slot stores before the elected lane's release)
          if (lane == 0) mbar_arrive(&w_full[ws]);
        } else {
          mbar_arrive(&w_full[ws]);  // 128 arrivals: the slot of this row is complete
        }
        useq++;
      }
    }
  } else {
    // =========================== Cholesky warps ======================================
    if constexpr (MX::kSetReg && MX::kRegsCholMix > MX::kLaunchRegs) umma::reg_alloc<MX::kRegsCholMix>();
    const int cw = warp - kFirstChol;
    constexpr int kGroups = (ALS_V2_LOCKSTEP > kCholWarps / 2) ? kCholWarps / 2 : ALS_V2_LOCKSTEP;
    constexpr int kGroupWarps = kGroups ? kCholWarps / kGroups : 1;
    float* scratch = reinterpret_cast<float*>(smem + S::off_scratch) + cw * CB::kScratch;
    int useq = 0;
    uint32_t base = 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      RowBatch B;
      load_batch<E>(p, rb, row_step, lane, base, B);
      // flat index of my row's first stage modulo the number of producers: producer
      // (start + s) % P owns stage s of the row
      const int nst_l = (B.cnt + E - 1) / E;
      const int smod_l = (int)((B.end - (uint32_t)nst_l) % (uint32_t)P);
      for (int ib = 0; ib < B.nb; ib++) {
        const int cnt = __shfl_sync(kFull, B.cnt, ib);
        if (cnt == 0) continue;
        if (useq % kCholWarps != cw) { useq++; continue; }
        const long long pos = rb + ib * row_step;
        const long long vr = p.row_order ? (long long)__ldg(p.row_order + pos) : pos;  // (virtual) row walked
        long long row = vr;  // the row it belongs to
        int acc = -1;
        if constexpr (kExt) {
          if (p.vrow) row = (long long)__ldg(p.vrow + vr);
          if (p.vacc) acc = __ldg(p.vacc + vr);
        }
        const int nst = (cnt + E - 1) / E;
        const int smod = __shfl_sync(kFull, smod_l, ib);
        const int bs = useq % kBSlots;
#if defined(ALS_PROFILE_WAITS) || defined(ALS_WATCHDOG)
        mbar_wait_id(&b_full[bs], (uint32_t)((useq / kBSlots) & 1), 7);
#else
        umma::mbar_wait_paced<MX::kPaceChol>(&b_full[bs], (uint32_t)((useq / kBSlots) & 1));
#endif
        float b[CB::kS];
#pragma unroll
        for (int s = 0; s < CB::kS; s++) b[s] = 0.f;
#pragma unroll
        for (int w = 0; w < P; w++) {
          int rel = w - smod;
          if (rel < 0) rel += P;
          if (rel < nst) {  // producer w had a stage in this row
            const float* bp = bpart + (bs * P + w) * KS;
#pragma unroll
            for (int s = 0; s < CB::kS; s++) b[s] += bp[lane + 32 * s];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_empty[bs]);
        const int ws = useq % kWSlots;
        float* slot = slots + ws * WP::kFloats;
#if defined(ALS_PROFILE_WAITS) || defined(ALS_WATCHDOG)
        mbar_wait_id(&w_full[ws], (uint32_t)((useq / kWSlots) & 1), 6);
#else
        umma::mbar_wait_paced<MX::kPaceChol>(&w_full[ws], (uint32_t)((useq / kWSlots) & 1));
#endif
#ifdef ALS_PROFILE_WAITS
        const long long tb = clock64();
        if (kGroups) bar_sync(1 + cw / kGroupWarps, kGroupWarps * 32);
        const long long ts = clock64();
        if (lane == 0) atomicAdd(&umma::g_wait_cycles[9], (unsigned long long)(ts - tb));
#else
        if (kGroups) bar_sync(1 + cw / kGroupWarps, kGroupWarps * 32);
#endif
#ifdef ALS_DEBUG_SLOT
        // development builds only (scripts/v2_debug.py): copy the slot (N = -W_u) and the rhs of
        // one row out before the solve touches them
        if (p.row_offset + row == g_debug_row) {
          for (int e = lane; e < WP::kFloats; e += 32) g_debug_slot[e] = slot[e];
#pragma unroll
          for (int s = 0; s < CB::kS; s++) g_debug_slot[WP::kFloats + lane + 32 * s] = b[s];
          __syncwarp();
        }
#endif
        if (lane == 0) bulk_wait_group_read0();  // the previous row's bulk stores have read the scratch vector
        __syncwarp();
        float* st = nullptr;
        if constexpr (kExt) {
        if (p.stash) st = p.stash + ((size_t)blockIdx.x * kCholWarps + cw) * (size_t)(WP::kFloats + KS);
        const float* ga = nullptr;
        if (acc >= 0) {
          // a chunk of a split row: add my partial -D and rhs to the row's record; whoever arrives last
          // assembles N = -(G + lambda alpha n_u I) + sum of the partials and solves
          float* gw = p.gacc + (size_t)acc * (size_t)(WP::kFloats + KS);
          for (int e = lane; e < WP::kFloats; e += 32) atomicAdd(gw + e, slot[e]);
#pragma unroll
          for (int s = 0; s < CB::kS; s++) atomicAdd(gw + WP::kFloats + lane + 32 * s, b[s]);
          __threadfence();
          __syncwarp();
          int arrived = 0;
          if (lane == 0) arrived = atomicAdd(p.gcount + acc, 1);
          arrived = __shfl_sync(kFull, arrived, 0);
          if (arrived != __ldg(p.acc_chunks + acc) - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_empty[ws]);
            useq++;
            continue;
          }
          __threadfence();
          ga = gw;
#pragma unroll
          for (int s = 0; s < CB::kS; s++) b[s] = __ldcg(ga + WP::kFloats + lane + 32 * s);
        }
        if (ga || st) {
          // the slot holds -D only (chunks, stash mode) or nothing yet (assembled row): N = -G - D - lambda alpha n_u I,
          // keeping a copy of -D and the rhs in stash mode
          const float lam_n = (float)(p.lambda_alpha * (double)(p.real_ptr[row + 1] - p.real_ptr[row]));
          for (int e = lane; e < WP::kFloats; e += 32) {
            const float d = ga ? __ldcg(ga + e) : slot[e];
            if (st) st[e] = d;
            slot[e] = ng[e] + d;
          }
          if (st) {
#pragma unroll
            for (int s = 0; s < CB::kS; s++) st[WP::kFloats + lane + 32 * s] = b[s];
          }
          __syncwarp();
          for (int i = lane; i < KS; i += 32) slot[WP::at(i, i)] -= (i < k) ? lam_n : 1.f;
          __syncwarp();
        }
        }  // kExt
        const float dmax = CB::diag_max(slot, lane, k);
        const bool ok = CB::factor_solve(slot, scratch, b, dmax, p.threshold, kCondLimit, lane, k);
#ifdef ALS_PROFILE_WAITS
        if (lane == 0) atomicAdd(&umma::g_wait_cycles[10], (unsigned long long)(clock64() - ts));
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&w_empty[ws]);
        if (ok) {
          // The solution sits in the warp's scratch vector in row order (padding entries 0): one
          // TMA bulk copy (cp.async.bulk shared -> global, 256 B, warp-uniform addresses) per
          // destination -- the local factor row and, sharded, the same row of every peer replica
          // over NVLink -- instead of per-lane stores.
          float* xs = scratch + 32;
          const long long eoff = (p.row_offset + row) * KS;
          fence_proxy_async_smem();  // the solver's generic-proxy writes of xs -> visible to the bulk copy
          __syncwarp();
          if (lane == 0) {
            bulk_s2g(p.out + eoff, smem_u32(xs), KS * 4);
#pragma unroll 1
            for (int r = 0; r < p.n_peers; r++) bulk_s2g(p.peer_out[r] + eoff, smem_u32(xs), KS * 4);
            bulk_commit_group();  // (the scratch vector is reused only after wait_group.read, below)
          }
        } else {
          // refused by the fp32 path.  Stash mode: my copy of -D and the rhs goes to the fp64 re-solve;
          // otherwise (or when its buffer is full) the row is gathered again by the fp64 kernel.
          bool handed = false;
          if (kExt && st) {
            int rs = 0;
            if (lane == 0) rs = atomicAdd(p.resolve_count, 1);
            rs = __shfl_sync(kFull, rs, 0);
            if (rs < p.resolve_cap) {
              float* rec = p.resolve_buf + (size_t)rs * (size_t)(WP::kFloats + KS);
              for (int e = lane; e < WP::kFloats + KS; e += 32) rec[e] = st[e];  // (each lane re-reads what it wrote)
              if (lane == 0) p.resolve_rows[rs] = (int)row;
              handed = true;
            }
          }
          if (!handed && lane == 0) {
            const int rs = atomicAdd(p.retry_count, 1);
            p.retry_rows[rs] = (int)row;
          }
        }
        useq++;
      }
      base = B.batch_end;
    }
    if (lane == 0) bulk_wait_group0();  // every bulk store of this warp has been written
    // tail: warps without a row in the last round still meet their group at the barrier
    if (kGroups && (useq % kCholWarps) != 0 && cw >= (useq % kCholWarps))
      bar_sync(1 + cw / kGroupWarps, kGroupWarps * 32);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, kTmemCols);
#ifdef ALS_PROFILE_WAITS
  if (tid == 0) atomicAdd(&umma::g_wait_cycles[8], (unsigned long long)(clock64() - prof_t0));
#endif
}

}  // namespace v2

template <int KS, class MX, bool kExt = false>
inline int launch_row_update_v2_t(const RowUpdateParams& p, const CUtensorMap& tmapM, int sm_count, cudaStream_t stream,
                                  char* err, size_t err_len) {
  using S = v2::Smem<KS, MX>;
  long long grid = sm_count;
  if (grid > p.n_rows) grid = p.n_rows > 0 ? p.n_rows : 1;
  v2::row_update_v2_kernel<KS, MX, kExt><<<(int)grid, MX::kThreads, S::kTotal, stream>>>(p, tmapM);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, err_len, "row_update_v2 launch: %s", cudaGetErrorString(e));
    return ALS_E_CUDA;
  }
  return ALS_OK;
}

}  // namespace als
