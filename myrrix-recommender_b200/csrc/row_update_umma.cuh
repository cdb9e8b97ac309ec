// row_update_umma.cuh -- tensor-core (tcgen05 + TMEM) row update for padded feature
// counts 32 and 64.
//
// Same contract as row_update_simt.cuh (Worker.call, AlternatingLeastSquares.java:438-502),
// different machine mapping.  One persistent CTA (640 threads, 20 warps) per SM, rows strided
// over CTAs, four warp-specialised roles connected by mbarrier rings:
//
//   producers (7 or 11)  each owns every 7th/11th 16-entry stage (32 entries at k=32) of the
//                        CTA's flat stage stream: gather the factor rows from HBM (coalesced
//                        16-byte loads, all of a stage's gathers issued up front), scale by
//                        sqrt(alpha*|r|), split into bf16 hi + bf16 lo (x ~= hi+lo to 2^-17,
//                        round-to-nearest twice), store both halves into the swizzled MN-major
//                        operand stage; accumulate the rhs b_u in fp32 on the side.
//   MMA issuer (1 warp)  per 16 entries one tcgen05.mma with A = B = the stage's operand tile:
//                        D[2k x 2k] += [hi;lo][hi;lo]^T, fp32 accumulate in TMEM (all four
//                        cross products of the rank-16 update).  The warp runs converged with
//                        running barrier addresses / descriptor words; one elected lane issues.
//                        This loop is the CTA's serial resource (one trip per stage).
//   drain (1 warpgroup)  the operand rows are ordered so that TMEM lane quarter q holds the hi
//                        and lo halves of matrix rows 16q..16q+15: each drain warp folds columns
//                        and partner lanes for its own 16 rows, adds G (fp32 copy in smem) and
//                        lambda*alpha*n_u and writes the pair-packed lower triangle of W_u into a
//                        W slot in shared memory.  The four warps never wait for each other.
//   Cholesky (8 or 4)    each takes one row's W slot: in-register fp32 LDL^T + solves
//                        (chol_warp.cuh), writes the fp32 factor row.  Rows whose pivots look
//                        singular / ill-conditioned are appended to a retry list and re-solved
//                        in fp64 by the CUDA-core kernel (which owns the error reporting).
//                        All Cholesky warps of a CTA enter each sweep together (one instruction
//                        stream for the 80 KB of straight-line code) and hand their input slots
//                        back a quarter of the way into the sweep (see kReleaseDiv).
//
// The split of the 15 non-drain, non-MMA warps between producers and Cholesky warps is a
// template parameter chosen per launch from the average row length (Mix<NCHOL> below).
//
// Every role walks the CTA's rows in the same order; row lengths are fetched 32 rows at a
// time (one load per lane, shared by shuffles) so no role waits on a dependent row_ptr load
// per row (the MMA warp instead loads them one row ahead from warp-uniform addresses).
//
// Long rows are cut into segments of kSegStages stages so no fp32 TMEM accumulator carries
// more than 64 MMA steps before it is folded into the planes.
#pragma once
#include "chol_warp.cuh"
#include "common.cuh"
#include "row_update_simt.cuh"  // RowUpdateParams
#include "umma_common.cuh"

namespace als {
namespace umma {

constexpr int kDrainWarps = 4;                               // warps 0..3   (TMEM lane quarters)
constexpr int kFirstChol = kDrainWarps;
constexpr int kMmaWarp = 19;                                 // last warp (shares the producers' warpgroup)
constexpr int kThreads = (kMmaWarp + 1) * 32;                // 640: 5 warps per SM sub-partition
constexpr int kAccSlots = 4;                                 // TMEM accumulators in flight
constexpr int kSegStages = 64;                               // stages per accumulation segment
constexpr int kTmemCols = 512;
constexpr int kRegsDrain = 64, kRegsChol = 128;

// Role mix of the 15 warps between the drain warpgroup and the MMA issuer: NCHOL Cholesky warps
// (warps 4 .. 3+NCHOL), the rest producers.  Short rows (the X<-Y half of the headline
// workload: 100 entries, one k x k solve per 7 stages) are bound by the solves: 8 Cholesky + 7
// producer warps.  Long rows (the Y<-X half: 1000 entries per solve) are bound by the gather /
// operand staging: 4 Cholesky + 11 producer warps, a deeper operand ring instead of W slots.
template <int NCHOL>
struct Mix {
  static_assert(NCHOL == 8 || NCHOL == 4, "whole warpgroups per role");
  static constexpr int kCholWarps = NCHOL;
  static constexpr int kProdWarps = 15 - NCHOL;
  static constexpr int kFirstProd = kFirstChol + NCHOL;
  static constexpr int kStages = (NCHOL == 8) ? 16 : 24;     // operand ring depth (4 KB each)
  static constexpr int kWSlots = NCHOL;                      // W slots (drain -> Cholesky), one per Cholesky warp
  static constexpr int kBSlots = NCHOL;                      // rhs ring depth (== kCholWarps)
  // 640 threads x 96 registers at launch; the register file is per SM sub-partition (16384
  // registers, warp w lives on sub-partition w % 4): each sub-partition hosts 1 drain warp,
  // NCHOL/4 Cholesky warps and (16-NCHOL)/4 producer/MMA warps, and setmaxnreg re-balances
  // within 5 x 96 x 32 = 15360 registers.
  static constexpr int kRegsProd = (NCHOL == 8) ? 80 : 96;
  static_assert(32 * (kRegsDrain + (NCHOL / 4) * kRegsChol + ((16 - NCHOL) / 4) * kRegsProd) <= 5 * 96 * 32,
                "register pool");
};
// Forward-sweep column (as a fraction of the sweep: column KS / kReleaseDiv) at which a
// Cholesky warp hands its W and rhs slots back to the drain / producers; 0 = right after the
// load.  Measured on the headline X<-Y half (A/B on one box): 1/8 174 ms, 3/16 173, 1/4 168,
// 5/16 175, 7/16 190, at once 200.
#ifndef ALS_RELEASE_DIV
#define ALS_RELEASE_DIV 4
#endif
constexpr int kReleaseDiv = ALS_RELEASE_DIV;
#ifndef ALS_CHOL_LOCKSTEP
#define ALS_CHOL_LOCKSTEP 1
#endif
constexpr bool kCholLockstep = ALS_CHOL_LOCKSTEP != 0;   // all Cholesky warps enter each sweep together
constexpr float kCondLimit = 256.f;  // max diag / min pivot above which a row goes to fp64
constexpr unsigned kFull = 0xffffffffu;

template <int KS, int NCHOL>
struct Smem {
  using CW = CholWarp<KS>;
  using MX = Mix<NCHOL>;
  static constexpr size_t kRing = (size_t)MX::kStages * 4096 + 2048;  // +pad: KS=32 A-operand overrun
  static constexpr size_t kPlaneBytes = sizeof(float) * ((CW::kPlane + 3) / 4 * 4);
  static constexpr size_t off_planes = kRing;                                   // [kWSlots]
  static constexpr size_t off_g32 = off_planes + MX::kWSlots * kPlaneBytes;
  static constexpr size_t off_bpart = off_g32 + kPlaneBytes;                    // [kBSlots][P][KS]
  static constexpr size_t off_colbuf = off_bpart + sizeof(float) * MX::kBSlots * MX::kProdWarps * KS;
  static constexpr size_t off_bars = off_colbuf + sizeof(float) * MX::kCholWarps * CW::kScratch;
  static constexpr int kNumBars = 2 * MX::kStages + 2 * kAccSlots + 2 * MX::kCholWarps + 2 * MX::kBSlots;
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

template <int KS, int NCHOL>
__global__ void __launch_bounds__(kThreads, 1) row_update_umma_kernel(const RowUpdateParams p) {
  using G = StageGeom<KS>;
  using S = Smem<KS, NCHOL>;
  using CW = CholWarp<KS>;
  using MX = Mix<NCHOL>;
  constexpr int kCholWarps = MX::kCholWarps, kProdWarps = MX::kProdWarps, kFirstProd = MX::kFirstProd;
  constexpr int kStages = MX::kStages, kWSlots = MX::kWSlots, kBSlots = MX::kBSlots;
  constexpr int kRegsProd = MX::kRegsProd;
  constexpr int kReleaseStep = kReleaseDiv > 0 ? KS / kReleaseDiv : -1;
  // 1024-byte alignment (SWIZZLE_128B atoms) comes from the declaration: no integer
  // round-trip on the pointer, so the compiler keeps every access in the shared window
  // (LDS/STS instead of generic loads); checked once below.
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  float* g32 = reinterpret_cast<float*>(smem + S::off_g32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bars);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + kAccSlots;
  uint64_t* w_full = acc_empty + kAccSlots;
  uint64_t* w_empty = w_full + kCholWarps;
  uint64_t* b_full = w_empty + kCholWarps;
  uint64_t* b_empty = b_full + kBSlots;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + S::off_misc);
  constexpr int kPlaneF = (int)(S::kPlaneBytes / sizeof(float));
  constexpr int E = G::kEntries;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int k = p.k;
#ifdef ALS_PROFILE_WAITS
  const long long prof_t0 = clock64();
#endif
  float* planes = reinterpret_cast<float*>(smem + S::off_planes);
  float* bpart = reinterpret_cast<float*>(smem + S::off_bpart);
  if (tid == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // operand atoms need 1024-byte alignment

  if (tid == 0) {
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full[i], 32);  // the 32 lanes of the producer warp that owns the stage
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccSlots; i++) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    // W hand-off barriers are per CONSUMER warp (row u -> barrier u % kCholWarps, data slot
    // u % kWSlots): a parity wait is only sound if its waiter sees every phase, and a slot is
    // refilled kCholWarps/kWSlots times between two visits of the same warp.
    for (int i = 0; i < kCholWarps; i++) {
      mbar_init(&w_full[i], 128);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < kBSlots; i++) {
      mbar_init(&b_full[i], kProdWarps);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init_fence();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_s, kTmemCols);
  // fp32 copy of G in the pair-packed layout of the W planes
  for (int e = tid; e < CW::kPlane; e += kThreads) g32[e] = 0.f;
  __syncthreads();
  for (int e = tid; e < KS * KS; e += kThreads) {
    const int i = e / KS, j = e % KS;  // row i, column j
    if (i >= j && i < k) g32[CW::offP(j >> 1) + 2 * (i - (j & ~1)) + (j & 1)] = (float)p.G[i * KS + j];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_s;
  const long long row_step = gridDim.x;

  if (warp >= kFirstProd) {
   if constexpr (kRegsProd < 96) reg_dealloc<kRegsProd>();
   if (warp < kMmaWarp) {
    // =========================== producers ===========================================
    const int pw = warp - kFirstProd;
    constexpr int CPR = G::kChunksPerRow;  // lanes per factor row
    constexpr int RPP = 32 / CPR;          // rows per pass of the warp
    constexpr int NPASS = E / RPP;         // 8
    const int q = lane % CPR;
    const int sub = lane / CPR;
    // byte offset of this lane's hi slot in a stage for pass i; the lo slot is +1024 (KS=64:
    // next MN atom) or ^64 (KS=32: other half of the 128-byte row, chunk index ^ 4)
    uint32_t offs[NPASS];
#pragma unroll
    for (int i = 0; i < NPASS; i++) {
      uint32_t oh, ol;
      G::slots(sub + RPP * i, q, oh, ol);
      offs[i] = oh;
    }
    uint32_t sbase = 0;  // flat index of stage 0 of the current row
    int useq = 0;        // sequence number of the current row among non-empty rows
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);

    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      // row table of this batch: one row per lane
      long long e0_l = 0;
      int cnt_l = 0;
      {
        const long long myrow = rb + lane * row_step;
        if (myrow < p.n_rows) {
          e0_l = p.row_ptr[myrow];
          cnt_l = (int)(p.row_ptr[myrow + 1] - e0_l);
        }
      }
      const long long left = (p.n_rows - rb + row_step - 1) / row_step;
      const int nb = left < 32 ? (int)left : 32;  // rows in this batch

      auto publish_b = [&]() {
        const int bslot = useq % kBSlots;
        float4 v = bacc;
#pragma unroll
        for (int off = CPR; off < 32; off <<= 1) {
          v.x += __shfl_xor_sync(kFull, v.x, off);
          v.y += __shfl_xor_sync(kFull, v.y, off);
          v.z += __shfl_xor_sync(kFull, v.z, off);
          v.w += __shfl_xor_sync(kFull, v.w, off);
        }
        mbar_wait_id(&b_empty[bslot], (uint32_t)(((useq / kBSlots) & 1) ^ 1), 0);
        if (lane < CPR)
          *reinterpret_cast<float4*>(bpart + (bslot * kProdWarps + pw) * KS + 4 * q) = v;
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_full[bslot]);
        bacc = make_float4(0.f, 0.f, 0.f, 0.f);
      };
      // Cursor over this warp's stages inside the batch: (row index i, stage st, flat base).
      // next_own() moves to the next stage whose flat index is == pw (mod kProdWarps); the
      // consuming cursor publishes the partial b of every non-empty row it leaves.
      int ci = 0, cst = 0;         // consuming cursor
      uint32_t cbase = sbase;
      int li = 0, lst = 0;         // look-ahead cursor (index prefetch)
      uint32_t lbase = sbase;
      auto next_own = [&](int& i, int& st, uint32_t& base, bool consuming) -> bool {
        while (i < nb) {
          const int cnt = __shfl_sync(kFull, cnt_l, i);
          const int nst = (cnt + E - 1) / E;
          const int s = st + (int)(((uint32_t)pw + kProdWarps - (base + (uint32_t)st) % kProdWarps) % kProdWarps);
          if (s < nst) { st = s; return true; }
          if (consuming && cnt > 0) { publish_b(); useq++; }
          base += nst; i++; st = 0;
        }
        return false;
      };
      int ci_n[NPASS];
      auto load_idx = [&](int i, int st) {
        const long long e0 = shfl_i64(e0_l, i);
        const int cnt = __shfl_sync(kFull, cnt_l, i);
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
          const int el = st * E + sub + RPP * ps;
          ci_n[ps] = (el < cnt) ? ld_stream_i32(p.col_idx + e0 + el) : -1;
        }
      };
      bool have = next_own(li, lst, lbase, false);
      if (have) load_idx(li, lst);
      while (next_own(ci, cst, cbase, true)) {
        // here (ci,cst) == (li,lst): its indices are in ci_n
        const long long e0 = shfl_i64(e0_l, ci);
        float4 y[NPASS];
        float r[NPASS];
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
          const bool ok = ci_n[ps] >= 0;
          r[ps] = ok ? ld_stream_f32(p.val + e0 + cst * E + sub + RPP * ps) : 0.f;
          y[ps] = ok ? ldg_f4(p.M + (long long)ci_n[ps] * KS + 4 * q)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        li = ci; lst = cst + 1; lbase = cbase;
        have = next_own(li, lst, lbase, false);
        if (have) load_idx(li, lst);  // indices of my next stage, while the rows are in flight

        const uint32_t sidx = cbase + (uint32_t)cst;
        const int slot = (int)(sidx % kStages);
        mbar_wait_id(&empty[slot], ((sidx / kStages) & 1) ^ 1, 1);
        unsigned char* stage = ring + (size_t)slot * G::kBytes;
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
          const float ar = p.alpha * fabsf(r[ps]);
          // SYRK weight (c_u - 1) = alpha*|r| (ALS.java:471-479); 0 when reconstructing R (:466-469)
          const float s = p.reconstruct_r ? 0.f : sqrt_approx(ar);
          const float cb = p.reconstruct_r ? r[ps] : (r[ps] > 0.f ? 1.f + ar : 0.f);  // :480-482
          uint2 hi, lo;
          split_bf16x2(make_float4(y[ps].x * s, y[ps].y * s, y[ps].z * s, y[ps].w * s), hi, lo);
          const uint32_t oh = offs[ps];
          const uint32_t ol = (KS == 64) ? (oh ^ 32u) : (oh ^ 64u);
          *reinterpret_cast<uint2*>(stage + oh) = hi;
          *reinterpret_cast<uint2*>(stage + ol) = lo;
          bacc.x = fmaf(cb, y[ps].x, bacc.x);
          bacc.y = fmaf(cb, y[ps].y, bacc.y);
          bacc.z = fmaf(cb, y[ps].z, bacc.z);
          bacc.w = fmaf(cb, y[ps].w, bacc.w);
        }
        fence_proxy_async_smem();
        mbar_arrive(&full[slot]);
        cst += 1;
      }
      sbase = cbase;  // next_own has advanced the base past every row of the batch
    }
   } else {
    // =========================== MMA issuer ==========================================
    // The whole warp runs this loop converged (every value below is warp-uniform, so it lives
    // in uniform registers); only the tcgen05 instructions themselves are issued by one
    // elected lane -- always the same one, as tcgen05.commit tracks the MMAs of its own thread.
    // This loop is the serial resource of the CTA: one trip per 16 entries.
    const uint32_t idesc = make_idesc_bf16_mn(G::kM, G::kN);
    const uint64_t desc0 = make_smem_desc(smem_u32(ring), G::kLBO, G::kSBO);  // slot 0, K-step 0
    const uint32_t dhi = (uint32_t)(desc0 >> 32);
    // running ring position: slot, its barriers, the low descriptor word, the phase parity
    uint32_t slot = 0, par = 0;
    uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), dlo = (uint32_t)desc0;
    uint32_t gseg = 0;
    // Row lengths come from loads at warp-uniform addresses (not from a per-lane table and
    // shuffles as in the other roles): only then does the compiler treat the loop as uniform.
    // The next row's length is fetched one row ahead.
    long long row = blockIdx.x;
    int cnt_next = 0;
#ifdef ALS_MMA_LANE0
    if (lane == 0)
#endif
    {
    if (row < p.n_rows) cnt_next = (int)(__ldg(p.row_ptr + row + 1) - __ldg(p.row_ptr + row));
    for (; row < p.n_rows; row += row_step) {
      const int cnt = cnt_next;
      const long long nrow = row + row_step;
      if (nrow < p.n_rows) cnt_next = (int)(__ldg(p.row_ptr + nrow + 1) - __ldg(p.row_ptr + nrow));
      const int nst = (cnt + E - 1) / E;
      for (int st0 = 0; st0 < nst; st0 += kSegStages, gseg++) {
        const int n = (nst - st0 < kSegStages) ? nst - st0 : kSegStages;  // stages of this segment
        const uint32_t a = gseg % kAccSlots;
        mbar_wait_id(&acc_empty[a], ((gseg / kAccSlots) & 1) ^ 1, 2);
        const uint32_t d_tmem = tmem_base + a * (uint32_t)G::kN;
        for (int t = 0; t < n; t++) {
          mbar_wait_addr(full_a, par);
          tc_fence_after_sync();
#pragma unroll
          for (int ks = 0; ks < G::kKSteps; ks++)
            mma_bf16_ss_same_elect(d_tmem, dlo + (uint32_t)(ks * (int)(G::kKStepBytes >> 4)), dhi, idesc,
                                   (t > 0 || ks > 0) ? 1u : 0u);
          mma_commit_addr_elect(empty_a);  // frees the operand stage once the MMAs have read it
          slot++; full_a += 8; empty_a += 8; dlo += (uint32_t)(G::kBytes >> 4);
          if (slot == kStages) {
            slot = 0; par ^= 1u;
            full_a -= 8 * kStages; empty_a -= 8 * kStages; dlo -= (uint32_t)(kStages * (G::kBytes >> 4));
          }
        }
        mma_commit_elect(&acc_full[a]);  // accumulator segment complete
      }
    }
    }
#ifdef ALS_MMA_LANE0
    __syncwarp();
#endif
   }
  } else if (warp < kDrainWarps) {
    // =========================== drain warpgroup =====================================
    reg_dealloc<kRegsDrain>();
   if constexpr (KS == 64) {
    // Warp q owns TMEM lane quarter q = matrix rows 16q..16q+15: lanes 0..15 hold the rows'
    // hi operand halves, lanes 16..31 the lo halves; D's column group g (32 columns) holds
    // [hi | lo] of features 16g..16g+15 (StageGeom<64>::slots).  Per 16-feature chunk a lane
    // folds the two column halves, swaps eight values with its partner lane (xor 16) so that
    // the pair finishes eight columns each, adds G (+ lambda*alpha*n_u on the diagonal) and
    // stores four float2 of the pair-packed plane.  The four warps never wait for each other.
    const int qd = warp;
    const bool upper = lane >= 16;
    const int i = 16 * qd + (lane & 15);  // matrix row of this lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t gseg = 0;
    int useq = 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      int cnt_l = 0;
      {
        const long long myrow = rb + lane * row_step;
        if (myrow < p.n_rows) cnt_l = (int)(p.row_ptr[myrow + 1] - p.row_ptr[myrow]);
      }
      const long long left = (p.n_rows - rb + row_step - 1) / row_step;
      const int nb = left < 32 ? (int)left : 32;
      for (int ib = 0; ib < nb; ib++) {
        const int cnt = __shfl_sync(kFull, cnt_l, ib);
        if (cnt == 0) continue;
        const int nst = (cnt + E - 1) / E;
        const int nseg = (nst + kSegStages - 1) / kSegStages;
        const int ws = useq % kWSlots;
        float* plane = planes + ws * kPlaneF;
        // W = G + lambda*alpha*n_u*I + ... (ALS.java:447-450, 488-492); padding rows (i >= k)
        // get a unit diagonal so the factorisation stays finite
        const float lam_n = (i < k) ? (float)(p.lambda_alpha * (double)cnt) : 1.f;
        // slot ws last held row useq - kWSlots, consumed by the same Cholesky warp
        mbar_wait_id(&w_empty[ws], (uint32_t)(((useq / kWSlots) & 1) ^ 1), 4);
        for (int seg = 0; seg < nseg; seg++, gseg++) {
          const int a = (int)(gseg % kAccSlots);
          mbar_wait_id(&acc_full[a], (gseg / kAccSlots) & 1, 5);
          tc_fence_after_sync();
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(a * G::kN);
          const float* src = (seg == 0) ? g32 : plane;  // later segments add to what this lane stored
          const float lam = (seg == 0) ? lam_n : 0.f;
#pragma unroll
          for (int fc = 0; fc < 4; fc++) {
            if (fc > qd) break;  // warp-uniform: chunk entirely right of the diagonal
            uint32_t v0[16], v1[16];
            tmem_ld_32x16(taddr + 32 * fc, v0);
            tmem_ld_32x16(taddr + 32 * fc + 16, v1);
            tmem_wait_ld();
            float r[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const float wl = __uint_as_float(v0[e]) + __uint_as_float(v1[e]);          // feature 16fc+e
              const float wu = __uint_as_float(v0[8 + e]) + __uint_as_float(v1[8 + e]);  // feature 16fc+8+e
              r[e] = (upper ? wu : wl) + __shfl_xor_sync(kFull, upper ? wl : wu, 16);
            }
            // this lane now holds W[i][j0 .. j0+7]
            const int P0 = 8 * fc + (upper ? 4 : 0);
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
              const int P = P0 + pp;  // column pair (2P, 2P+1)
              const int o = 2 * (KS * P - P * (P - 1)) + 2 * (i - 2 * P);
              if (fc < qd) {          // strictly below the diagonal: no predicates
                const float2 g = *reinterpret_cast<const float2*>(src + o);
                *reinterpret_cast<float2*>(plane + o) = make_float2(r[2 * pp] + g.x, r[2 * pp + 1] + g.y);
              } else if (i >= 2 * P) {
                const float2 g = *reinterpret_cast<const float2*>(src + o);
                float2 v = make_float2(r[2 * pp] + g.x, r[2 * pp + 1] + g.y);
                if (i == 2 * P) { v.x += lam; v.y = 0.f; }
                else if (i == 2 * P + 1) v.y += lam;
                *reinterpret_cast<float2*>(plane + o) = v;
              }
            }
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[a]);  // accumulator may be overwritten by the next segment
        }
        mbar_arrive(&w_full[ws]);  // 128 arrivals: the plane of this row is complete
        useq++;
      }
    }
   } else {
    const int t = tid;  // 0..127 == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool is_hi = t < KS, is_lo = t >= KS && t < 2 * KS;
    const int i = is_hi ? t : t - KS;  // matrix row held by this thread
    const int warp_max_row = (warp * 32 + 31) % KS;  // warp-uniform: lower triangle only
    uint32_t gseg = 0;
    int useq = 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      int cnt_l = 0;
      {
        const long long myrow = rb + lane * row_step;
        if (myrow < p.n_rows) cnt_l = (int)(p.row_ptr[myrow + 1] - p.row_ptr[myrow]);
      }
      const long long left = (p.n_rows - rb + row_step - 1) / row_step;
      const int nb = left < 32 ? (int)left : 32;
      for (int ib = 0; ib < nb; ib++) {
        const int cnt = __shfl_sync(kFull, cnt_l, ib);
        if (cnt == 0) continue;
        const int nst = (cnt + E - 1) / E;
        const int nseg = (nst + kSegStages - 1) / kSegStages;
        const int ws = useq % kWSlots;
        float* plane = planes + ws * kPlaneF;
        const float lam_n = (float)(p.lambda_alpha * (double)cnt);
        // slot ws last held row useq - kWSlots, consumed by the same Cholesky warp
        mbar_wait_id(&w_empty[ws], (uint32_t)(((useq / kWSlots) & 1) ^ 1), 4);
        for (int seg = 0; seg < nseg; seg++, gseg++) {
          const int a = (int)(gseg % kAccSlots);
          mbar_wait_id(&acc_full[a], (gseg / kAccSlots) & 1, 5);
          tc_fence_after_sync();
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(a * G::kN);
          // hi-row warps write (W = G + lambda*alpha*n_u*I + hi part), then the lo-row warps add
          // their part: one warpgroup barrier between the two (whole warps are hi or lo)
          if (!is_hi) bar_sync(2, 128);
#pragma unroll 1
          for (int jc = 0; jc < KS / 16; jc++) {
            if (jc * 16 > warp_max_row) break;  // chunk entirely above the diagonal for this warp
            uint32_t v0[16], v1[16];
            tmem_ld_32x16(taddr + jc * 16, v0);
            tmem_ld_32x16(taddr + KS + jc * 16, v1);
            tmem_wait_ld();
            if (is_hi || is_lo) {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 2) {
                const int j = jc * 16 + jj;  // even column: pair P = j/2 holds columns j, j+1
                if (i >= j) {
                  float2 v;
                  v.x = __uint_as_float(v0[jj]) + __uint_as_float(v1[jj]);
                  v.y = (i > j) ? __uint_as_float(v0[jj + 1]) + __uint_as_float(v1[jj + 1]) : 0.f;
                  const int P = j >> 1;
                  const int o = 2 * (KS * P - P * (P - 1)) + 2 * (i - j);
                  if (is_hi && seg == 0) {
                    // W = G + lambda*alpha*n_u*I + ... (ALS.java:447-450, 488-492); padding rows
                    // (i >= k) get a unit diagonal so the factorisation stays finite
                    const float2 g = *reinterpret_cast<const float2*>(g32 + o);
                    v.x += g.x + ((i == j) ? (i < k ? lam_n : 1.f) : 0.f);
                    v.y += g.y + ((i == j + 1) ? (i < k ? lam_n : 1.f) : 0.f);
                  } else {
                    const float2 old = *reinterpret_cast<const float2*>(plane + o);
                    v.x += old.x;
                    v.y += old.y;
                  }
                  *reinterpret_cast<float2*>(plane + o) = v;
                }
              }
            }
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[a]);  // accumulator may be overwritten by the next segment
          if (is_hi) bar_sync(2, 128);    // releases the lo-row warps of this segment
          if (seg + 1 < nseg) bar_sync(3, 128);  // next segment's hi pass reads what lo wrote
        }
        mbar_arrive(&w_full[ws]);  // 128 arrivals: the plane of this row is complete
        useq++;
      }
    }
   }
  } else {
    // =========================== Cholesky warps ======================================
    reg_alloc<kRegsChol>();
    const int cw = warp - kFirstChol;
    float* scratch = reinterpret_cast<float*>(smem + S::off_colbuf) + cw * CW::kScratch;
    int useq = 0;
    for (long long rb = blockIdx.x; rb < p.n_rows; rb += 32 * row_step) {
      int cnt_l = 0;
      {
        const long long myrow = rb + lane * row_step;
        if (myrow < p.n_rows) cnt_l = (int)(p.row_ptr[myrow + 1] - p.row_ptr[myrow]);
      }
      const long long left = (p.n_rows - rb + row_step - 1) / row_step;
      const int nb = left < 32 ? (int)left : 32;
      for (int ib = 0; ib < nb; ib++) {
        const int cnt = __shfl_sync(kFull, cnt_l, ib);
        if (cnt == 0) continue;
        if (useq % kCholWarps != cw) { useq++; continue; }
        const long long row = rb + ib * row_step;
        const int ws = useq % kWSlots;
        mbar_wait_id(&w_full[cw], (uint32_t)((useq / kCholWarps) & 1), 6);
        typename CW::Rows R;
        CW::load(planes + ws * kPlaneF, lane, R);
        __syncwarp();
        // The W slot and the rhs slot are handed back to the drain / producers only when the
        // sweep below reaches column kReleaseStep: the refill they trigger would otherwise
        // compete with the first, shared-memory-heaviest third of the sweep.
        if (kReleaseStep < 0 && lane == 0) mbar_arrive(&w_empty[cw]);
        mbar_wait_id(&b_full[cw], (uint32_t)((useq / kBSlots) & 1), 7);
        float bx[CW::kRowsPerLane];  // rhs entries of this lane's rows, then the solution
#pragma unroll
        for (int s = 0; s < CW::kRowsPerLane; s++) bx[s] = 0.f;
#pragma unroll
        for (int w = 0; w < kProdWarps; w++) {
          const float* bp = bpart + (cw * kProdWarps + w) * KS;
#pragma unroll
          for (int s = 0; s < CW::kRowsPerLane; s++) bx[s] += bp[CW::row_of(lane, s)];
        }
        __syncwarp();
        if (kReleaseStep < 0 && lane == 0) mbar_arrive(&b_empty[cw]);
        // All Cholesky warps enter the sweep together: they then run the same instruction
        // stream in near lockstep, so the instruction cache sees one stream instead of eight.
        if (kCholLockstep) {
#ifdef ALS_PROFILE_WAITS
          const long long tb = clock64();
          bar_sync(1, kCholWarps * 32);
          if (lane == 0) atomicAdd(&g_wait_cycles[9], (unsigned long long)(clock64() - tb));
#else
          bar_sync(1, kCholWarps * 32);
#endif
        }
#ifdef ALS_PROFILE_WAITS
        const long long ts = clock64();
#endif
        const bool ok = CW::template factor_solve<kReleaseStep>(
            R, scratch, bx, p.threshold, kCondLimit, lane, k, [&]() {
              if (lane == 0) {
                mbar_arrive(&w_empty[cw]);
                mbar_arrive(&b_empty[cw]);
              }
            });
#ifdef ALS_PROFILE_WAITS
        if (lane == 0) atomicAdd(&g_wait_cycles[10], (unsigned long long)(clock64() - ts));
#endif
        if (ok) {
          float* dst = p.out + (p.row_offset + row) * KS;
#pragma unroll
          for (int s = 0; s < CW::kRowsPerLane; s++) {
            const int r = CW::row_of(lane, s);
            if (CW::writes(lane, s) && r < k) {
              dst[r] = bx[s];
              push_to_peers(p, (p.row_offset + row) * KS + r, bx[s]);
            }
          }
        } else if (lane == 0) {
          const int slot = atomicAdd(p.retry_count, 1);
          p.retry_rows[slot] = (int)row;
        }
        useq++;
      }
    }
    // tail: warps without a row in the last round still meet the others at the barrier
    if (kCholLockstep && (useq % kCholWarps) != 0 && cw >= (useq % kCholWarps))
      bar_sync(1, kCholWarps * 32);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
#ifdef ALS_PROFILE_WAITS
  if (tid == 0) atomicAdd(&g_wait_cycles[8], (unsigned long long)(clock64() - prof_t0));
#endif
}

}  // namespace umma

inline bool umma_supported(int ks) { return ks == 32 || ks == 64; }

template <int KS, int NCHOL>
inline int launch_row_update_umma_t(const RowUpdateParams& p, int sm_count, cudaStream_t stream,
                                    char* err, size_t err_len) {
  // (the dynamic shared memory opt-in is per device: als_create does it for the handle's device)
  using S = umma::Smem<KS, NCHOL>;
  long long grid = sm_count;
  if (grid > p.n_rows) grid = p.n_rows > 0 ? p.n_rows : 1;
  umma::row_update_umma_kernel<KS, NCHOL><<<(int)grid, umma::kThreads, S::kTotal, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, err_len, "row_update_umma launch: %s", cudaGetErrorString(e));
    return ALS_E_CUDA;
  }
  return ALS_OK;
}

// Average entries per row at which the producer-heavy role mix wins: with 8 Cholesky / 7
// producer warps a row costs max(solve/8, stages * stage_time/7), with 4 / 11 it costs
// max(solve/4, stages * stage_time/11); measured solve ~30k cycles, stage_time ~2.5k cycles
// (16 entries) => crossover near 340 entries per row.
constexpr long long kLongRowEntries = 384;

// long_rows: producer-heavy role mix (4 Cholesky + 11 producer warps) instead of 8 + 7.
inline int launch_row_update_umma(int ks, const RowUpdateParams& p, bool long_rows, int sm_count,
                                  cudaStream_t stream, char* err, size_t err_len) {
  switch (ks) {
    case 32:
      return long_rows ? launch_row_update_umma_t<32, 4>(p, sm_count, stream, err, err_len)
                       : launch_row_update_umma_t<32, 8>(p, sm_count, stream, err, err_len);
    case 64:
      return long_rows ? launch_row_update_umma_t<64, 4>(p, sm_count, stream, err, err_len)
                       : launch_row_update_umma_t<64, 8>(p, sm_count, stream, err, err_len);
    default:
      snprintf(err, err_len, "tcgen05 kernel supports padded feature counts 32 and 64 only");
      return ALS_E_UNSUPPORTED;
  }
}

}  // namespace als
