// row_update_umma.cuh -- tensor-core (tcgen05 + TMEM) row update for padded feature
// counts 32 and 64.
//
// Same contract as row_update_simt.cuh (Worker.call, AlternatingLeastSquares.java:438-502),
// different machine mapping.  One persistent CTA (640 threads) per SM, rows strided over
// CTAs, four warp-specialised roles connected by mbarrier rings:
//
//   producers (7 warps)  each owns every 7th 16/32-entry stage of the CTA's flat stage
//                        stream: gather the factor rows from HBM (coalesced 16-byte loads),
//                        scale by sqrt(alpha*|r|), split into bf16 hi + bf16 lo (x ~= hi+lo to
//                        2^-17, round-to-nearest twice), store both halves into the swizzled
//                        MN-major operand stage; accumulate the rhs b_u in fp32 on the side.
//   MMA issuer (1 thread) per 16 entries one tcgen05.mma with A = B = [hi;lo]:
//                        D[2k x 2k] += [hi;lo][hi;lo]^T, fp32 accumulate in TMEM (all four
//                        cross products hi*hi, hi*lo, lo*hi, lo*lo of the rank-16 update).
//   drain (1 warpgroup)  TMEM -> registers, fold the column halves, add G (fp32 copy in smem)
//                        and lambda*alpha*n_u, write two packed lower-triangular planes
//                        (hi rows / lo rows) into a W slot in shared memory.
//   Cholesky (8 warps)   each takes one row's W slot: in-register fp32 LDL^T + solves
//                        (chol_warp.cuh), writes the fp32 factor row.  Rows whose pivots look
//                        singular / ill-conditioned are appended to a retry list and re-solved
//                        in fp64 by the CUDA-core kernel (which owns the error reporting).
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
constexpr int kCholWarps = 8;                                // warps 4..11
constexpr int kProdWarps = 7;                                // warps 12..18
constexpr int kFirstChol = kDrainWarps;
constexpr int kFirstProd = kFirstChol + kCholWarps;
constexpr int kMmaWarp = kFirstProd + kProdWarps;            // warp 19
constexpr int kThreads = (kMmaWarp + 1) * 32;                // 640
constexpr int kStages = 16;                                  // operand ring depth (4 KB each)
constexpr int kAccSlots = 4;                                 // TMEM accumulators in flight
constexpr int kWSlots = 4;                                   // W slots (drain -> Cholesky)
constexpr int kBSlots = 8;                                   // rhs ring depth (== kCholWarps)
static_assert(kBSlots == kCholWarps && kCholWarps % kWSlots == 0, "ring/consumer phase bookkeeping");
constexpr int kSegStages = 64;                               // stages per accumulation segment
constexpr int kTmemCols = 512;
constexpr int kRegsProd = 80, kRegsDrain = 56, kRegsChol = 128;  // 256*80+128*56+256*128 <= 61440
constexpr float kCondLimit = 256.f;  // max diag / min pivot above which a row goes to fp64

template <int KS>
struct Smem {
  using CW = CholWarp<KS>;
  static constexpr size_t kRing = (size_t)kStages * 4096 + 2048;  // +pad: KS=32 A-operand overrun
  static constexpr size_t kPlaneBytes = sizeof(float) * ((CW::kPlane + 3) / 4 * 4);
  static constexpr size_t off_planes = kRing;                                   // [kWSlots][2]
  static constexpr size_t off_g32 = off_planes + kWSlots * 2 * kPlaneBytes;
  static constexpr size_t off_bpart = off_g32 + kPlaneBytes;                    // [kBSlots][P][KS]
  static constexpr size_t off_colbuf = off_bpart + sizeof(float) * kBSlots * kProdWarps * KS;
  static constexpr size_t off_bars = off_colbuf + sizeof(float) * kCholWarps * 2 * CW::kColBuf;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccSlots + 2 * kCholWarps + 2 * kBSlots;
  static constexpr size_t off_misc = (off_bars + sizeof(uint64_t) * kNumBars + 15) / 16 * 16;
  static constexpr size_t kTotal = off_misc + 64 + 1024;  // + slack for 1024-byte alignment
};

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) row_update_umma_kernel(const RowUpdateParams p) {
  using G = StageGeom<KS>;
  using S = Smem<KS>;
  using CW = CholWarp<KS>;
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  float* planes = reinterpret_cast<float*>(smem + S::off_planes);
  float* g32 = reinterpret_cast<float*>(smem + S::off_g32);
  float* bpart = reinterpret_cast<float*>(smem + S::off_bpart);
  float* colbufs = reinterpret_cast<float*>(smem + S::off_colbuf);
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

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int k = p.k;

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
  // fp32 copy of G (packed lower, column-major) for the drain warps
  for (int e = tid; e < KS * KS; e += kThreads) {
    const int j = e / KS, i = e % KS;  // column j, row i
    if (i >= j) g32[CW::off(j) + i - j] = (i < k) ? (float)p.G[i * KS + j] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_s;
  const long long row_step = gridDim.x;

  // Register budget (640 threads x 96 at launch = 61440): the Cholesky warpgroups hold a
  // k x k triangle in registers and take 128 each; drain and producer/MMA warpgroups give
  // registers back.  setmaxnreg is warpgroup-wide: roles are laid out on 4-warp boundaries.
  if (warp >= kFirstProd) {
   reg_dealloc<kRegsProd>();
   if (warp < kMmaWarp) {
    // =========================== producers ===========================================
    const int pw = warp - kFirstProd;
    constexpr int CPR = G::kChunksPerRow;  // lanes per factor row
    constexpr int RPP = 32 / CPR;          // rows per pass of the warp
    constexpr int NPASS = G::kEntries / RPP;
    const int q = lane % CPR;
    const int sub = lane / CPR;
    struct Cursor {
      long long row, e0, e1, nst, st, sidx;
      int useq;
    };
    auto open_row = [&](Cursor& c) {
      while (c.row < p.n_rows) {
        c.e0 = p.row_ptr[c.row];
        c.e1 = p.row_ptr[c.row + 1];
        if (c.e1 > c.e0) break;
        c.row += row_step;
      }
      if (c.row < p.n_rows) c.nst = (c.e1 - c.e0 + G::kEntries - 1) / G::kEntries;
    };
    Cursor cur;
    cur.row = blockIdx.x; cur.st = 0; cur.sidx = 0; cur.useq = 0; cur.e0 = cur.e1 = cur.nst = 0;
    open_row(cur);
    Cursor la = cur;
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto publish_b = [&](int useq) {
      const int bslot = useq % kBSlots;
      float4 v = bacc;
#pragma unroll
      for (int off = CPR; off < 32; off <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
      }
      mbar_wait(&b_empty[bslot], (uint32_t)(((useq / kBSlots) & 1) ^ 1));
      if (lane < CPR)
        *reinterpret_cast<float4*>(bpart + ((size_t)bslot * kProdWarps + pw) * KS + 4 * q) = v;
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bslot]);
      bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // advance c to this warp's next stage (flat index == pw mod kProdWarps); the consuming
    // cursor publishes the partial b of every row it leaves (also rows it owns no stage of).
    auto seek = [&](Cursor& c, bool consuming) -> bool {
      for (;;) {
        if (c.row >= p.n_rows) return false;
        const long long base = c.sidx - c.st;  // flat index of stage 0 of this row
        const long long st = c.st + ((pw - (base + c.st)) % kProdWarps + kProdWarps) % kProdWarps;
        if (st < c.nst) {
          c.sidx = base + st;
          c.st = st;
          return true;
        }
        if (consuming) publish_b(c.useq);
        c.sidx = base + c.nst;
        c.st = 0;
        c.useq++;
        c.row += row_step;
        open_row(c);
      }
    };
    int ci_n[NPASS];
    float r_n[NPASS];
    auto load_meta = [&](const Cursor& c) {
#pragma unroll
      for (int i = 0; i < NPASS; i++) {
        const long long e = c.e0 + c.st * G::kEntries + sub + RPP * i;
        const bool ok = e < c.e1;
        ci_n[i] = ok ? ld_stream_i32(p.col_idx + e) : -1;
        r_n[i] = ok ? ld_stream_f32(p.val + e) : 0.f;
      }
    };
    bool have_la = seek(la, false);
    if (have_la) load_meta(la);
    while (seek(cur, true)) {
      float4 y[NPASS];
      float r[NPASS];
#pragma unroll
      for (int i = 0; i < NPASS; i++) {
        r[i] = r_n[i];
        y[i] = (ci_n[i] >= 0) ? ldg_f4(p.M + (long long)ci_n[i] * KS + 4 * q)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      la.st += 1;
      la.sidx += 1;
      have_la = seek(la, false);
      if (have_la) load_meta(la);  // (index,value) of my next stage, while the rows are in flight

      const int slot = (int)(cur.sidx % kStages);
      mbar_wait(&empty[slot], (uint32_t)(((cur.sidx / kStages) & 1) ^ 1));
      unsigned char* stage = ring + (size_t)slot * G::kBytes;
#pragma unroll
      for (int i = 0; i < NPASS; i++) {
        const float ar = p.alpha * fabsf(r[i]);
        // SYRK weight (c_u - 1) = alpha*|r| (ALS.java:471-479); 0 when reconstructing R (:466-469)
        const float s = p.reconstruct_r ? 0.f : sqrtf(ar);
        const float cb = p.reconstruct_r ? r[i] : (r[i] > 0.f ? 1.f + ar : 0.f);  // :480-482
        uint2 hi, lo;
        split_bf16x2(make_float4(y[i].x * s, y[i].y * s, y[i].z * s, y[i].w * s), hi, lo);
        uint32_t off_hi, off_lo;
        G::slots(sub + RPP * i, q, off_hi, off_lo);
        *reinterpret_cast<uint2*>(stage + off_hi) = hi;
        *reinterpret_cast<uint2*>(stage + off_lo) = lo;
        bacc.x = fmaf(cb, y[i].x, bacc.x);
        bacc.y = fmaf(cb, y[i].y, bacc.y);
        bacc.z = fmaf(cb, y[i].z, bacc.z);
        bacc.w = fmaf(cb, y[i].w, bacc.w);
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[slot]);
      cur.st += 1;
      cur.sidx += 1;
    }
   } else {
    // =========================== MMA issuer ==========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_mn(G::kM, G::kN);
      long long sidx = 0, gseg = 0;
      for (long long row = blockIdx.x; row < p.n_rows; row += row_step) {
        const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
        if (e1 == e0) continue;
        const long long nst = (e1 - e0 + G::kEntries - 1) / G::kEntries;
        uint32_t d_tmem = 0;
        int a = 0;
        for (long long st = 0; st < nst; st++, sidx++) {
          const bool seg_first = (st % kSegStages) == 0;
          if (seg_first) {
            a = (int)(gseg % kAccSlots);
            mbar_wait(&acc_empty[a], (uint32_t)(((gseg / kAccSlots) & 1) ^ 1));
            tc_fence_after_sync();
            d_tmem = tmem_base + (uint32_t)(a * G::kN);
          }
          const int slot = (int)(sidx % kStages);
          mbar_wait(&full[slot], (uint32_t)((sidx / kStages) & 1));
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(ring + (size_t)slot * G::kBytes);
#pragma unroll
          for (int ks = 0; ks < G::kKSteps; ks++) {
            const uint64_t desc = make_smem_desc(sa + ks * G::kKStepBytes, G::kLBO, G::kSBO);
            mma_bf16_ss(d_tmem, desc, desc, idesc, (seg_first && ks == 0) ? 0u : 1u);
          }
          mma_commit(&empty[slot]);  // frees the operand stage once the MMAs have read it
          if ((st % kSegStages) == kSegStages - 1 || st == nst - 1) {
            mma_commit(&acc_full[a]);  // accumulator segment complete
            gseg++;
          }
        }
      }
    }
   }
  } else if (warp < kDrainWarps) {
    // =========================== drain warpgroup =====================================
    reg_dealloc<kRegsDrain>();
    const int t = tid;  // 0..127 == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool is_hi = t < KS, is_lo = t >= KS && t < 2 * KS;
    const int i = is_hi ? t : t - KS;  // matrix row held by this thread
    // warp-uniform: highest matrix row any lane of this warp holds (lower triangle only)
    const int warp_max_row = (warp * 32 + 31) % KS;
    long long gseg = 0;
    int useq = 0;
    for (long long row = blockIdx.x; row < p.n_rows; row += row_step) {
      const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
      if (e1 == e0) continue;
      const long long nst = (e1 - e0 + G::kEntries - 1) / G::kEntries;
      const long long nseg = (nst + kSegStages - 1) / kSegStages;
      const int ws = useq % kWSlots;
      float* plane = planes + (size_t)(ws * 2 + (is_lo ? 1 : 0)) * kPlaneF;
      const float lam_n = (float)(p.lambda_alpha * (double)(e1 - e0));
      if (useq >= kWSlots) {  // slot last held row useq - kWSlots: wait until it was loaded
        const int prev = useq - kWSlots;
        mbar_wait(&w_empty[prev % kCholWarps], (uint32_t)((prev / kCholWarps) & 1));
      }
      for (long long seg = 0; seg < nseg; seg++, gseg++) {
        const int a = (int)(gseg % kAccSlots);
        mbar_wait(&acc_full[a], (uint32_t)((gseg / kAccSlots) & 1));
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)(a * G::kN);
#pragma unroll
        for (int jc = 0; jc < KS / 16; jc++) {
          if (jc * 16 > warp_max_row) continue;  // whole chunk above the diagonal for this warp
          uint32_t v0[16], v1[16];
          tmem_ld_32x16(taddr + jc * 16, v0);
          tmem_ld_32x16(taddr + KS + jc * 16, v1);
          tmem_wait_ld();
          if (is_hi || is_lo) {
#pragma unroll
            for (int jj = 0; jj < 16; jj++) {
              const int j = jc * 16 + jj;
              if (j <= i) {
                float v = __uint_as_float(v0[jj]) + __uint_as_float(v1[jj]);
                const int o = CW::off(j) + i - j;
                if (seg == 0) {
                  // W = G + lambda*alpha*n_u*I + ... (ALS.java:447-450, 488-492): hi plane only
                  if (is_hi) v += g32[o] + ((i == j && i < k) ? lam_n : 0.f);
                } else {
                  v += plane[o];
                }
                plane[o] = v;
              }
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&acc_empty[a]);  // accumulator may be overwritten by the next segment
      }
      mbar_arrive(&w_full[useq % kCholWarps]);  // 128 arrivals: both planes of this row complete
      useq++;
    }
  } else {
    // =========================== Cholesky warps ======================================
    reg_alloc<kRegsChol>();
    const int cw = warp - kFirstChol;
    const uint32_t colbuf = smem_u32(colbufs + (size_t)cw * 2 * CW::kColBuf);
    const uint32_t planes_s = smem_u32(planes);
    const uint32_t bpart_s = smem_u32(bpart);
    int useq = 0;
    for (long long row = blockIdx.x; row < p.n_rows; row += row_step) {
      const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
      if (e1 == e0) continue;
      if (useq % kCholWarps != cw) { useq++; continue; }
      const int ws = useq % kWSlots;
      const int bslot = useq % kBSlots;
      mbar_wait(&w_full[cw], (uint32_t)((useq / kCholWarps) & 1));
      typename CW::Rows R;
      CW::load(planes_s + (uint32_t)((ws * 2) * kPlaneF) * 4u,
               planes_s + (uint32_t)((ws * 2 + 1) * kPlaneF) * 4u, lane, R);
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_empty[cw]);  // slot free: the row now lives in registers
      mbar_wait(&b_full[bslot], (uint32_t)((useq / kBSlots) & 1));
      float b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int w = 0; w < kProdWarps; w++) {
        const uint32_t bp = bpart_s + (uint32_t)(((bslot * kProdWarps + w) * KS + lane) * 4);
        b0 += lds_f32(bp);
        if (KS == 64) b1 += lds_f32(bp + 128u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[bslot]);
      // padding rows (k < KS) have a zero diagonal: give them a unit pivot so the sweep stays
      // finite; their solution entries are exactly 0 and are never written.
      if (k < KS) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j == lane && j >= k) { if (j & 1) R.A0[j >> 1].y = 1.f; else R.A0[j >> 1].x = 1.f; }
        if (KS == 64) {
#pragma unroll
          for (int j = 32; j < 64; j++)
            if (j == lane + 32 && j >= k) { if (j & 1) R.A1[j >> 1].y = 1.f; else R.A1[j >> 1].x = 1.f; }
        }
      }
      float x0, x1;
      const bool ok = CW::factor_solve(R, colbuf, b0, b1, p.threshold, kCondLimit, lane, k, x0, x1);
      if (ok) {
        float* dst = p.out + (p.row_offset + row) * KS;
        if (lane < k) dst[lane] = x0;
        if (KS == 64 && lane + 32 < k) dst[lane + 32] = x1;
      } else if (lane == 0) {
        const int slot = atomicAdd(p.retry_count, 1);
        p.retry_rows[slot] = (int)row;
      }
      useq++;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace umma

inline bool umma_supported(int ks) { return ks == 32 || ks == 64; }

template <int KS>
inline int launch_row_update_umma_t(const RowUpdateParams& p, int sm_count, cudaStream_t stream,
                                    char* err, size_t err_len) {
  using S = umma::Smem<KS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma::row_update_umma_kernel<KS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)S::kTotal);
    if (e != cudaSuccess) {
      snprintf(err, err_len, "cudaFuncSetAttribute(umma): %s", cudaGetErrorString(e));
      return ALS_E_CUDA;
    }
    configured = true;
  }
  long long grid = sm_count;
  if (grid > p.n_rows) grid = p.n_rows > 0 ? p.n_rows : 1;
  umma::row_update_umma_kernel<KS><<<(int)grid, umma::kThreads, S::kTotal, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, err_len, "row_update_umma launch: %s", cudaGetErrorString(e));
    return ALS_E_CUDA;
  }
  return ALS_OK;
}

inline int launch_row_update_umma(int ks, const RowUpdateParams& p, int sm_count,
                                  cudaStream_t stream, char* err, size_t err_len) {
  switch (ks) {
    case 32: return launch_row_update_umma_t<32>(p, sm_count, stream, err, err_len);
    case 64: return launch_row_update_umma_t<64>(p, sm_count, stream, err, err_len);
    default:
      snprintf(err, err_len, "tcgen05 kernel supports padded feature counts 32 and 64 only");
      return ALS_E_UNSUPPORTED;
  }
}

}  // namespace als
