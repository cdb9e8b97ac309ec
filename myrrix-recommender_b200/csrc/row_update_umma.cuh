// row_update_umma.cuh -- tensor-core (tcgen05 + TMEM) row update for padded feature
// counts 32 and 64.
//
// Same contract as row_update_simt.cuh (Worker.call, AlternatingLeastSquares.java:438-502),
// different machine mapping.  One persistent CTA per SM, rows strided over CTAs, three
// warp-specialised roles connected by mbarrier rings:
//
//   producers (8 warps)  gather the row's factor rows from HBM (coalesced 16-byte loads),
//                        scale each by sqrt(alpha*|r|), split it into bf16 hi + bf16 lo
//                        (x ~= hi+lo to 2^-17, round-to-nearest twice) and store both
//                        halves into the swizzled MN-major operand stage; accumulate the
//                        right-hand side b_u in fp32 on the side.
//   MMA issuer (1 thread) per 16 entries one tcgen05.mma, A = B = [hi;lo]:
//                        D[2k x 2k] += [hi;lo] [hi;lo]^T  (fp32 accumulate in TMEM), i.e. all
//                        four cross products hi*hi, hi*lo, lo*hi, lo*lo of the rank-16 update.
//   epilogue (2 warpgroups) read D from TMEM, fold the four blocks into the fp64 W_u held
//                        in shared memory (W_u starts as G + lambda*alpha*n_u*I), then fp64
//                        LDL^T + solves and write the fp32 row.
//
// Long rows are cut into segments of kSegStages stages so no fp32 accumulator ever carries
// more than 512 (k=64) / 1024 (k=32) entries before it is folded into fp64.
#pragma once
#include "common.cuh"
#include "row_update_simt.cuh"  // RowUpdateParams
#include "solve_fp64.cuh"
#include "umma_common.cuh"

namespace als {
namespace umma {

constexpr int kEpiWGs = 2;                  // warps 0..7
constexpr int kProdWarps = 8;               // warps 8..15
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMmaWarp = kEpiWGs * 4 + kProdWarps;  // warp 16
constexpr int kThreads = (kMmaWarp + 1) * 32;       // 544
constexpr int kStages = 16;                 // operand ring depth (4 KB each)
constexpr int kAccSlots = 4;                // TMEM accumulators in flight
constexpr int kBSlots = 8;                  // right-hand-side ring depth (rows in flight)
constexpr int kSegStages = 32;              // stages per accumulation segment
constexpr int kTmemCols = 512;

template <int KS>
struct Smem {
  static constexpr int LDW = KS + 1;
  static constexpr size_t kRing = (size_t)kStages * 4096 + 2048;  // +pad: KS=32 A-operand overrun
  static constexpr size_t kW = sizeof(double) * KS * LDW;
  static constexpr size_t off_W = kRing;
  static constexpr size_t off_bvec = off_W + kEpiWGs * kW;
  static constexpr size_t off_invd = off_bvec + sizeof(double) * kEpiWGs * KS;
  static constexpr size_t off_bpart = off_invd + sizeof(double) * kEpiWGs * KS;
  static constexpr size_t off_bars = off_bpart + sizeof(float) * kBSlots * kProdWarps * KS;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccSlots + 2 * kBSlots;
  static constexpr size_t off_misc = off_bars + sizeof(uint64_t) * kNumBars;
  static constexpr size_t kTotal = off_misc + 64 + 1024;  // + slack for 1024-byte alignment
};

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) row_update_umma_kernel(const RowUpdateParams p) {
  using G = StageGeom<KS>;
  using S = Smem<KS>;
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::off_bars);
  uint64_t* full = bars;
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_empty = acc_full + kAccSlots;
  uint64_t* b_full = acc_empty + kAccSlots;
  uint64_t* b_empty = b_full + kBSlots;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + S::off_misc);
  float* bpart = reinterpret_cast<float*>(smem + S::off_bpart);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full[i], kProdThreads);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kAccSlots; i++) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    for (int i = 0; i < kBSlots; i++) {
      mbar_init(&b_full[i], kProdWarps);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init_fence();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_s, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_s;

  const long long row_step = gridDim.x;

  if (warp >= kEpiWGs * 4 && warp < kMmaWarp) {
    // =========================== producers ===========================================
    const int pt = tid - kEpiWGs * 128;  // 0..255
    const int q = pt % G::kChunksPerRow;  // which float4 of the factor row
    const int el = pt / G::kChunksPerRow; // entry slot inside a stage
    const int pw = pt >> 5;
    uint32_t off_hi, off_lo;
    G::slots(el, q, off_hi, off_lo);
    long long sidx = 0;
    int useq = 0;
    for (long long row = blockIdx.x; row < p.n_rows; row += row_step) {
      const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
      if (e1 == e0) continue;
      const int bslot = useq % kBSlots;
      mbar_wait(&b_empty[bslot], ((useq / kBSlots) & 1) ^ 1);
      float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
      const long long nst = (e1 - e0 + G::kEntries - 1) / G::kEntries;
      // software pipeline: the gather for stage st+1 is in flight while stage st is split/stored
      float4 y_next = make_float4(0.f, 0.f, 0.f, 0.f);
      float r_next = 0.f;
      {
        const long long e = e0 + el;
        if (e < e1) {
          const int ci = ld_stream_i32(p.col_idx + e);
          r_next = ld_stream_f32(p.val + e);
          y_next = ldg_f4(p.M + (long long)ci * KS + 4 * q);
        }
      }
      for (long long st = 0; st < nst; st++, sidx++) {
        const float4 y = y_next;
        const float r = r_next;
        y_next = make_float4(0.f, 0.f, 0.f, 0.f);
        r_next = 0.f;
        {
          const long long e = e0 + (st + 1) * G::kEntries + el;
          if (e < e1) {
            const int ci = ld_stream_i32(p.col_idx + e);
            r_next = ld_stream_f32(p.val + e);
            y_next = ldg_f4(p.M + (long long)ci * KS + 4 * q);
          }
        }
        const float ar = p.alpha * fabsf(r);
        // SYRK weight (c_u - 1) = alpha*|r| (ALS.java:471-479); 0 when reconstructing R (:466-469)
        const float s = p.reconstruct_r ? 0.f : sqrtf(ar);
        const float cb = p.reconstruct_r ? r : (r > 0.f ? 1.f + ar : 0.f);  // :480-482
        uint2 hi, lo;
        split_bf16x2(make_float4(y.x * s, y.y * s, y.z * s, y.w * s), hi, lo);
        const int slot = (int)(sidx % kStages);
        mbar_wait(&empty[slot], (uint32_t)(((sidx / kStages) & 1) ^ 1));
        unsigned char* stage = ring + (size_t)slot * G::kBytes;
        *reinterpret_cast<uint2*>(stage + off_hi) = hi;
        *reinterpret_cast<uint2*>(stage + off_lo) = lo;
        fence_proxy_async_smem();
        mbar_arrive(&full[slot]);
        bacc.x = fmaf(cb, y.x, bacc.x);
        bacc.y = fmaf(cb, y.y, bacc.y);
        bacc.z = fmaf(cb, y.z, bacc.z);
        bacc.w = fmaf(cb, y.w, bacc.w);
      }
      // reduce b over the lanes of this warp that hold the same chunk q, then publish the
      // per-warp partial; the epilogue sums the 8 partials in fp64.
#pragma unroll
      for (int off = G::kChunksPerRow; off < 32; off <<= 1) {
        bacc.x += __shfl_xor_sync(0xffffffffu, bacc.x, off);
        bacc.y += __shfl_xor_sync(0xffffffffu, bacc.y, off);
        bacc.z += __shfl_xor_sync(0xffffffffu, bacc.z, off);
        bacc.w += __shfl_xor_sync(0xffffffffu, bacc.w, off);
      }
      if (lane < G::kChunksPerRow)
        *reinterpret_cast<float4*>(bpart + ((size_t)bslot * kProdWarps + pw) * KS + 4 * q) = bacc;
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bslot]);
      useq++;
    }
  } else if (warp == kMmaWarp) {
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
  } else {
    // =========================== epilogue warpgroups =================================
    const int g = warp >> 2;        // warpgroup index
    const int t = tid & 127;        // thread in warpgroup == TMEM lane
    const int bar_id = 1 + g;
    double* W = reinterpret_cast<double*>(smem + S::off_W + (size_t)g * S::kW);
    double* bvec = reinterpret_cast<double*>(smem + S::off_bvec) + g * KS;
    double* invd = reinterpret_cast<double*>(smem + S::off_invd) + g * KS;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int k = p.k;
    long long gseg = 0;
    int useq = 0;
    for (long long row = blockIdx.x; row < p.n_rows; row += row_step) {
      const long long e0 = p.row_ptr[row], e1 = p.row_ptr[row + 1];
      if (e1 == e0) continue;
      const long long nst = (e1 - e0 + G::kEntries - 1) / G::kEntries;
      const long long nseg = (nst + kSegStages - 1) / kSegStages;
      if (useq % kEpiWGs != g) {
        gseg += nseg;
        useq++;
        continue;
      }
      // W = G + lambda*alpha*n_u*I, lower triangle, fp64 (ALS.java:447-450, 488-492)
      group_barrier(bar_id);  // previous row's solve has finished with W / bvec
      for (int e = t; e < KS * KS; e += 128) {
        const int r = e / KS, c = e % KS;
        if (c <= r) {
          double v = (r < k) ? p.G[r * KS + c] : 0.0;
          if (r == c && r < k) v += p.lambda_alpha * (double)(e1 - e0);
          W[r * S::LDW + c] = v;
        }
      }
      group_barrier(bar_id);
      for (long long seg = 0; seg < nseg; seg++, gseg++) {
        const int a = (int)(gseg % kAccSlots);
        mbar_wait(&acc_full[a], (uint32_t)((gseg / kAccSlots) & 1));
        tc_fence_after_sync();
        // thread t holds row t of D = [hi;lo][hi;lo]^T: fold the two column halves
        float rsum[KS];
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)(a * G::kN);
#pragma unroll
        for (int jc = 0; jc < KS / 16; jc++) {
          uint32_t v0[16], v1[16];
          tmem_ld_32x16(taddr + jc * 16, v0);
          tmem_ld_32x16(taddr + KS + jc * 16, v1);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j++)
            rsum[jc * 16 + j] = __uint_as_float(v0[j]) + __uint_as_float(v1[j]);
        }
        tc_fence_before_sync();
        mbar_arrive(&acc_empty[a]);  // accumulator may be overwritten by the next segment
        // rows [0,KS) are the hi half, [KS,2KS) the lo half: two phases so the two threads
        // that own the same W row never collide
        if (t < KS) {
#pragma unroll
          for (int j = 0; j < KS; j++)
            if (j <= t) W[t * S::LDW + j] += (double)rsum[j];
        }
        group_barrier(bar_id);
        if (t >= KS && t < 2 * KS) {
          const int i = t - KS;
#pragma unroll
          for (int j = 0; j < KS; j++)
            if (j <= i) W[i * S::LDW + j] += (double)rsum[j];
        }
        group_barrier(bar_id);
      }
      const int bslot = useq % kBSlots;
      mbar_wait(&b_full[bslot], (uint32_t)((useq / kBSlots) & 1));
      if (t < KS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kProdWarps; w++)
          s += (double)bpart[((size_t)bslot * kProdWarps + w) * KS + t];
        bvec[t] = s;
      }
      group_barrier(bar_id);
      if (t == 0) mbar_arrive(&b_empty[bslot]);  // b partials consumed
      ldlt_solve_fp64<KS>(W, bvec, invd, k, t, bar_id, (double)p.threshold, p.status, p.which,
                          p.row_offset + row, p.out + (p.row_offset + row) * KS);
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
