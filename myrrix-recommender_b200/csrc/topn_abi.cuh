// topn_abi.cuh -- C ABI of the top-N scoring path (include/myrrix_als.h: als_recommend,
// als_recommend_batch, als_top_n) over csrc/topn.cuh.  Included by als_abi.cu after the handle.
#pragma once
#include "topn.cuh"

struct TopNState {
  unsigned* excl = nullptr;   // [kMaxQ][words] filtered-item bitmaps of the pass
  unsigned* tmp = nullptr;    // [words] scratch bitmap (known-item intersection)
  long long words = 0;
  unsigned long long* cand = nullptr;
  size_t cand_cap = 0;
  unsigned long long* gthr = nullptr;  // [kMaxQ]
  int* nonfinite = nullptr;
  int* ints = nullptr;        // qrow | vec_query of a pass, or the users of a batch
  size_t ints_cap = 0;
  float* vecs = nullptr;      // [kMaxVec][ks] staging of caller-supplied feature vectors
  int* pack = nullptr;        // single-query results in one read-back: items[kMaxN] | values[kMaxN] | count | nonfinite
  int* out_items = nullptr;
  float* out_values = nullptr;
  int* out_counts = nullptr;
  size_t out_cap = 0;         // queries
  int out_n = 0;              // how_many the output buffers were sized for
  CUtensorMap map;
  const float* map_base = nullptr;
  long long map_rows = 0;
  int grid = 0;
  unsigned attr_mask = 0;     // instantiations whose shared-memory opt-in is done
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

void topn_free(als_handle* h) {
  TopNState* t = h->topn;
  if (!t) return;
  cudaFree(t->excl); cudaFree(t->tmp); cudaFree(t->cand); cudaFree(t->gthr); cudaFree(t->nonfinite);
  cudaFree(t->ints); cudaFree(t->vecs); cudaFree(t->pack); cudaFree(t->out_items); cudaFree(t->out_values); cudaFree(t->out_counts);
  delete t;
  h->topn = nullptr;
}

template <typename T>
int topn_grow(als_handle* h, T** p, size_t* cap, size_t need) {
  if (*cap >= need && *p) return ALS_OK;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t n = need < 1024 ? 1024 : need;
  if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return fail(h, ALS_E_OOM, "top-N scratch of %zu bytes", n * sizeof(T));
  }
  *cap = n;
  return ALS_OK;
}

template <int KS, int G>
int topn_launch_t(als_handle* h, TopNState* t, const topn::Params& p, int slot) {
  using S = topn::Shape<KS>;
  auto kern = topn::topn_score_kernel<KS, G>;
  if (!(t->attr_mask & (1u << slot))) {
    CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal));
    t->attr_mask |= 1u << slot;
  }
  const long long n_tiles = (p.n_items + topn::kTile - 1) / topn::kTile;
  int per_sm = 1;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, topn::kThreads, S::kTotal));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  long long grid = (long long)h->sm_count * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > t->grid) return fail(h, ALS_E_STATE, "top-N candidate buffer sized for %d CTAs", t->grid);
  kern<<<(int)grid, topn::kThreads, S::kTotal, h->stream>>>(t->map, p);
  CU(h, cudaGetLastError());
  topn::topn_merge_kernel<<<p.n_q, topn::kMergeThreads, 0, h->stream>>>(p.cand, (int)grid, p.gthr, p.how_many,
                                                                        t->out_items, t->out_values, t->out_counts);
  CU(h, cudaGetLastError());
  h->launches += 2;
  return ALS_OK;
}

template <int KS>
int topn_launch_g(als_handle* h, TopNState* t, const topn::Params& p, int ks_slot) {
  const int g = p.n_vec >= 8 ? 8 : (p.n_vec > 2 ? 4 : p.n_vec);
  switch (g) {
    case 1: return topn_launch_t<KS, 1>(h, t, p, ks_slot * 4 + 0);
    case 2: return topn_launch_t<KS, 2>(h, t, p, ks_slot * 4 + 1);
    case 4: return topn_launch_t<KS, 4>(h, t, p, ks_slot * 4 + 2);
    default: return topn_launch_t<KS, 8>(h, t, p, ks_slot * 4 + 3);
  }
}

// out_items / out_values / out_counts of the state receive the pass's n_q results at their start;
// the caller offsets them afterwards (batch) or copies them out.
int topn_launch(als_handle* h, TopNState* t, const topn::Params& p) {
  switch (h->ks) {
    case 4: return topn_launch_g<4>(h, t, p, 0);
    case 8: return topn_launch_g<8>(h, t, p, 1);
    case 16: return topn_launch_g<16>(h, t, p, 2);
    case 32: return topn_launch_g<32>(h, t, p, 3);
    case 64: return topn_launch_g<64>(h, t, p, 4);
    case 128: return topn_launch_g<128>(h, t, p, 5);
  }
  return fail(h, ALS_E_UNSUPPORTED, "top-N: padded features %d", h->ks);
}

// scratch + tensor map for scoring the rows of F ([rows][ks], resident)
int topn_prepare(als_handle* h, const float* F, long long rows, int how_many, size_t n_queries_out) {
  if (!F || rows <= 0) return fail(h, ALS_E_STATE, "top-N: no resident factors (set interactions / factors first)");
  if (how_many < 1 || how_many > topn::kMaxN)
    return fail(h, ALS_E_UNSUPPORTED, "top-N: howMany must be in 1..%d", topn::kMaxN);
  if (rows >= 0x7fffffffLL) return fail(h, ALS_E_UNSUPPORTED, "top-N: more than 2^31 - 1 rows");
  if (!h->topn) {
    h->topn = new (std::nothrow) TopNState();
    if (!h->topn) return fail(h, ALS_E_OOM, "top-N state");
  }
  TopNState* t = h->topn;
  const long long words = (rows + 31) / 32;
  if (words > t->words) {
    cudaFree(t->excl); cudaFree(t->tmp);
    t->excl = t->tmp = nullptr;
    t->words = 0;
    if (cudaMalloc(&t->excl, sizeof(unsigned) * (size_t)words * topn::kMaxQ) != cudaSuccess ||
        cudaMalloc(&t->tmp, sizeof(unsigned) * (size_t)words) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, ALS_E_OOM, "top-N filter bitmaps");
    }
    t->words = words;
    CU(h, cudaMemsetAsync(t->tmp, 0, sizeof(unsigned) * (size_t)words, h->stream));
  }
  if (!t->gthr) {
    CU(h, cudaMalloc(&t->gthr, sizeof(unsigned long long) * topn::kMaxQ));
    CU(h, cudaMalloc(&t->nonfinite, sizeof(int)));
    CU(h, cudaMalloc(&t->vecs, sizeof(float) * (size_t)topn::kMaxVec * kMaxFeatures));
    CU(h, cudaMalloc(&t->pack, sizeof(int) * (2 * topn::kMaxN + 2)));
  }
  t->grid = h->sm_count * 2;
  int rc;
  if ((rc = topn_grow(h, &t->cand, &t->cand_cap, (size_t)topn::kMaxQ * t->grid * how_many)) != ALS_OK) return rc;
  if (t->out_cap < n_queries_out || t->out_n < how_many) {
    cudaFree(t->out_items); cudaFree(t->out_values); cudaFree(t->out_counts);
    t->out_items = nullptr; t->out_values = nullptr; t->out_counts = nullptr;
    t->out_cap = 0;
    const size_t nq = n_queries_out < 64 ? 64 : n_queries_out;
    if (cudaMalloc(&t->out_items, sizeof(int) * nq * how_many) != cudaSuccess ||
        cudaMalloc(&t->out_values, sizeof(float) * nq * how_many) != cudaSuccess ||
        cudaMalloc(&t->out_counts, sizeof(int) * nq) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, ALS_E_OOM, "top-N result buffers");
    }
    t->out_cap = nq;
    t->out_n = how_many;
  }
  if (t->map_base != F || t->map_rows != rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(h, ALS_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int box_cols = h->ks < 32 ? h->ks : 32;
    const cuuint64_t gdim[2] = {(cuuint64_t)h->ks, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)h->ks * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)topn::kTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&t->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)F, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           h->ks >= 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, ALS_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    t->map_base = F;
    t->map_rows = rows;
  }
  return ALS_OK;
}

int topn_check_nonfinite(als_handle* h, TopNState* t) {
  int nf = 0;
  CU(h, cudaMemcpyAsync(&nf, t->nonfinite, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (nf) return fail(h, ALS_E_NONFINITE, "Bad recommendation value (RecommendIterator.java:99)");
  return ALS_OK;
}

}  // namespace

// One query: the mean score of `n_vec` feature vectors against every row of the item (which = 1)
// or user (which = 0) factor.  features: host, [n_vec][k].  rows_of_x >= 0 entries take row
// rows_of_x[v] of X instead (features may then be null).
static int topn_one_query(als_handle* h, int32_t which, const float* features, const int32_t* rows_of_x,
                          int32_t n_vec, const int32_t* exclude, int32_t n_exclude, bool exclude_known,
                          int32_t how_many, int32_t* out_ids, float* out_values, int32_t* out_count) {
  if (!h || !out_ids || !out_values || !out_count || (which != 0 && which != 1)) return ALS_E_ARG;
  if (n_vec < 1 || n_vec > topn::kMaxVec)
    return fail(h, ALS_E_UNSUPPORTED, "top-N: 1..%d feature vectors per query", topn::kMaxVec);
  if (n_exclude < 0 || (n_exclude > 0 && !exclude)) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  const float* F = which == 0 ? h->X : h->Y;
  const long long rows = which == 0 ? h->n_users : h->n_items;
  int rc = topn_prepare(h, F, rows, how_many, 1);
  if (rc != ALS_OK) return rc;
  TopNState* t = h->topn;
  for (int i = 0; i < n_exclude; i++)
    if (exclude[i] < 0 || exclude[i] >= rows) return fail(h, ALS_E_ARG, "excluded id %d out of range", exclude[i]);
  // vectors
  const size_t n_ints = (size_t)2 * topn::kMaxVec + (size_t)n_exclude;
  if ((rc = topn_grow(h, &t->ints, &t->ints_cap, n_ints)) != ALS_OK) return rc;
  int hi[2 * topn::kMaxVec];
  memset(hi, 0, sizeof(hi));  // vec_query: all vectors belong to query 0
  const float* qbase;
  if (rows_of_x) {
    for (int v = 0; v < n_vec; v++) {
      if (rows_of_x[v] < 0 || rows_of_x[v] >= h->n_users) return fail(h, ALS_E_ARG, "user %d out of range", rows_of_x[v]);
      hi[v] = rows_of_x[v];
    }
    qbase = h->X;
  } else {
    if (!features) return ALS_E_ARG;
    float hv[topn::kMaxVec * kMaxFeatures];
    memset(hv, 0, sizeof(hv));
    for (int v = 0; v < n_vec; v++) {
      for (int f = 0; f < h->k; f++) hv[v * h->ks + f] = features[(size_t)v * h->k + f];
      hi[v] = v;
    }
    CU(h, cudaMemcpyAsync(t->vecs, hv, sizeof(float) * (size_t)n_vec * h->ks, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));  // hv is on this frame
    qbase = t->vecs;
  }
  CU(h, cudaMemcpyAsync(t->ints, hi, sizeof(hi), cudaMemcpyHostToDevice, h->stream));
  // filter bitmap
  bool filtered = false;
  if (exclude_known || n_exclude > 0) CU(h, cudaMemsetAsync(t->excl, 0, sizeof(unsigned) * (size_t)t->words, h->stream));
  if (exclude_known) {
    // the intersection of the known-item sets of the users that have one (ServerRecommender.java:402-421)
    if (!rows_of_x || which != 1) return fail(h, ALS_E_ARG, "known items are defined for user queries against items");
    const Csr& R = h->by_user;
    bool first = true;
    if (n_vec == 1) {
      // one user: mark its row (a row without entries marks nothing) -- no need to know its length here
      const long long row = (long long)rows_of_x[0] - R.row_begin;
      if (row < 0 || row >= R.rows)
        return fail(h, ALS_E_ARG, "user %d is not in this rank's block: its known items live on another rank", rows_of_x[0]);
      topn::mark_row_kernel<<<8, 256, 0, h->stream>>>(R.ptr, R.idx, row, t->excl);
      h->launches += 1;
      filtered = true;
    }
    for (int v = 0; v < n_vec && n_vec > 1; v++) {
      const long long row = (long long)rows_of_x[v] - R.row_begin;
      if (row < 0 || row >= R.rows)
        return fail(h, ALS_E_ARG, "user %d is not in this rank's block: its known items live on another rank", rows_of_x[v]);
      long long pp[2];
      CU(h, cudaMemcpyAsync(pp, R.ptr + row, sizeof(pp), cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
      if (pp[1] == pp[0]) continue;  // (knownItemIDs.get(userID) == null -> continue)
      const int blocks = (int)((pp[1] - pp[0] + 255) / 256 > 64 ? 64 : (pp[1] - pp[0] + 255) / 256);
      if (first) {
        topn::mark_row_kernel<<<blocks, 256, 0, h->stream>>>(R.ptr, R.idx, row, t->excl);
        first = false;
      } else {
        topn::mark_row_kernel<<<blocks, 256, 0, h->stream>>>(R.ptr, R.idx, row, t->tmp);
        topn::and_clear_kernel<<<(int)((t->words + 255) / 256), 256, 0, h->stream>>>(t->excl, t->tmp, t->words);
        h->launches += 1;
      }
      h->launches += 1;
      filtered = true;
    }
  }
  if (n_exclude > 0) {
    int* d_ex = t->ints + 2 * topn::kMaxVec;
    CU(h, cudaMemcpyAsync(d_ex, exclude, sizeof(int) * (size_t)n_exclude, cudaMemcpyHostToDevice, h->stream));
    topn::mark_list_kernel<<<(n_exclude + 255) / 256, 256, 0, h->stream>>>(d_ex, n_exclude, t->excl, t->words, 1);
    h->launches += 1;
    filtered = true;
  }
  CU(h, cudaMemsetAsync(t->gthr, 0, sizeof(unsigned long long) * topn::kMaxQ, h->stream));
  CU(h, cudaMemsetAsync(t->pack + 2 * topn::kMaxN, 0, 2 * sizeof(int), h->stream));
  topn::Params p;
  p.qbase = qbase;
  p.qrow = t->ints;
  p.vec_query = t->ints + topn::kMaxVec;
  p.n_vec = n_vec;
  p.n_q = 1;
  p.excl = filtered ? t->excl : nullptr;
  p.excl_words = t->words;
  p.n_items = rows;
  p.how_many = how_many;
  p.gthr = t->gthr;
  p.cand = t->cand;
  p.nonfinite = t->pack + 2 * topn::kMaxN + 1;
  // the pass's results land in the packed buffer: one read-back, one synchronisation
  int* const oi = t->out_items;
  float* const ov = t->out_values;
  int* const oc = t->out_counts;
  t->out_items = t->pack;
  t->out_values = reinterpret_cast<float*>(t->pack + topn::kMaxN);
  t->out_counts = t->pack + 2 * topn::kMaxN;
  nvtxRangePushA("als:top_n");
  rc = topn_launch(h, t, p);
  nvtxRangePop();
  t->out_items = oi;
  t->out_values = ov;
  t->out_counts = oc;
  if (rc != ALS_OK) return rc;
  int host[2 * topn::kMaxN + 2];
  CU(h, cudaMemcpyAsync(host, t->pack, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (host[2 * topn::kMaxN + 1]) return fail(h, ALS_E_NONFINITE, "Bad recommendation value (RecommendIterator.java:99)");
  memcpy(out_ids, host, sizeof(int) * (size_t)how_many);
  memcpy(out_values, host + topn::kMaxN, sizeof(float) * (size_t)how_many);
  *out_count = host[2 * topn::kMaxN];
  return ALS_OK;
}

int als_top_n(als_handle* h, int32_t which, const float* features, int32_t n_vectors, const int32_t* exclude,
              int32_t n_exclude, int32_t how_many, int32_t* out_ids, float* out_values, int32_t* out_count) {
  return topn_one_query(h, which, features, nullptr, n_vectors, exclude, n_exclude, false, how_many, out_ids,
                        out_values, out_count);
}

int als_recommend(als_handle* h, const int32_t* users, int32_t n_users, int32_t how_many,
                  int32_t consider_known_items, const int32_t* exclude, int32_t n_exclude, int32_t* out_items,
                  float* out_values, int32_t* out_count) {
  if (!h || !users) return ALS_E_ARG;
  return topn_one_query(h, 1, nullptr, users, n_users, exclude, n_exclude, consider_known_items == 0, how_many,
                        out_items, out_values, out_count);
}

int als_recommend_batch(als_handle* h, const int32_t* users, int64_t n_queries, int32_t how_many,
                        int32_t consider_known_items, int32_t* out_items, float* out_values, int32_t* out_counts) {
  if (!h || n_queries < 0 || (n_queries > 0 && (!users || !out_items || !out_values || !out_counts))) return ALS_E_ARG;
  if (n_queries == 0) return ALS_OK;
  CU(h, cudaSetDevice(h->device));
  const long long chunk_max = 1 << 16;  // queries per result read-back
  int rc = topn_prepare(h, h->Y, h->n_items, how_many, (size_t)(n_queries < chunk_max ? n_queries : chunk_max));
  if (rc != ALS_OK) return rc;
  TopNState* t = h->topn;
  const Csr& R = h->by_user;
  for (int64_t i = 0; i < n_queries; i++) {
    if (users[i] < 0 || users[i] >= h->n_users) return fail(h, ALS_E_ARG, "user %d out of range", users[i]);
    if (!consider_known_items && (users[i] < R.row_begin || users[i] >= R.row_begin + R.rows))
      return fail(h, ALS_E_ARG, "user %d is not in this rank's block: its known items live on another rank", users[i]);
  }
  CU(h, cudaMemsetAsync(t->nonfinite, 0, sizeof(int), h->stream));
  nvtxRangePushA("als:recommend_batch");
  for (int64_t c0 = 0; c0 < n_queries; c0 += chunk_max) {
    const int64_t cn = (n_queries - c0 < chunk_max) ? n_queries - c0 : chunk_max;
    if ((rc = topn_grow(h, &t->ints, &t->ints_cap, (size_t)cn)) != ALS_OK) { nvtxRangePop(); return rc; }
    CU(h, cudaMemcpyAsync(t->ints, users + c0, sizeof(int) * (size_t)cn, cudaMemcpyHostToDevice, h->stream));
    int* oi = t->out_items;
    float* ov = t->out_values;
    int* oc = t->out_counts;
    for (int64_t q0 = 0; q0 < cn; q0 += topn::kMaxQ) {
      const int nq = (int)((cn - q0 < topn::kMaxQ) ? cn - q0 : topn::kMaxQ);
      if (!consider_known_items) {
        CU(h, cudaMemsetAsync(t->excl, 0, sizeof(unsigned) * (size_t)t->words * nq, h->stream));
        topn::mark_rows_kernel<<<dim3(8, nq), 256, 0, h->stream>>>(R.ptr, R.idx, t->ints + q0, R.row_begin, R.rows,
                                                                     t->excl, t->words);
        h->launches += 1;
      }
      CU(h, cudaMemsetAsync(t->gthr, 0, sizeof(unsigned long long) * topn::kMaxQ, h->stream));
      topn::Params p;
      p.qbase = h->X;
      p.qrow = t->ints + q0;
      p.vec_query = nullptr;  // vector v = query v
      p.n_vec = nq;
      p.n_q = nq;
      p.excl = consider_known_items ? nullptr : t->excl;
      p.excl_words = t->words;
      p.n_items = h->n_items;
      p.how_many = how_many;
      p.gthr = t->gthr;
      p.cand = t->cand;
      p.nonfinite = t->nonfinite;
      // results of the pass land at the start of the state's buffers: point them at this pass
      t->out_items = oi + q0 * how_many;
      t->out_values = ov + q0 * how_many;
      t->out_counts = oc + q0;
      rc = topn_launch(h, t, p);
      t->out_items = oi;
      t->out_values = ov;
      t->out_counts = oc;
      if (rc != ALS_OK) { nvtxRangePop(); return rc; }
    }
    CU(h, cudaMemcpyAsync(out_items + c0 * how_many, oi, sizeof(int) * (size_t)cn * how_many, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(out_values + c0 * how_many, ov, sizeof(float) * (size_t)cn * how_many, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(out_counts + c0, oc, sizeof(int) * (size_t)cn, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  nvtxRangePop();
  return topn_check_nonfinite(h, t);
}
