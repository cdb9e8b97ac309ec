// als_abi.cu -- the C ABI (include/myrrix_als.h) over the sm_100a kernels.
//
// Plain CUDA runtime: no torch types anywhere in this library.  One handle owns
// all HBM: both orientations of R (CSR by user, CSR by item), the two factor
// matrices with a padded row stride, the fp64 Gramian and its partials.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "aux_kernels.cuh"
#include "foldin_dev.cuh"
#include "common.cuh"
#include "gramian.cuh"
#include "row_update_simt.cuh"
#include "row_update_umma.cuh"
#include "row_update_v2.cuh"
#include "resolve_fp64.cuh"
#include <nvtx3/nvToolsExt.h>

using namespace als;

// ---------------------------------------------------------------------------
// NCCL is resolved at run time (dlopen) so the library loads on a box without it
// and a single-GPU caller never pays for it.
#include <nccl.h>
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
bool load_nccl() {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return false;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.Send = (decltype(g_nccl.Send))dlsym(g_nccl.lib, "ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(g_nccl.lib, "ncclRecv");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather &&
              g_nccl.AllReduce && g_nccl.GetErrorString && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart &&
              g_nccl.GroupEnd;
  return g_nccl.ok;
}
}  // namespace

// ---------------------------------------------------------------------------
struct Csr {
  long long* ptr = nullptr;  // [rows+1]
  int* idx = nullptr;        // [nnz]
  float* val = nullptr;      // [nnz]
  long long rows = 0;        // local rows
  long long nnz = 0;
  long long row_begin = 0;   // global index of local row 0
};

#include <cuda.h>  // CUtensorMap

struct TopNState;  // topn_abi.cuh

struct als_handle {
  als_config cfg;
  TopNState* topn = nullptr;  // top-N scoring scratch, created on first use
  // TMA tensor maps of the factor matrices for the row update's gathers ([rows][ks] fp32, one box = one row)
  CUtensorMap gather_map[2];
  const float* gather_map_base[2] = {nullptr, nullptr};
  long long gather_map_rows[2] = {0, 0};
  // visiting order of the row updates per orientation (rows by length, longest first), built lazily;
  // invalidated whenever an interaction array is replaced (free_csr)
  int* d_order[2] = {nullptr, nullptr};
  long long order_rows[2] = {0, 0};
  bool order_valid[2] = {false, false};
  // long rows split into chunks for the tensor-core kernel (ensure_split): virtual rows per orientation
  struct Split {
    bool valid = false;      // built for the current interaction arrays
    long long rows = 0;      // rows of the orientation it was built for
    long long n_virtual = 0; // virtual rows (= rows when nothing is split)
    long long n_acc = 0;     // split rows (0: the orientation is walked as stored)
    long long* vptr = nullptr;
    int* vrow = nullptr;
    int* vacc = nullptr;
    int* acc_chunks = nullptr;
    float* gacc = nullptr;
    int* gcount = nullptr;
    size_t acc_floats = 0;   // floats of one accumulation record
  } split[2];
  long long split_limit = 8192;  // entries; MYRRIX_ALS_SPLIT_ROWS (0: never split)
  // stash mode of the tensor-core kernel (RowUpdateParams::stash), per orientation: switched on when the
  // previous launch of the same half handed many rows to the fp64 paths, off again when it no longer does
  int stash_policy = -1;          // MYRRIX_ALS_STASH: 0 never, 1 always, unset: by the previous launch's count
  bool stash_on[2] = {false, false};
  bool fail_pending[2] = {false, false};
  cudaEvent_t fail_ev[2] = {nullptr, nullptr};
  int* h_fail = nullptr;          // pinned: [which][retry count, resolve count] of the last launch
  float* d_stash = nullptr;       // [CTA][solving warp][record]
  size_t stash_floats = 0;
  float* d_resolve_buf = nullptr;
  int* d_resolve_rows = nullptr;
  int* d_resolve_count = nullptr;
  long long resolve_cap = 0;
  size_t resolve_rec = 0;         // floats per record
  // fold-in solver state of the generation (als_set_fold_in_state): [0] X'X, [1] Y'Y
  double* fi_qrt[2] = {nullptr, nullptr};
  double* fi_rdiag[2] = {nullptr, nullptr};
  int* fi_perm[2] = {nullptr, nullptr};
  double fi_learn_rate = 1.0;
  int k = 0, ks = 0;
  int kernel = ALS_KERNEL_SIMT;
  int device = 0;
  int sm_count = 0;
  long long n_users = 0, n_items = 0;      // global
  long long users_alloc = 0, items_alloc = 0;  // rows allocated (padded to world blocks)
  Csr by_user, by_item;
  bool have_by_item = false;
  float* X = nullptr;
  float* Y = nullptr;
  double* G = nullptr;
  double* G_partial = nullptr;
  int g_partials = 0;
  DeviceStatus* d_status = nullptr;
  unsigned long long* d_ticket = nullptr;
  int* d_retry_rows = nullptr;   // rows the tensor-core kernel handed to the fp64 kernel
  long long retry_cap = 0;
  int* d_retry_count = nullptr;
  long long* d_retry_total = nullptr;
  double* d_scratch = nullptr;  // rank kernel scratch + probe output
  int* d_rank = nullptr;
  // rows that are keys of the reference's map but have no entries (InputFilesReader.removeSmall
  // leaves the empty maps in place): local row indices per orientation, solved as W = G, b = 0
  int* d_pe_rows[2] = {nullptr, nullptr};
  int* d_pe_count[2] = {nullptr, nullptr};
  long long pe_rows[2] = {0, 0};
  // per-handle launch configuration (function attributes and occupancy are per device)
  int simt_blocks_per_sm = 1;
  int mix_override = 0;   // MYRRIX_ALS_MIX=4|8, read once at als_create
  bool legacy_umma = false;  // MYRRIX_ALS_V1=1: round-1 tensor-core kernel (A/B runs)
  bool no_row_order = false; // MYRRIX_ALS_NO_ROW_ORDER=1: visit rows as stored (A/B runs)
  // probe scratch (grown on demand, freed with the handle)
  int* d_probe_idx = nullptr;
  double* d_probe_out = nullptr;
  long long probe_idx_cap = 0, probe_out_cap = 0;
  int* d_flag = nullptr;  // validation / cross-rank status word
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  long long device_bytes = 0;
  // multi-GPU
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  // peer replicas of X and Y (cudaIpc mappings; entry `rank` is unused): finished rows are pushed
  // into them from the solve epilogue.  p2p = false -> exchange by ncclAllGather after the kernel.
  bool p2p = false, p2p_disabled = false;
  bool factors_ipc = false;  // X / Y are plain cudaMalloc allocations (mappable by peers), not pool memory
  float* peer_X[16] = {nullptr};
  float* peer_Y[16] = {nullptr};
  // profiling
  bool profile = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  als_timings tm;
  struct PendingEv { cudaEvent_t a, b; int kind; };
  PendingEv* pending = nullptr;
  int n_pending = 0, cap_pending = 0;
  int launches = 0;
  // errors
  char err[512];
  int singular_rank = 0;
  int sticky = ALS_OK;
};

namespace {

int fail(als_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(h, expr)                                                                       \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      const int _c = (_e == cudaErrorMemoryAllocation) ? ALS_E_OOM : ALS_E_CUDA;          \
      return fail((h), _c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                  __FILE__, __LINE__);                                                    \
    }                                                                                     \
  } while (0)

// Working storage (interaction arrays, transposition scratch, small state) comes from the device's
// stream-ordered memory pool, whose release threshold als_create raises to "never": a serving
// process rebuilds the model again and again, and cudaMalloc / cudaFree of tens of GB cost ~8 ms per
// GB each time (measured: 150 ms to tear down one C3 handle, 12 % of a 5-iteration call).
template <typename T>
int dev_alloc(als_handle* h, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
#ifdef ALS_NO_POOL  // (A/B builds)
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
#else
  cudaError_t e = cudaMallocAsync((void**)p, count * sizeof(T), h->stream);
#endif
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, ALS_E_OOM, "cudaMallocAsync of %zu bytes failed: %s", count * sizeof(T),
                cudaGetErrorString(e));
  }
  h->device_bytes += (long long)(count * sizeof(T));
  return ALS_OK;
}
template <typename T>
void dev_free(als_handle* h, T** p, size_t count) {
  if (*p) {
#ifdef ALS_NO_POOL
    cudaFree(*p);
#else
    cudaFreeAsync(*p, h->stream);
#endif
    h->device_bytes -= (long long)((count ? count : 1) * sizeof(T));
    *p = nullptr;
  }
}
// The factor replicas are mapped by peer processes (cudaIpc): plain cudaMalloc.
template <typename T>
int dev_alloc_ipc(als_handle* h, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, ALS_E_OOM, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T),
                cudaGetErrorString(e));
  }
  h->device_bytes += (long long)(count * sizeof(T));
  return ALS_OK;
}
template <typename T>
void dev_free_ipc(als_handle* h, T** p, size_t count) {
  if (*p) {
    cudaFree(*p);
    h->device_bytes -= (long long)((count ? count : 1) * sizeof(T));
    *p = nullptr;
  }
}

void free_split(als_handle* h, int which) {
  als_handle::Split& s = h->split[which];
  dev_free(h, &s.vptr, (size_t)s.n_virtual + 1);
  dev_free(h, &s.vrow, (size_t)s.n_virtual);
  dev_free(h, &s.vacc, (size_t)s.n_virtual);
  dev_free(h, &s.acc_chunks, (size_t)s.n_acc);
  dev_free(h, &s.gacc, (size_t)s.n_acc * s.acc_floats);
  dev_free(h, &s.gcount, (size_t)s.n_acc);
  s = als_handle::Split();
}

void free_csr(als_handle* h, Csr* c) {
  h->order_valid[0] = h->order_valid[1] = false;
  h->split[0].valid = h->split[1].valid = false;
  dev_free(h, &c->ptr, (size_t)c->rows + 1);
  dev_free(h, &c->idx, (size_t)c->nnz);
  dev_free(h, &c->val, (size_t)c->nnz);
  c->rows = c->nnz = 0;
}

long long block_rows(long long n, int world) { return (n + world - 1) / world; }

}  // namespace
#include "topn_abi.cuh"
namespace {

void close_peers(als_handle* h) {
  for (int r = 0; r < 16; r++) {
    if (h->peer_X[r]) cudaIpcCloseMemHandle(h->peer_X[r]);
    if (h->peer_Y[r]) cudaIpcCloseMemHandle(h->peer_Y[r]);
    h->peer_X[r] = h->peer_Y[r] = nullptr;
  }
  h->p2p = false;
}

// Map every other rank's X and Y replica into this process (cudaIpc handles exchanged with one
// ncclAllGather).  Collective.  Any failure leaves p2p off: the exchange then falls back to
// ncclAllGather after each kernel.
int setup_peers(als_handle* h) {
  close_peers(h);
  if (!h->comm || h->world < 2 || h->world > 16 || h->p2p_disabled) return ALS_OK;
  struct Pair { cudaIpcMemHandle_t x, y; int ok; int pad[3]; };
  Pair mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = cudaIpcGetMemHandle(&mine.x, h->X) == cudaSuccess && cudaIpcGetMemHandle(&mine.y, h->Y) == cudaSuccess;
  cudaGetLastError();
  Pair* d_all = nullptr;
  CU(h, cudaMalloc(&d_all, sizeof(Pair) * h->world));
  CU(h, cudaMemcpyAsync(d_all + h->rank, &mine, sizeof(Pair), cudaMemcpyHostToDevice, h->stream));
  ncclResult_t r = g_nccl.AllGather(d_all + h->rank, d_all, sizeof(Pair), ncclChar, h->comm, h->stream);
  if (r != ncclSuccess) { cudaFree(d_all); return fail(h, ALS_E_NCCL, "ncclAllGather(ipc handles): %s", g_nccl.GetErrorString(r)); }
  Pair all[16];
  CU(h, cudaMemcpyAsync(all, d_all, sizeof(Pair) * h->world, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_all);
  bool ok = true;
  for (int q = 0; q < h->world; q++) ok = ok && all[q].ok;
  for (int q = 0; ok && q < h->world; q++) {
    if (q == h->rank) continue;
    void *px = nullptr, *py = nullptr;
    if (cudaIpcOpenMemHandle(&px, all[q].x, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&py, all[q].y, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      if (px) cudaIpcCloseMemHandle(px);
      ok = false;
      break;
    }
    h->peer_X[q] = (float*)px;
    h->peer_Y[q] = (float*)py;
  }
  // every rank must agree, or some would wait for rows nobody pushes
  int* d_ok = h->d_flag + 3;
  const int mine_ok = ok ? 1 : 0;
  CU(h, cudaMemcpyAsync(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  r = g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream);
  if (r != ncclSuccess) return fail(h, ALS_E_NCCL, "ncclAllReduce(p2p agreement): %s", g_nccl.GetErrorString(r));
  int all_ok = 0;
  CU(h, cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (!all_ok) close_peers(h);
  else h->p2p = true;
  return ALS_OK;
}

void free_factors(als_handle* h) {
  if (h->factors_ipc) {
    dev_free_ipc(h, &h->X, (size_t)h->users_alloc * h->ks);
    dev_free_ipc(h, &h->Y, (size_t)h->items_alloc * h->ks);
  } else {
    dev_free(h, &h->X, (size_t)h->users_alloc * h->ks);
    dev_free(h, &h->Y, (size_t)h->items_alloc * h->ks);
  }
  // cached views of the old allocations
  h->gather_map_base[0] = h->gather_map_base[1] = nullptr;
}

int alloc_factors(als_handle* h) {
  // Factor replicas are padded to world * block rows so the per-half all-gather is a
  // plain in-place ncclAllGather of equal blocks.
  const long long ua = block_rows(h->n_users, h->world) * h->world;
  const long long ia = block_rows(h->n_items, h->world) * h->world;
  // single-GPU handles take the replicas from the pool like everything else (cudaFree / cudaMalloc of
  // 2.8 GB cost 25-400 ms per model build depending on the box); only handles with a communicator
  // need cudaIpc-mappable (plain) allocations
  const bool want_ipc = h->comm != nullptr;
  if (h->X && ua == h->users_alloc && h->Y && ia == h->items_alloc && h->factors_ipc == want_ipc) return ALS_OK;
  if (h->p2p) {
    // collective: every rank re-allocates for the same new sizes; nobody frees a replica while
    // a peer still has it mapped
    close_peers(h);
    g_nccl.AllReduce(h->d_flag + 3, h->d_flag + 3, 1, ncclInt, ncclMin, h->comm, h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
  }
  free_factors(h);
  h->users_alloc = ua;
  h->items_alloc = ia;
  h->factors_ipc = want_ipc;
  int rc;
  if (want_ipc) {
    if ((rc = dev_alloc_ipc(h, &h->X, (size_t)ua * h->ks)) != ALS_OK) return rc;
    if ((rc = dev_alloc_ipc(h, &h->Y, (size_t)ia * h->ks)) != ALS_OK) return rc;
  } else {
    if ((rc = dev_alloc(h, &h->X, (size_t)ua * h->ks)) != ALS_OK) return rc;
    if ((rc = dev_alloc(h, &h->Y, (size_t)ia * h->ks)) != ALS_OK) return rc;
  }
  CU(h, cudaMemsetAsync(h->X, 0, sizeof(float) * (size_t)ua * h->ks, h->stream));
  CU(h, cudaMemsetAsync(h->Y, 0, sizeof(float) * (size_t)ia * h->ks, h->stream));
  return setup_peers(h);
}

// local block of `n` rows owned by this rank
void local_block(const als_handle* h, long long n, long long* begin, long long* end) {
  const long long b = block_rows(n, h->world);
  *begin = (long long)h->rank * b;
  if (*begin > n) *begin = n;
  *end = *begin + b;
  if (*end > n) *end = n;
}

// ---- profiling helpers -----------------------------------------------------
void prof_begin(als_handle* h, cudaEvent_t* a) {
  *a = nullptr;
  if (!h->profile) return;
  cudaEventCreate(a);
  cudaEventRecord(*a, h->stream);
}
void prof_end(als_handle* h, cudaEvent_t a, int kind) {
  if (!h->profile || !a) return;
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, h->stream);
  if (h->n_pending == h->cap_pending) {
    h->cap_pending = h->cap_pending ? 2 * h->cap_pending : 64;
    h->pending = (als_handle::PendingEv*)realloc(h->pending,
                                                 sizeof(als_handle::PendingEv) * h->cap_pending);
  }
  h->pending[h->n_pending++] = {a, b, kind};
}
void prof_drain(als_handle* h) {
  for (int i = 0; i < h->n_pending; i++) {
    float ms = 0.f;
    cudaEventSynchronize(h->pending[i].b);
    cudaEventElapsedTime(&ms, h->pending[i].a, h->pending[i].b);
    switch (h->pending[i].kind) {
      case 0: h->tm.gramian_ms += ms; break;
      case 1: h->tm.update_x_ms += ms; break;
      case 2: h->tm.update_y_ms += ms; break;
      default: h->tm.exchange_ms += ms; break;
    }
    cudaEventDestroy(h->pending[i].a);
    cudaEventDestroy(h->pending[i].b);
  }
  h->n_pending = 0;
}

// ---- kernels launch --------------------------------------------------------
template <int KS>
int launch_gramian_t(als_handle* h, const float* M, long long n_rows) {
  using S = GramShape<KS>;
  int grid = h->sm_count * 2;
  const long long chunks = (n_rows + kGramChunk - 1) / kGramChunk;
  if (chunks < grid) grid = (int)(chunks > 0 ? chunks : 1);
  const int n_partials = grid * S::kPartialsPerCta;
  if (n_partials > h->g_partials) {
    dev_free(h, &h->G_partial, (size_t)h->g_partials * h->ks * h->ks);
    h->g_partials = h->sm_count * 2 * S::kPartialsPerCta;
    int rc = dev_alloc(h, &h->G_partial, (size_t)h->g_partials * h->ks * h->ks);
    if (rc != ALS_OK) return rc;
  }
  gramian_partial_kernel<KS><<<grid, kGramThreads, 0, h->stream>>>(M, n_rows, h->G_partial);
  const int kk = KS * KS;
  gramian_reduce_kernel<<<(kk + 63) / 64, 64 * kGramReduceSlices, 0, h->stream>>>(h->G_partial, n_partials, KS, h->G);
  h->launches += 2;
  CU(h, cudaGetLastError());
  return ALS_OK;
}

int launch_gramian_local(als_handle* h, const float* M, long long n_rows);

// G = M^T M over all n_rows rows of the replica.  With a communicator every rank reduces only
// its own block of rows (the block it wrote in the previous half) and the k x k fp64 partials
// are summed with one ncclAllReduce (32 KB at k = 64): the serial Gramian pass of the
// reference (MatrixUtils.transposeTimesSelf) scales with the number of GPUs like the row
// updates do.  Every rank receives the same bits.
int launch_gramian(als_handle* h, const float* M, long long n_rows) {
  if (h->world == 1 || !h->comm) return launch_gramian_local(h, M, n_rows);
  const long long b = block_rows(n_rows, h->world);
  const long long lo = b * h->rank;
  long long cnt = n_rows - lo;
  if (cnt > b) cnt = b;
  if (cnt < 0) cnt = 0;
  int rc = launch_gramian_local(h, M + (size_t)lo * h->ks, cnt);
  if (rc != ALS_OK) return rc;
  cudaEvent_t a;
  prof_begin(h, &a);
  ncclResult_t r = g_nccl.AllReduce(h->G, h->G, (size_t)h->ks * h->ks, ncclDouble, ncclSum, h->comm,
                                    h->stream);
  if (r != ncclSuccess) return fail(h, ALS_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
  h->launches += 1;
  prof_end(h, a, 0);
  return ALS_OK;
}

int launch_gramian_local(als_handle* h, const float* M, long long n_rows) {
  cudaEvent_t a;
  prof_begin(h, &a);
  nvtxRangePushA("als:gramian");
  int rc;
  switch (h->ks) {
    case 4: rc = launch_gramian_t<4>(h, M, n_rows); break;
    case 8: rc = launch_gramian_t<8>(h, M, n_rows); break;
    case 16: rc = launch_gramian_t<16>(h, M, n_rows); break;
    case 32: rc = launch_gramian_t<32>(h, M, n_rows); break;
    case 64: rc = launch_gramian_t<64>(h, M, n_rows); break;
    case 128: rc = launch_gramian_t<128>(h, M, n_rows); break;
    default: rc = fail(h, ALS_E_UNSUPPORTED, "padded feature count %d unsupported", h->ks); break;
  }
  nvtxRangePop();
  prof_end(h, a, 0);
  return rc;
}

template <int KS>
int launch_simt_t(als_handle* h, const RowUpdateParams& p) {
  using S = SimtShape<KS>;
  const size_t smem = S::smem_bytes();
  long long grid = (long long)h->sm_count * h->simt_blocks_per_sm;
  if (grid > p.n_rows) grid = p.n_rows > 0 ? p.n_rows : 1;
  row_update_simt_kernel<KS><<<(int)grid, kSimtThreads, smem, h->stream>>>(p);
  h->launches += 1;
  CU(h, cudaGetLastError());
  return ALS_OK;
}

int launch_simt(als_handle* h, const RowUpdateParams& p) {
  switch (h->ks) {
    case 4: return launch_simt_t<4>(h, p);
    case 8: return launch_simt_t<8>(h, p);
    case 16: return launch_simt_t<16>(h, p);
    case 32: return launch_simt_t<32>(h, p);
    case 64: return launch_simt_t<64>(h, p);
    case 128: return launch_simt_t<128>(h, p);
    default: return fail(h, ALS_E_UNSUPPORTED, "padded feature count %d unsupported", h->ks);
  }
}

// cudaFuncSetAttribute / occupancy are per device: done once per handle, on the handle's device
// (als_create), for every kernel instantiation this handle can launch.
template <int KS>
int configure_simt_t(als_handle* h) {
  const size_t smem = SimtShape<KS>::smem_bytes();
  CU(h, cudaFuncSetAttribute(row_update_simt_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem));
  int b = 1;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, row_update_simt_kernel<KS>, kSimtThreads, smem));
  h->simt_blocks_per_sm = b < 1 ? 1 : b;
  return ALS_OK;
}
template <class K>
int set_smem_attr(als_handle* h, K kernel, size_t bytes) {
  CU(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return ALS_OK;
}
int configure_kernels(als_handle* h) {
  int rc;
  switch (h->ks) {
    case 4: rc = configure_simt_t<4>(h); break;
    case 8: rc = configure_simt_t<8>(h); break;
    case 16: rc = configure_simt_t<16>(h); break;
    case 32: rc = configure_simt_t<32>(h); break;
    case 64: rc = configure_simt_t<64>(h); break;
    case 128: rc = configure_simt_t<128>(h); break;
    default: return fail(h, ALS_E_UNSUPPORTED, "padded feature count %d unsupported", h->ks);
  }
  if (rc != ALS_OK || h->kernel != ALS_KERNEL_TCGEN05) return rc;
  if (h->ks == 32) {
    if ((rc = set_smem_attr(h, umma::row_update_umma_kernel<32, 4>, umma::Smem<32, 4>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, umma::row_update_umma_kernel<32, 8>, umma::Smem<32, 8>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<32, v2::Mixes<32>::Long>, v2::Smem<32, v2::Mixes<32>::Long>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<32, v2::Mixes<32>::Long, true>, v2::Smem<32, v2::Mixes<32>::Long>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<32, v2::Mixes<32>::Short>, v2::Smem<32, v2::Mixes<32>::Short>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<32, v2::Mixes<32>::Short, true>, v2::Smem<32, v2::Mixes<32>::Short>::kTotal)) != ALS_OK) return rc;
  } else if (h->ks == 64) {
    if ((rc = set_smem_attr(h, umma::row_update_umma_kernel<64, 4>, umma::Smem<64, 4>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, umma::row_update_umma_kernel<64, 8>, umma::Smem<64, 8>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<64, v2::MixLong>, v2::Smem<64, v2::MixLong>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<64, v2::MixLong, true>, v2::Smem<64, v2::MixLong>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<64, v2::MixShort>, v2::Smem<64, v2::MixShort>::kTotal)) != ALS_OK) return rc;
    if ((rc = set_smem_attr(h, v2::row_update_v2_kernel<64, v2::MixShort, true>, v2::Smem<64, v2::MixShort>::kTotal)) != ALS_OK) return rc;
  }
  return ALS_OK;
}

// rows of R by length, longest first (stable: equal lengths keep their stored order, so uniform
// workloads are visited exactly as before)
int ensure_row_order(als_handle* h, const Csr& R, int which) {
  if (h->order_valid[which] && h->order_rows[which] == R.rows) return ALS_OK;
  if (R.rows >= (1LL << 31)) return fail(h, ALS_E_UNSUPPORTED, "more than 2^31 - 1 local rows");
  dev_free(h, &h->d_order[which], (size_t)h->order_rows[which]);
  h->order_rows[which] = 0;
  h->order_valid[which] = false;
  const size_t n = (size_t)R.rows;
  int rc;
  unsigned *k_in = nullptr, *k_out = nullptr;
  int* v_in = nullptr;
  if ((rc = dev_alloc(h, &h->d_order[which], n)) != ALS_OK) return rc;
  h->order_rows[which] = R.rows;
  if ((rc = dev_alloc(h, &k_in, n)) != ALS_OK || (rc = dev_alloc(h, &k_out, n)) != ALS_OK ||
      (rc = dev_alloc(h, &v_in, n)) != ALS_OK) {
    dev_free(h, &k_in, n); dev_free(h, &k_out, n); dev_free(h, &v_in, n);
    return rc;
  }
  row_length_keys_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(R.ptr, R.rows, k_in, v_in);
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, k_in, k_out, v_in, h->d_order[which],
                                                            (int)n, 0, 32, h->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, h->stream);
  if (e == cudaSuccess)
    e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, k_in, k_out, v_in, h->d_order[which], (int)n, 0, 32,
                                                  h->stream);
  if (tmp) cudaFreeAsync(tmp, h->stream);
  dev_free(h, &k_in, n); dev_free(h, &k_out, n); dev_free(h, &v_in, n);
  h->launches += 2;
  if (e != cudaSuccess) return fail(h, ALS_E_CUDA, "row order sort: %s", cudaGetErrorString(e));
  h->order_valid[which] = true;
  return ALS_OK;
}

// Stash mode for the coming launch of orientation `which` (see RowUpdateParams::stash): decided from the
// number of rows the previous launch of the same half handed to the fp64 paths (known by now: at least
// one other launch has been queued since), with hysteresis; buffers are created on first use.
int prepare_stash(als_handle* h, const Csr& R, int which) {
  if (h->stash_policy >= 0) {
    h->stash_on[which] = h->stash_policy == 1;
  } else if (h->fail_pending[which]) {
    CU(h, cudaEventSynchronize(h->fail_ev[which]));
    h->fail_pending[which] = false;
    const long long fails = (long long)h->h_fail[2 * which] + (long long)h->h_fail[2 * which + 1];
    const long long on_at = R.rows / 64 > 256 ? R.rows / 64 : 256;
    if (!h->stash_on[which] && fails >= on_at) h->stash_on[which] = true;
    else if (h->stash_on[which] && fails < on_at / 4) h->stash_on[which] = false;
  }
  if (!h->stash_on[which]) return ALS_OK;
  int rc;
  const size_t rec = (h->ks == 64 ? (size_t)WPanels<64>::kFloats : (size_t)WPanels<32>::kFloats) + (size_t)h->ks;
  const size_t stash_floats = (size_t)h->sm_count * 16 * rec;  // (at most 16 solving warps per CTA)
  if (!h->d_stash || h->stash_floats < stash_floats) {
    dev_free(h, &h->d_stash, h->stash_floats);
    h->stash_floats = 0;
    if ((rc = dev_alloc(h, &h->d_stash, stash_floats)) != ALS_OK) return rc;
    h->stash_floats = stash_floats;
  }
  if (!h->d_resolve_count && (rc = dev_alloc(h, &h->d_resolve_count, 1)) != ALS_OK) return rc;
  // records for every row, up to 2 GB (rows beyond that take the retry list)
  long long cap = (long long)((2ULL << 30) / (rec * sizeof(float)));
  if (cap > R.rows) cap = R.rows;
  if (cap < 1) cap = 1;
  if (h->resolve_cap < cap || h->resolve_rec != rec) {
    dev_free(h, &h->d_resolve_buf, (size_t)h->resolve_cap * h->resolve_rec);
    dev_free(h, &h->d_resolve_rows, (size_t)h->resolve_cap);
    h->resolve_cap = 0;
    if ((rc = dev_alloc(h, &h->d_resolve_buf, (size_t)cap * rec)) != ALS_OK) return rc;
    if ((rc = dev_alloc(h, &h->d_resolve_rows, (size_t)cap)) != ALS_OK) return rc;
    h->resolve_cap = cap;
    h->resolve_rec = rec;
  }
  return ALS_OK;
}

// Rows longer than h->split_limit entries are walked as several virtual rows (equal chunks) by the
// tensor-core kernel, so that one 10^5..10^6-entry row of power-law data (a popular item's column) is shared by
// many CTAs instead of deciding the kernel's tail; see RowUpdateParams::vrow.  Built once per
// interaction array.  n_acc == 0 afterwards: nothing to split, the orientation is walked as stored.
int ensure_split(als_handle* h, const Csr& R, int which) {
  als_handle::Split& s = h->split[which];
  if (s.valid && s.rows == R.rows) return ALS_OK;
  free_split(h, which);
  s.rows = R.rows;
  s.n_virtual = R.rows;
  if (h->split_limit <= 0 || R.rows == 0 || R.nnz <= h->split_limit) { s.valid = true; return ALS_OK; }
  const size_t n1 = (size_t)R.rows + 1;
  long long *nch = nullptr, *vfirst = nullptr;
  int *is_long = nullptr, *accfirst = nullptr;
  void* tmp = nullptr;
  int rc = ALS_OK;
  auto done = [&](int code) {
    if (tmp) cudaFreeAsync(tmp, h->stream);
    dev_free(h, &nch, n1); dev_free(h, &vfirst, n1); dev_free(h, &is_long, n1); dev_free(h, &accfirst, n1);
    return code;
  };
  if ((rc = dev_alloc(h, &nch, n1)) != ALS_OK || (rc = dev_alloc(h, &vfirst, n1)) != ALS_OK ||
      (rc = dev_alloc(h, &is_long, n1)) != ALS_OK || (rc = dev_alloc(h, &accfirst, n1)) != ALS_OK)
    return done(rc);
  split_counts_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(R.ptr, R.rows, h->split_limit, nch, is_long);
  size_t b1 = 0, b2 = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, b1, nch, vfirst, (int)n1, h->stream);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, b2, is_long, accfirst, (int)n1, h->stream);
  const size_t tb = b1 > b2 ? b1 : b2;
  if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tb ? tb : 1, h->stream);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, b1, nch, vfirst, (int)n1, h->stream);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, b2, is_long, accfirst, (int)n1, h->stream);
  long long n_virtual = 0;
  int n_acc = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_virtual, vfirst + R.rows, sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_acc, accfirst + R.rows, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  h->launches += 3;
  if (e != cudaSuccess) return done(fail(h, ALS_E_CUDA, "row split: %s", cudaGetErrorString(e)));
  if (n_acc == 0) { s.valid = true; return done(ALS_OK); }
  if (n_virtual >= (1LL << 31)) return done(fail(h, ALS_E_UNSUPPORTED, "more than 2^31 - 1 local (virtual) rows"));
  s.n_virtual = n_virtual;
  s.n_acc = n_acc;
  s.acc_floats = (h->ks == 64 ? (size_t)WPanels<64>::kFloats : (size_t)WPanels<32>::kFloats) + (size_t)h->ks;
  if ((rc = dev_alloc(h, &s.vptr, (size_t)n_virtual + 1)) != ALS_OK || (rc = dev_alloc(h, &s.vrow, (size_t)n_virtual)) != ALS_OK ||
      (rc = dev_alloc(h, &s.vacc, (size_t)n_virtual)) != ALS_OK || (rc = dev_alloc(h, &s.acc_chunks, (size_t)n_acc)) != ALS_OK ||
      (rc = dev_alloc(h, &s.gacc, (size_t)n_acc * s.acc_floats)) != ALS_OK || (rc = dev_alloc(h, &s.gcount, (size_t)n_acc)) != ALS_OK) {
    free_split(h, which);
    return done(rc);
  }
  split_fill_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(R.ptr, R.rows, h->split_limit, vfirst, accfirst, s.vptr, s.vrow,
                                                           s.vacc, s.acc_chunks);
  h->launches += 1;
  if ((e = cudaGetLastError()) != cudaSuccess) {
    free_split(h, which);
    return done(fail(h, ALS_E_CUDA, "row split: %s", cudaGetErrorString(e)));
  }
  s.valid = true;
  return done(ALS_OK);
}

int launch_row_update(als_handle* h, const Csr& R, const float* M, float* out, int which) {
  RowUpdateParams p;
  p.row_ptr = R.ptr;
  p.col_idx = R.idx;
  p.val = R.val;
  p.n_rows = R.rows;
  p.row_offset = R.row_begin;
  p.M = M;
  p.G = h->G;
  p.out = out;
  p.k = h->k;
  p.alpha = (float)h->cfg.alpha;
  p.lambda_alpha = h->cfg.lambda * h->cfg.alpha;
  p.reconstruct_r = h->cfg.reconstruct_r;
  p.loss_ignores_unspecified = h->cfg.loss_ignores_unspecified;
  p.threshold = (float)h->cfg.singularity_threshold;
  p.which = which;
  p.status = h->d_status;
  p.ticket = h->d_ticket;
  p.row_list = nullptr;
  p.row_list_count = nullptr;
  p.row_order = nullptr;
  p.vrow = nullptr;
  p.vacc = nullptr;
  p.acc_chunks = nullptr;
  p.real_ptr = R.ptr;
  p.gacc = nullptr;
  p.gcount = nullptr;
  // the second-generation tensor-core kernel walks rows of more than split_limit entries as chunks
  const bool v2_kernel = h->kernel == ALS_KERNEL_TCGEN05 && (h->ks == 64 || h->ks == 32) && !h->legacy_umma;
  const als_handle::Split* sp = nullptr;
  if (v2_kernel) {
    const int src = ensure_split(h, R, which);
    if (src != ALS_OK) return src;
    if (h->split[which].n_acc > 0) sp = &h->split[which];
  }
  Csr Rv = R;  // the rows the kernel walks (virtual rows when split)
  if (sp) {
    Rv.ptr = sp->vptr;
    Rv.rows = sp->n_virtual;
  }
  if (!h->no_row_order && R.rows > 0) {
    const int orc = ensure_row_order(h, Rv, which);
    if (orc != ALS_OK) return orc;
    p.row_order = h->d_order[which];
  }
  p.solve_empty = 0;
  p.retry_rows = nullptr;
  p.retry_count = nullptr;
  p.n_peers = 0;
  if (h->p2p) {
    float* const* peers = which == 0 ? h->peer_X : h->peer_Y;
    for (int r = 0; r < h->world; r++)
      if (r != h->rank) p.peer_out[p.n_peers++] = peers[r];
  }
  CU(h, cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned long long), h->stream));
  int rc = ALS_OK;
  if (h->kernel == ALS_KERNEL_TCGEN05) {
    if (h->retry_cap < R.rows) {
      dev_free(h, &h->d_retry_rows, (size_t)h->retry_cap);
      h->retry_cap = R.rows;
      if ((rc = dev_alloc(h, &h->d_retry_rows, (size_t)h->retry_cap)) != ALS_OK) return rc;
    }
    CU(h, cudaMemsetAsync(h->d_retry_count, 0, sizeof(int), h->stream));
  }
  cudaEvent_t a;
  prof_begin(h, &a);
  nvtxRangePushA(which == 0 ? "als:row_update_x" : "als:row_update_y");
  if (h->kernel == ALS_KERNEL_TCGEN05) {
    p.retry_rows = h->d_retry_rows;
    p.retry_count = h->d_retry_count;
    // role mix by average row length unless MYRRIX_ALS_MIX forced one at als_create
    const bool long_rows = h->mix_override ? (h->mix_override == 4)
                                           : (R.rows > 0 && R.nnz / R.rows >= kLongRowEntries);
    const int mi = which == 0 ? 1 : 0;  // the X half gathers rows of Y and vice versa
    if ((h->ks == 64 || h->ks == 32) && !h->legacy_umma) {
      const long long m_rows = which == 0 ? h->n_items : h->n_users;
      if (m_rows >= (long long)v2::kOobRow) {  // (the gathers' "no entry" row index must lie past the tensor)
        nvtxRangePop();
        return fail(h, ALS_E_UNSUPPORTED, "factor matrices of 2^30 rows or more");
      }
      if (h->gather_map_base[mi] != M || h->gather_map_rows[mi] != m_rows) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) { nvtxRangePop(); return fail(h, ALS_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver"); }
        const cuuint64_t gdim[2] = {(cuuint64_t)h->ks, (cuuint64_t)m_rows};
        const cuuint64_t gstr[1] = {(cuuint64_t)h->ks * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)h->ks, 1};  // tile::gather4 takes four one-row boxes
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = enc(&h->gather_map[mi], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)M, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { nvtxRangePop(); return fail(h, ALS_E_CUDA, "cuTensorMapEncodeTiled(factor) failed (%d)", (int)cr); }
        h->gather_map_base[mi] = M;
        h->gather_map_rows[mi] = m_rows;
      }
    }
    const CUtensorMap& tm = h->gather_map[mi];
    p.stash = nullptr;
    p.resolve_buf = nullptr;
    p.resolve_rows = nullptr;
    p.resolve_count = nullptr;
    p.resolve_cap = 0;
    if (v2_kernel) {
      const int src = prepare_stash(h, R, which);
      if (src != ALS_OK) { nvtxRangePop(); return src; }
      if (h->stash_on[which]) {
        p.stash = h->d_stash;
        p.resolve_buf = h->d_resolve_buf;
        p.resolve_rows = h->d_resolve_rows;
        p.resolve_count = h->d_resolve_count;
        p.resolve_cap = (int)h->resolve_cap;
        cudaMemsetAsync(h->d_resolve_count, 0, sizeof(int), h->stream);
      }
    }
    const RowUpdateParams p_rows = p;  // the rows as stored: what the fp64 re-solves below walk
    if (sp) {
      p.row_ptr = sp->vptr;
      p.n_rows = sp->n_virtual;
      p.vrow = sp->vrow;
      p.vacc = sp->vacc;
      p.acc_chunks = sp->acc_chunks;
      p.gacc = sp->gacc;
      p.gcount = sp->gcount;
      cudaMemsetAsync(sp->gacc, 0, (size_t)sp->n_acc * sp->acc_floats * sizeof(float), h->stream);
      cudaMemsetAsync(sp->gcount, 0, (size_t)sp->n_acc * sizeof(int), h->stream);
    }
    const bool ext = sp != nullptr || p.stash != nullptr;  // virtual rows / stash mode: the extended variant
    if (h->ks == 64 && !h->legacy_umma) {
      if (ext)
        rc = long_rows ? launch_row_update_v2_t<64, v2::MixLong, true>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err))
                       : launch_row_update_v2_t<64, v2::MixShort, true>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err));
      else
        rc = long_rows ? launch_row_update_v2_t<64, v2::MixLong>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err))
                       : launch_row_update_v2_t<64, v2::MixShort>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err));
    } else if (h->ks == 32 && !h->legacy_umma) {
      if (ext)
        rc = long_rows ? launch_row_update_v2_t<32, v2::Mixes<32>::Long, true>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err))
                       : launch_row_update_v2_t<32, v2::Mixes<32>::Short, true>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err));
      else
        rc = long_rows ? launch_row_update_v2_t<32, v2::Mixes<32>::Long>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err))
                       : launch_row_update_v2_t<32, v2::Mixes<32>::Short>(p, tm, h->sm_count, h->stream, h->err, sizeof(h->err));
    } else {
      rc = launch_row_update_umma(h->ks, p, long_rows, h->sm_count, h->stream, h->err, sizeof(h->err));
    }
    if (rc == ALS_OK) {
      h->launches += 1;
      // fp64 re-solve of the rows the fp32 tensor-core path refused (normally none): the
      // CUDA-core kernel in row-list mode; it exits immediately when the list is empty.
      accumulate_count_kernel<<<1, 1, 0, h->stream>>>(h->d_retry_count, h->d_retry_total);
      h->launches += 1;
      p = p_rows;
      if (p.stash) {
        // rows refused in stash mode: fp64 solve from the fp64 Gramian and the stashed data term
        accumulate_resolve_kernel<<<1, 1, 0, h->stream>>>(h->d_resolve_count, p.resolve_cap, h->d_retry_total);
        // (latency-bound: one barrier per column; as many CTAs per SM as the 35 KB / 10 KB of shared memory allow)
        if (h->ks == 64) resolve_fp64_kernel<64><<<h->sm_count * 6, 128, 0, h->stream>>>(p);
        else resolve_fp64_kernel<32><<<h->sm_count * 12, 128, 0, h->stream>>>(p);
        h->launches += 2;
      }
      if (v2_kernel && h->stash_policy < 0) {
        // how many rows left the fp32 path: read before the next launch of this half (prepare_stash)
        if (!h->h_fail) {
          CU(h, cudaHostAlloc((void**)&h->h_fail, 4 * sizeof(int), cudaHostAllocDefault));
          memset(h->h_fail, 0, 4 * sizeof(int));
        }
        if (!h->fail_ev[which]) CU(h, cudaEventCreateWithFlags(&h->fail_ev[which], cudaEventDisableTiming));
        h->h_fail[2 * which + 1] = 0;
        CU(h, cudaMemcpyAsync(&h->h_fail[2 * which], h->d_retry_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        if (p.stash)
          CU(h, cudaMemcpyAsync(&h->h_fail[2 * which + 1], h->d_resolve_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaEventRecord(h->fail_ev[which], h->stream));
        h->fail_pending[which] = true;
      }
      RowUpdateParams q = p;
      q.row_list = h->d_retry_rows;
      q.row_list_count = h->d_retry_count;
      q.retry_rows = nullptr;
      q.retry_count = nullptr;
      rc = launch_simt(h, q);
    }
  } else {
    rc = launch_simt(h, p);
  }
  if (rc == ALS_OK && h->pe_rows[which] > 0) {
    // present-but-empty rows: W_u = G, b_u = 0 (AlternatingLeastSquares.java:391-410 walks every
    // map entry; InputFilesReader.java:202-211 leaves emptied maps in place) -> x_u = 0, or
    // ALS_E_SINGULAR when G itself is singular, through the same fp64 solve as every other row
    RowUpdateParams q = p;
    q.row_list = h->d_pe_rows[which];
    q.row_list_count = h->d_pe_count[which];
    q.solve_empty = 1;
    q.retry_rows = nullptr;
    q.retry_count = nullptr;
    CU(h, cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned long long), h->stream));
    rc = launch_simt(h, q);
  }
  nvtxRangePop();
  prof_end(h, a, which == 0 ? 1 : 2);
  return rc;
}

int exchange(als_handle* h, float* F, long long n_global) {
  if (h->world == 1 || !h->comm) return ALS_OK;
  // p2p: the row-update kernels already stored every finished row into all replicas.  The next
  // half starts with the Gramian all-reduce, which no rank leaves before every rank has
  // finished (and thereby flushed) its row update: no further synchronisation is needed here.
  if (h->p2p) return ALS_OK;
  cudaEvent_t a;
  prof_begin(h, &a);
  const long long b = block_rows(n_global, h->world);
  const size_t count = (size_t)b * h->ks;
  ncclResult_t r = g_nccl.AllGather(F + (size_t)h->rank * count, F, count, ncclFloat, h->comm,
                                    h->stream);
  if (r != ncclSuccess) return fail(h, ALS_E_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
  h->launches += 1;
  prof_end(h, a, 3);
  return ALS_OK;
}

// Build a transposed orientation on the device: stable LSD radix sort of
// (column key, row<<32|value) so each column keeps ascending row order. Source A holds
// rows [A.row_begin, A.row_begin + A.rows); only columns [col_begin, col_begin + n_cols)
// may occur in it; T gets n_cols rows (row_begin = col_begin) whose indices are GLOBAL rows of A.
int build_transpose_from(als_handle* h, const Csr& A, long long col_begin, long long n_cols, Csr* Tp) {
  Csr& T = *Tp;
  free_csr(h, &T);
  T.rows = n_cols;
  T.nnz = A.nnz;
  T.row_begin = col_begin;
  int rc;
  if ((rc = dev_alloc(h, &T.ptr, (size_t)T.rows + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &T.idx, (size_t)T.nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &T.val, (size_t)T.nnz)) != ALS_OK) return rc;
  if (A.nnz == 0) {
    CU(h, cudaMemsetAsync(T.ptr, 0, sizeof(long long) * ((size_t)T.rows + 1), h->stream));
    return ALS_OK;
  }
  if (A.nnz >= (1LL << 31)) return fail(h, ALS_E_UNSUPPORTED, "nnz >= 2^31 per device");
  int *keys_in = nullptr, *keys_out = nullptr;
  unsigned long long *pk_in = nullptr, *pk_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  const size_t n = (size_t)A.nnz;
  auto cleanup = [&]() {
    dev_free(h, &keys_in, n); dev_free(h, &keys_out, n);
    dev_free(h, &pk_in, n); dev_free(h, &pk_out, n);
    if (tmp) { cudaFreeAsync(tmp, h->stream); h->device_bytes -= (long long)tmp_bytes; }
  };
  if ((rc = dev_alloc(h, &keys_in, n)) != ALS_OK || (rc = dev_alloc(h, &keys_out, n)) != ALS_OK ||
      (rc = dev_alloc(h, &pk_in, n)) != ALS_OK || (rc = dev_alloc(h, &pk_out, n)) != ALS_OK) {
    cleanup();
    return rc;
  }
  expand_rows_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(A.ptr, A.rows, A.row_begin,
                                                            (int)col_begin, A.idx, A.val, keys_in,
                                                            pk_in);
  int end_bit = 1;
  while ((1LL << end_bit) < n_cols && end_bit < 31) end_bit++;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, pk_in,
                                                  pk_out, (int)n, 0, end_bit, h->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, h->stream);
  if (e == cudaSuccess) {
    h->device_bytes += (long long)tmp_bytes;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, pk_in, pk_out, (int)n, 0,
                                        end_bit, h->stream);
  }
  if (e != cudaSuccess) {
    cleanup();
    return fail(h, e == cudaErrorMemoryAllocation ? ALS_E_OOM : ALS_E_CUDA, "transpose sort: %s",
                cudaGetErrorString(e));
  }
  build_ptr_unpack_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(keys_out, pk_out, A.nnz, T.rows,
                                                                 T.ptr, T.idx, T.val);
  h->launches += 3;
  e = cudaStreamSynchronize(h->stream);
  cleanup();
  if (e != cudaSuccess) return fail(h, ALS_E_CUDA, "transpose: %s", cudaGetErrorString(e));
  return ALS_OK;
}

// off[r] = first position in the ascending key array with key >= r * block (r = 0 .. world)
__global__ void block_boundaries_kernel(const int* __restrict__ keys, long long n, long long block, int world,
                                        long long* __restrict__ off) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > world) return;
  const long long target = (long long)r * block;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  off[r] = lo;
}
__global__ void shift_keys_kernel(int* keys, long long n, int delta) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) keys[e] -= delta;
}

// Sharded: build this rank's by-item block from every rank's by-user block on the devices.
// Each rank sorts its entries by item (stable: users stay ascending), which makes the entries
// bound for one item block a contiguous run; the runs are exchanged with grouped ncclSend /
// ncclRecv (an all-to-all), and the received runs -- ordered by source rank, i.e. by user range --
// are merged by one more stable sort on the item.  Collective.
int build_by_item_distributed(als_handle* h) {
  const Csr& A = h->by_user;
  const int W = h->world;
  const long long ib_rows = block_rows(h->n_items, W);
  long long ib, ie;
  local_block(h, h->n_items, &ib, &ie);
  if (A.nnz >= (1LL << 31)) return fail(h, ALS_E_UNSUPPORTED, "nnz >= 2^31 per device");
  const size_t n = (size_t)A.nnz;
  int *k_in = nullptr, *k_out = nullptr, *rk = nullptr, *rk2 = nullptr;
  unsigned long long *p_in = nullptr, *p_out = nullptr, *rp = nullptr, *rp2 = nullptr;
  long long *d_off = nullptr, *d_cnt = nullptr;
  void* tmp = nullptr;
  size_t n_recv = 0;
  int rc = ALS_OK;
  auto cleanup = [&]() {
    cudaFree(k_in); cudaFree(k_out); cudaFree(p_in); cudaFree(p_out); cudaFree(rk); cudaFree(rk2);
    cudaFree(rp); cudaFree(rp2); cudaFree(d_off); cudaFree(d_cnt); cudaFree(tmp);
  };
#define DT(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      cleanup();                                                                                  \
      return fail(h, _e == cudaErrorMemoryAllocation ? ALS_E_OOM : ALS_E_CUDA, "%s: %s", #expr,   \
                  cudaGetErrorString(_e));                                                        \
    }                                                                                             \
  } while (0)
#define NC(expr)                                                                                  \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) {                                                                      \
      cleanup();                                                                                  \
      return fail(h, ALS_E_NCCL, "%s: %s", #expr, g_nccl.GetErrorString(_r));                     \
    }                                                                                             \
  } while (0)
  DT(cudaMalloc(&k_in, sizeof(int) * (n ? n : 1)));
  DT(cudaMalloc(&k_out, sizeof(int) * (n ? n : 1)));
  DT(cudaMalloc(&p_in, sizeof(unsigned long long) * (n ? n : 1)));
  DT(cudaMalloc(&p_out, sizeof(unsigned long long) * (n ? n : 1)));
  DT(cudaMalloc(&d_off, sizeof(long long) * (W + 1)));
  DT(cudaMalloc(&d_cnt, sizeof(long long) * (size_t)W * W));
  int item_bits = 1;
  while ((1LL << item_bits) < h->n_items && item_bits < 31) item_bits++;
  size_t tmp_bytes = 0;
  if (n) {
    expand_rows_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(A.ptr, A.rows, A.row_begin, 0, A.idx, A.val, k_in, p_in);
    DT(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, p_in, p_out, (int)n, 0, item_bits, h->stream));
    DT(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    DT(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, p_in, p_out, (int)n, 0, item_bits, h->stream));
    h->launches += 2;
  }
  block_boundaries_kernel<<<1, 64, 0, h->stream>>>(k_out, (long long)n, ib_rows, W, d_off);
  long long off[17], cnt_all[16 * 16];
  DT(cudaMemcpyAsync(off, d_off, sizeof(long long) * (W + 1), cudaMemcpyDeviceToHost, h->stream));
  DT(cudaStreamSynchronize(h->stream));
  long long mine[16];
  for (int r = 0; r < W; r++) mine[r] = off[r + 1] - off[r];
  DT(cudaMemcpyAsync(d_cnt + (size_t)h->rank * W, mine, sizeof(long long) * W, cudaMemcpyHostToDevice, h->stream));
  NC(g_nccl.AllGather(d_cnt + (size_t)h->rank * W, d_cnt, (size_t)W, ncclInt64, h->comm, h->stream));
  DT(cudaMemcpyAsync(cnt_all, d_cnt, sizeof(long long) * (size_t)W * W, cudaMemcpyDeviceToHost, h->stream));
  DT(cudaStreamSynchronize(h->stream));
  long long roff[17];
  roff[0] = 0;
  for (int src = 0; src < W; src++) roff[src + 1] = roff[src] + cnt_all[src * W + h->rank];
  n_recv = (size_t)roff[W];
  if (n_recv >= (1ULL << 31)) { cleanup(); return fail(h, ALS_E_UNSUPPORTED, "nnz >= 2^31 per device"); }
  DT(cudaMalloc(&rk, sizeof(int) * (n_recv ? n_recv : 1)));
  DT(cudaMalloc(&rp, sizeof(unsigned long long) * (n_recv ? n_recv : 1)));
  NC(g_nccl.GroupStart());
  for (int r = 0; r < W; r++) {
    if (mine[r] > 0) {
      NC(g_nccl.Send(k_out + off[r], (size_t)mine[r], ncclInt32, r, h->comm, h->stream));
      NC(g_nccl.Send(p_out + off[r], (size_t)mine[r], ncclUint64, r, h->comm, h->stream));
    }
    const long long c = cnt_all[r * W + h->rank];
    if (c > 0) {
      NC(g_nccl.Recv(rk + roff[r], (size_t)c, ncclInt32, r, h->comm, h->stream));
      NC(g_nccl.Recv(rp + roff[r], (size_t)c, ncclUint64, r, h->comm, h->stream));
    }
  }
  NC(g_nccl.GroupEnd());
  h->launches += 1;
  // merge the runs: keys local to my item block, stable sort, pointers + unpack
  Csr& T = h->by_item;
  free_csr(h, &T);
  T.rows = ie - ib;
  T.nnz = (long long)n_recv;
  T.row_begin = ib;
  if ((rc = dev_alloc(h, &T.ptr, (size_t)T.rows + 1)) != ALS_OK || (rc = dev_alloc(h, &T.idx, n_recv)) != ALS_OK ||
      (rc = dev_alloc(h, &T.val, n_recv)) != ALS_OK) {
    cleanup();
    return rc;
  }
  if (n_recv == 0) {
    DT(cudaMemsetAsync(T.ptr, 0, sizeof(long long) * ((size_t)T.rows + 1), h->stream));
  } else {
    DT(cudaMalloc(&rk2, sizeof(int) * n_recv));
    DT(cudaMalloc(&rp2, sizeof(unsigned long long) * n_recv));
    shift_keys_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(rk, (long long)n_recv, (int)ib);
    int bits = 1;
    while ((1LL << bits) < T.rows && bits < 31) bits++;
    size_t tb2 = 0;
    DT(cub::DeviceRadixSort::SortPairs(nullptr, tb2, rk, rk2, rp, rp2, (int)n_recv, 0, bits, h->stream));
    if (tb2 > tmp_bytes) { cudaFree(tmp); tmp = nullptr; DT(cudaMalloc(&tmp, tb2)); tmp_bytes = tb2; }
    DT(cub::DeviceRadixSort::SortPairs(tmp, tb2, rk, rk2, rp, rp2, (int)n_recv, 0, bits, h->stream));
    build_ptr_unpack_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(rk2, rp2, T.nnz, T.rows, T.ptr, T.idx, T.val);
    h->launches += 3;
  }
  DT(cudaStreamSynchronize(h->stream));
  cleanup();
#undef DT
#undef NC
  h->have_by_item = true;
  return ALS_OK;
}

int build_transpose(als_handle* h) {
  int rc = build_transpose_from(h, h->by_user, 0, h->n_items, &h->by_item);
  if (rc == ALS_OK) h->have_by_item = true;
  return rc;
}

int upload_csr(als_handle* h, Csr* c, long long rows, long long row_begin, const long long* ptr,
               const int* idx, const float* val, cudaMemcpyKind kind) {
  free_csr(h, c);
  long long first = 0, last = 0;
  if (kind == cudaMemcpyHostToDevice) {
    first = ptr[0];
    last = ptr[rows];
  } else {
    CU(h, cudaMemcpy(&first, ptr, sizeof(long long), cudaMemcpyDeviceToHost));
    CU(h, cudaMemcpy(&last, ptr + rows, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (first != 0 || last < 0) return fail(h, ALS_E_ARG, "row_ptr must start at 0 and be non-negative");
  if (last > 0 && (!idx || !val)) return fail(h, ALS_E_ARG, "null col_idx/val with nnz > 0");
  c->rows = rows;
  c->nnz = last;
  c->row_begin = row_begin;
  int rc;
  if ((rc = dev_alloc(h, &c->ptr, (size_t)rows + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &c->idx, (size_t)c->nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &c->val, (size_t)c->nnz)) != ALS_OK) return rc;
  CU(h, cudaMemcpyAsync(c->ptr, ptr, sizeof(long long) * ((size_t)rows + 1), kind, h->stream));
  CU(h, cudaMemcpyAsync(c->idx, idx, sizeof(int) * (size_t)c->nnz, kind, h->stream));
  CU(h, cudaMemcpyAsync(c->val, val, sizeof(float) * (size_t)c->nnz, kind, h->stream));
  return ALS_OK;
}

// Structural check of an uploaded CSR before any kernel dereferences it: row pointers must be
// non-decreasing and end at nnz, indices must lie in [0, n_cols).  (A bad index would otherwise
// become an out-of-bounds gather and poison the CUDA context.)
__global__ void validate_csr_kernel(const long long* __restrict__ ptr, long long rows, long long nnz,
                                    const int* __restrict__ idx, long long n_cols, int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  for (long long r = t0; r < rows; r += stride) bad |= ptr[r] > ptr[r + 1] || ptr[r] < 0;
  if (t0 == 0) bad |= ptr[rows] != nnz;
  for (long long e = t0; e < nnz; e += stride) bad |= idx[e] < 0 || (long long)idx[e] >= n_cols;
  if (bad) atomicOr(flag, 1);
}
int validate_csr(als_handle* h, const Csr& c, long long n_cols, const char* what) {
  CU(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  validate_csr_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(c.ptr, c.rows, c.nnz, c.idx, n_cols, h->d_flag);
  h->launches += 1;
  int flag = 0;
  CU(h, cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (flag) return fail(h, ALS_E_ARG, "%s: row pointers not monotone or an index outside [0, %lld)", what, n_cols);
  return ALS_OK;
}

// new interactions: forget deferred errors and the present-but-empty lists of the old ones
int reset_for_new_interactions(als_handle* h) {
  h->sticky = ALS_OK;
  h->singular_rank = 0;
  CU(h, cudaMemsetAsync(h->d_status, 0, sizeof(DeviceStatus), h->stream));
  for (int w = 0; w < 2; w++) {
    h->pe_rows[w] = 0;
    CU(h, cudaMemsetAsync(h->d_pe_count[w], 0, sizeof(int), h->stream));
    h->stash_on[w] = false;  // (what the old interactions' launches told about the fp32 path's yield)
    h->fail_pending[w] = false;
  }
  return ALS_OK;
}

int check_ready(als_handle* h) {
  if (!h) return ALS_E_ARG;
  if (!h->by_user.ptr || !h->have_by_item || !h->X || !h->Y)
    return fail(h, ALS_E_STATE, "interactions not set");
  return ALS_OK;
}

}  // namespace

template <int KS>
__global__ void debug_solve_blocked_kernel(const float* __restrict__ W, const float* __restrict__ b, int k,
                                           float threshold, float* __restrict__ x, int* __restrict__ ok) {
  using WP = WPanels<KS>;
  using CB = CholBlocked<KS>;
  __shared__ __align__(16) float slot[WP::kFloats];
  __shared__ __align__(16) float scratch[CB::kScratch];
  const int lane = threadIdx.x;
  for (int e = lane; e < WP::kFloats; e += 32) slot[e] = 0.f;
  __syncwarp();
  for (int e = lane; e < KS * KS; e += 32) {
    const int i = e / KS, j = e % KS;
    if (i >= j) {
      float v = 0.f;
      if (i < k) v = W[i * k + j];
      else if (i == j) v = 1.f;  // padding rows: unit diagonal
      slot[WP::at(i, j)] = -v;
    }
  }
  __syncwarp();
  float bv[CB::kS];
#pragma unroll
  for (int s = 0; s < CB::kS; s++) bv[s] = (lane + 32 * s < k) ? b[lane + 32 * s] : 0.f;
  const long long t0 = clock64();
  const float dmax = CB::diag_max(slot, lane, k);
  const bool good = CB::factor_solve(slot, scratch, bv, dmax, threshold, v2::kCondLimit, lane, k);
  const long long t1 = clock64();
#pragma unroll
  for (int s = 0; s < CB::kS; s++)
    if (lane + 32 * s < k) x[lane + 32 * s] = bv[s];
  if (lane == 0) {
    ok[0] = good ? 1 : 0;
    ok[1] = (int)(t1 - t0);  // cycles of one solve on an otherwise idle SM
  }
}

// ===========================================================================
extern "C" {

int als_abi_version(void) { return MYRRIX_ALS_ABI_VERSION; }

int als_config_default(als_config* cfg) {
  if (!cfg) return ALS_E_ARG;
  memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(als_config);
  cfg->features = 30;       // MatrixFactorizer.DEFAULT_FEATURES (MatrixFactorizer.java:34)
  cfg->alpha = 1.0;         // AlternatingLeastSquares.DEFAULT_ALPHA (:71)
  cfg->lambda = 0.1;        // DEFAULT_LAMBDA (:73)
  cfg->singularity_threshold = 1.0e-5;  // LinearSystemSolver.java:33-34
  cfg->device = 0;
  cfg->kernel = ALS_KERNEL_AUTO;
  return ALS_OK;
}

int als_create(const als_config* cfg, als_handle** out) {
  if (!cfg || !out) return ALS_E_ARG;
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(als_config)) return ALS_E_ARG;
  // Preconditions of the reference constructor (AlternatingLeastSquares.java:137-141).
  if (cfg->features <= 0 || cfg->features > kMaxFeatures) return ALS_E_ARG;
  if (!(cfg->alpha == cfg->alpha) || !(cfg->lambda == cfg->lambda)) return ALS_E_ARG;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return ALS_E_CUDA;  // no CPU fallback: without a GPU this library does not run
  }
  if (cfg->device < 0 || cfg->device >= n_dev) return ALS_E_ARG;
  als_handle* h = new (std::nothrow) als_handle();
  if (!h) return ALS_E_OOM;
  h->cfg = *cfg;
  h->err[0] = 0;
  memset(&h->tm, 0, sizeof(h->tm));
  h->tm.struct_size = (int32_t)sizeof(als_timings);
  h->k = cfg->features;
  h->ks = padded_features(cfg->features);
  h->device = cfg->device;
  *out = h;
  CU(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CU(h, cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    return fail(h, ALS_E_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only",
                prop.major, prop.minor);
  }
  // kernel selection
  const bool umma_ok = umma_supported(h->ks) && !cfg->loss_ignores_unspecified;
  if (cfg->kernel == ALS_KERNEL_TCGEN05 && !umma_ok)
    return fail(h, ALS_E_UNSUPPORTED, "tcgen05 kernel not available for features=%d", cfg->features);
  h->kernel = (cfg->kernel == ALS_KERNEL_SIMT) ? ALS_KERNEL_SIMT
              : (umma_ok ? ALS_KERNEL_TCGEN05 : ALS_KERNEL_SIMT);
  // development switches, read once per handle (never on the launch path)
  if (const char* e = getenv("MYRRIX_ALS_MIX")) {
    const int v = atoi(e);
    if (v == 4 || v == 8) h->mix_override = v;
  }
  if (const char* e = getenv("MYRRIX_ALS_V1")) h->legacy_umma = atoi(e) != 0;
  if (const char* e = getenv("MYRRIX_ALS_NO_P2P")) h->p2p_disabled = atoi(e) != 0;
  if (const char* e = getenv("MYRRIX_ALS_NO_ROW_ORDER")) h->no_row_order = atoi(e) != 0;
  if (const char* e = getenv("MYRRIX_ALS_STASH")) h->stash_policy = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("MYRRIX_ALS_SPLIT_ROWS")) {
    const long long v = atoll(e);
    h->split_limit = v <= 0 ? 0 : (v < 64 ? 64 : v);
  }
  CU(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  {
    // keep freed working storage in the device's pool (see dev_alloc)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  int rc;
  if ((rc = configure_kernels(h)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_flag, 4)) != ALS_OK) return rc;
  for (int w = 0; w < 2; w++) {
    if ((rc = dev_alloc(h, &h->d_pe_count[w], 1)) != ALS_OK) return rc;
    CU(h, cudaMemsetAsync(h->d_pe_count[w], 0, sizeof(int), h->stream));
  }
  if ((rc = dev_alloc(h, &h->G, (size_t)h->ks * h->ks)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_status, 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_ticket, 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_rank, 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_retry_count, 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &h->d_retry_total, 2)) != ALS_OK) return rc;  // [0] all fp64 rows, [1] of those: re-solved from the stash
  CU(h, cudaMemsetAsync(h->d_retry_total, 0, 2 * sizeof(long long), h->stream));
  const size_t scratch = (size_t)2 * h->k * h->k + 2 * h->k + 1;
  if ((rc = dev_alloc(h, &h->d_scratch, scratch > 10000 ? scratch : 10000)) != ALS_OK) return rc;
  CU(h, cudaMemsetAsync(h->d_status, 0, sizeof(DeviceStatus), h->stream));
  CU(h, cudaMemsetAsync(h->G, 0, sizeof(double) * h->ks * h->ks, h->stream));
  return ALS_OK;
}

int als_destroy(als_handle* h) {
  if (!h) return ALS_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  prof_drain(h);
  free(h->pending);
  if (h->p2p) {
    // nobody frees a replica while a peer still has it mapped
    close_peers(h);
    if (h->comm && h->d_flag) {
      g_nccl.AllReduce(h->d_flag + 3, h->d_flag + 3, 1, ncclInt, ncclMin, h->comm, h->stream);
      cudaStreamSynchronize(h->stream);
    }
  }
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  topn_free(h);
  for (int w = 0; w < 2; w++) { cudaFree(h->fi_qrt[w]); cudaFree(h->fi_rdiag[w]); cudaFree(h->fi_perm[w]); }
  for (int w = 0; w < 2; w++) dev_free(h, &h->d_order[w], (size_t)h->order_rows[w]);
  for (int w = 0; w < 2; w++) free_split(h, w);
  free_csr(h, &h->by_user);
  free_csr(h, &h->by_item);
  free_factors(h);
  // pool allocations go back to the pool (stream-ordered; nothing is unmapped)
  void* pooled[] = {h->G, h->G_partial, h->d_status, h->d_ticket, h->d_rank, h->d_scratch, h->d_retry_rows,
                    h->d_retry_count, h->d_retry_total, h->d_flag, h->d_probe_idx, h->d_probe_out,
                    h->d_pe_rows[0], h->d_pe_rows[1], h->d_pe_count[0], h->d_pe_count[1],
                    h->d_stash, h->d_resolve_buf, h->d_resolve_rows, h->d_resolve_count};
  for (void* q : pooled)
    if (q) cudaFreeAsync(q, h->stream);
  cudaStreamSynchronize(h->stream);
  for (int w = 0; w < 2; w++)
    if (h->fail_ev[w]) cudaEventDestroy(h->fail_ev[w]);
  if (h->h_fail) cudaFreeHost(h->h_fail);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return ALS_OK;
}

int als_set_stream(als_handle* h, void* cuda_stream) {
  if (!h) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return ALS_OK;
}

static int set_dims(als_handle* h, int64_t n_users, int64_t n_items) {
  if (n_users <= 0 || n_items <= 0 || n_users >= (1LL << 31) || n_items >= (1LL << 31))
    return fail(h, ALS_E_ARG, "n_users/n_items must be in [1, 2^31)");
  h->n_users = n_users;
  h->n_items = n_items;
  return alloc_factors(h);
}

int als_set_interactions(als_handle* h, int64_t n_users, int64_t n_items, const int64_t* row_ptr,
                         const int32_t* col_idx, const float* val) {
  if (!h || !row_ptr) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  int rc = set_dims(h, n_users, n_items);
  if (rc != ALS_OK) return rc;
  long long ub, ue;
  local_block(h, h->n_users, &ub, &ue);
  h->have_by_item = false;
  if ((rc = reset_for_new_interactions(h)) != ALS_OK) return rc;
  rc = upload_csr(h, &h->by_user, ue - ub, ub, (const long long*)row_ptr, col_idx, val,
                  cudaMemcpyHostToDevice);
  if (rc != ALS_OK) return rc;
  if ((rc = validate_csr(h, h->by_user, h->n_items, "als_set_interactions")) != ALS_OK) {
    free_csr(h, &h->by_user);
    return rc;
  }
  if (h->world == 1) return build_transpose(h);
  // sharded with a communicator: the by-item blocks are built from the by-user blocks on the
  // devices (collective; a later als_set_interactions_by_column overrides it)
  if (h->comm) return build_by_item_distributed(h);
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;  // partition-only handle: caller must follow with als_set_interactions_by_column
}

int als_set_interactions_by_column(als_handle* h, const int64_t* col_ptr, const int32_t* row_idx,
                                   const float* val) {
  if (!h || !col_ptr) return ALS_E_ARG;
  if (!h->by_user.ptr) return fail(h, ALS_E_STATE, "call als_set_interactions first");
  CU(h, cudaSetDevice(h->device));
  long long ib, ie;
  local_block(h, h->n_items, &ib, &ie);
  int rc = upload_csr(h, &h->by_item, ie - ib, ib, (const long long*)col_ptr, row_idx, val,
                      cudaMemcpyHostToDevice);
  if (rc != ALS_OK) return rc;
  if ((rc = validate_csr(h, h->by_item, h->n_users, "als_set_interactions_by_column")) != ALS_OK) {
    free_csr(h, &h->by_item);
    return rc;
  }
  h->have_by_item = true;
  return ALS_OK;
}

int als_set_interactions_device(als_handle* h, int64_t n_users, int64_t n_items,
                                const int64_t* d_row_ptr, const int32_t* d_col_idx,
                                const float* d_val) {
  if (!h || !d_row_ptr) return ALS_E_ARG;
  if (h->world != 1) return fail(h, ALS_E_UNSUPPORTED, "device upload is single-GPU only");
  CU(h, cudaSetDevice(h->device));
  int rc = set_dims(h, n_users, n_items);
  if (rc != ALS_OK) return rc;
  h->have_by_item = false;
  if ((rc = reset_for_new_interactions(h)) != ALS_OK) return rc;
  rc = upload_csr(h, &h->by_user, n_users, 0, (const long long*)d_row_ptr, d_col_idx, d_val,
                  cudaMemcpyDeviceToDevice);
  if (rc != ALS_OK) return rc;
  if ((rc = validate_csr(h, h->by_user, h->n_items, "als_set_interactions_device")) != ALS_OK) {
    free_csr(h, &h->by_user);
    return rc;
  }
  return build_transpose(h);
}

static int set_factor(als_handle* h, float* dst, long long rows, const float* src) {
  if (!h || !src) return ALS_E_ARG;
  if (!dst) return fail(h, ALS_E_STATE, "interactions not set");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpy2DAsync(dst, sizeof(float) * h->ks, src, sizeof(float) * h->k,
                          sizeof(float) * h->k, (size_t)rows, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}
int als_set_y(als_handle* h, const float* y) { return set_factor(h, h ? h->Y : nullptr, h ? h->n_items : 0, y); }
int als_set_x(als_handle* h, const float* x) { return set_factor(h, h ? h->X : nullptr, h ? h->n_users : 0, x); }

static int get_factor(als_handle* h, const float* src, long long rows, float* out) {
  if (!h || !out) return ALS_E_ARG;
  if (!src) return fail(h, ALS_E_STATE, "interactions not set");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpy2DAsync(out, sizeof(float) * h->k, src, sizeof(float) * h->ks,
                          sizeof(float) * h->k, (size_t)rows, cudaMemcpyDeviceToHost, h->stream));
  return als_sync(h);
}
int als_get_x(als_handle* h, float* out) { return get_factor(h, h ? h->X : nullptr, h ? h->n_users : 0, out); }
int als_get_y(als_handle* h, float* out) { return get_factor(h, h ? h->Y : nullptr, h ? h->n_items : 0, out); }

int als_get_factor_block(als_handle* h, int32_t which, int64_t first_row, int64_t n_rows, float* out) {
  if (!h || (which != 0 && which != 1) || first_row < 0 || n_rows < 0 || (n_rows > 0 && !out)) return ALS_E_ARG;
  const float* F = which == 0 ? h->X : h->Y;
  const long long limit = which == 0 ? h->n_users : h->n_items;
  if (!F) return fail(h, ALS_E_STATE, "interactions not set");
  if (first_row + n_rows > limit) return fail(h, ALS_E_ARG, "row block out of range");
  if (n_rows == 0) return ALS_OK;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpy2DAsync(out, sizeof(float) * h->k, F + (size_t)first_row * h->ks, sizeof(float) * h->ks,
                          sizeof(float) * h->k, (size_t)n_rows, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}

int als_get_rows(als_handle* h, int32_t which, const int32_t* rows, int32_t n, float* out) {
  if (!h || (which != 0 && which != 1) || n < 0 || (n > 0 && (!rows || !out))) return ALS_E_ARG;
  const float* F = which == 0 ? h->X : h->Y;
  const long long limit = which == 0 ? h->n_users : h->n_items;
  if (!F) return fail(h, ALS_E_STATE, "interactions not set");
  if (n == 0) return ALS_OK;
  for (int i = 0; i < n; i++)
    if (rows[i] < 0 || rows[i] >= limit) return fail(h, ALS_E_ARG, "row %d out of range", rows[i]);
  CU(h, cudaSetDevice(h->device));
  int* d_rows = nullptr;
  float* d_out = nullptr;
  CU(h, cudaMalloc(&d_rows, sizeof(int) * (size_t)n));
  if (cudaMalloc(&d_out, sizeof(float) * (size_t)n * h->k) != cudaSuccess) { cudaFree(d_rows); return fail(h, ALS_E_OOM, "als_get_rows scratch"); }
  cudaMemcpyAsync(d_rows, rows, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  gather_rows_kernel<<<(int)(((long long)n * h->k + 255) / 256), 256, 0, h->stream>>>(F, h->ks, h->k, d_rows, n, d_out);
  h->launches += 1;
  cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)n * h->k, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_rows); cudaFree(d_out);
  if (e != cudaSuccess) return fail(h, ALS_E_CUDA, "als_get_rows: %s", cudaGetErrorString(e));
  return ALS_OK;
}

int als_half_x(als_handle* h) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  CU(h, cudaSetDevice(h->device));
  // RealMatrix YTY = MatrixUtils.transposeTimesSelf(Y)  (ALS.java:342)
  if ((rc = launch_gramian(h, h->Y, h->n_items)) != ALS_OK) return rc;
  // addWorkers(RbyRow, Y, YTY, X, ...)  (ALS.java:344)
  if ((rc = launch_row_update(h, h->by_user, h->Y, h->X, 0)) != ALS_OK) return rc;
  h->tm.n_half_x++;
  return exchange(h, h->X, h->n_users);
}

int als_half_y(als_handle* h) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  CU(h, cudaSetDevice(h->device));
  if ((rc = launch_gramian(h, h->X, h->n_users)) != ALS_OK) return rc;   // ALS.java:369
  if ((rc = launch_row_update(h, h->by_item, h->X, h->Y, 1)) != ALS_OK) return rc;  // :371
  h->tm.n_half_y++;
  return exchange(h, h->Y, h->n_items);
}

int als_iterate(als_handle* h, int32_t n_iterations) {
  if (!h || n_iterations < 0) return ALS_E_ARG;
  for (int it = 0; it < n_iterations; it++) {
    int rc = als_half_x(h);
    if (rc != ALS_OK) return rc;
    rc = als_half_y(h);
    if (rc != ALS_OK) return rc;
  }
  return ALS_OK;
}

int als_set_present_empty_rows(als_handle* h, int32_t which, const int32_t* rows, int64_t n) {
  if (!h || (which != 0 && which != 1) || n < 0 || (n > 0 && !rows)) return ALS_E_ARG;
  const Csr& c = which == 0 ? h->by_user : h->by_item;
  if (!c.ptr || (which == 1 && !h->have_by_item)) return fail(h, ALS_E_STATE, "interactions not set");
  CU(h, cudaSetDevice(h->device));
  // keep the rows of this rank's block, as local indices; they must have no entries
  int* local = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!local) return ALS_E_OOM;
  long long m = 0;
  const long long n_global = which == 0 ? h->n_users : h->n_items;
  for (int64_t i = 0; i < n; i++) {
    const long long r = rows[i];
    if (r < 0 || r >= n_global) { free(local); return fail(h, ALS_E_ARG, "present-empty row %lld out of range", r); }
    if (r >= c.row_begin && r < c.row_begin + c.rows) local[m++] = (int)(r - c.row_begin);
  }
  dev_free(h, &h->d_pe_rows[which], (size_t)h->pe_rows[which]);
  h->pe_rows[which] = 0;
  int rc = ALS_OK;
  if (m > 0) {
    if ((rc = dev_alloc(h, &h->d_pe_rows[which], (size_t)m)) != ALS_OK) { free(local); return rc; }
    cudaMemcpyAsync(h->d_pe_rows[which], local, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, h->stream);
  }
  const int mi = (int)m;
  cudaMemcpyAsync(h->d_pe_count[which], &mi, sizeof(int), cudaMemcpyHostToDevice, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  free(local);
  if (e != cudaSuccess) return fail(h, ALS_E_CUDA, "present-empty upload: %s", cudaGetErrorString(e));
  h->pe_rows[which] = m;
  return ALS_OK;
}

int als_probe(als_handle* h, const int32_t* users, int32_t n_users, const int32_t* items,
              int32_t n_items, double* out) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  if (!users || !items || !out || n_users < 0 || n_items < 0) return ALS_E_ARG;
  const long long n = (long long)n_users * n_items;
  if (n == 0) return ALS_OK;
  if (n > 1000000) return fail(h, ALS_E_ARG, "probe too large");
  for (int i = 0; i < n_users; i++)
    if (users[i] < 0 || users[i] >= h->n_users) return fail(h, ALS_E_ARG, "probe user out of range");
  for (int j = 0; j < n_items; j++)
    if (items[j] < 0 || items[j] >= h->n_items) return fail(h, ALS_E_ARG, "probe item out of range");
  CU(h, cudaSetDevice(h->device));
  // scratch lives with the handle: the stop rule probes once per iteration
  if (h->probe_idx_cap < (long long)n_users + n_items) {
    dev_free(h, &h->d_probe_idx, (size_t)h->probe_idx_cap);
    h->probe_idx_cap = (long long)n_users + n_items;
    if ((rc = dev_alloc(h, &h->d_probe_idx, (size_t)h->probe_idx_cap)) != ALS_OK) { h->probe_idx_cap = 0; return rc; }
  }
  if (h->probe_out_cap < n) {
    dev_free(h, &h->d_probe_out, (size_t)h->probe_out_cap);
    h->probe_out_cap = n;
    if ((rc = dev_alloc(h, &h->d_probe_out, (size_t)n)) != ALS_OK) { h->probe_out_cap = 0; return rc; }
  }
  int* d_u = h->d_probe_idx;
  int* d_i = h->d_probe_idx + n_users;
  CU(h, cudaMemcpyAsync(d_u, users, sizeof(int) * n_users, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(d_i, items, sizeof(int) * n_items, cudaMemcpyHostToDevice, h->stream));
  probe_kernel<<<(int)((n + 127) / 128), 128, 0, h->stream>>>(h->X, h->Y, h->ks, h->k, d_u, n_users,
                                                             d_i, n_items, h->d_probe_out);
  h->launches += 1;
  CU(h, cudaMemcpyAsync(out, h->d_probe_out, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}

// AlternatingLeastSquares.call's loop (ALS.java:206-257) with the stop rule on the device.
int als_call(als_handle* h, const int32_t* test_users, int32_t n_test_users, const int32_t* test_items,
             int32_t n_test_items, int32_t max_iterations, double convergence_threshold, int32_t random_y,
             int32_t x_is_empty, int32_t* iterations_run, double* last_convergence_value) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  if (n_test_users < 0 || n_test_items < 0 || (n_test_users > 0 && !test_users) || (n_test_items > 0 && !test_items))
    return ALS_E_ARG;
  if (!(convergence_threshold > 0.0 && convergence_threshold < 1.0))  // ALS.java:140
    return fail(h, ALS_E_ARG, "convergence threshold must be in (0,1)");
  const long long n = (long long)n_test_users * n_test_items;
  if (n > 1000000) return fail(h, ALS_E_ARG, "too many convergence test pairs");
  if (max_iterations <= 0 && n == 0) return fail(h, ALS_E_ARG, "no iteration limit and nothing to test convergence on");
  if (iterations_run) *iterations_run = 0;
  if (last_convergence_value) *last_convergence_value = NAN;
  CU(h, cudaSetDevice(h->device));
  double* d_est = nullptr;   // estimates[][] (:215), then the statistic
  double* fresh = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!fresh) return ALS_E_OOM;
  if (cudaMalloc(&d_est, sizeof(double) * (size_t)(n + 2)) != cudaSuccess) { free(fresh); cudaGetLastError(); return fail(h, ALS_E_OOM, "stop-rule scratch"); }
  auto done = [&](int code) { cudaFree(d_est); free(fresh); return code; };
  cudaMemsetAsync(d_est, 0, sizeof(double) * (size_t)(n + 2), h->stream);
  if (n > 0 && !x_is_empty) {
    // estimates of the model as it stands (:216-222); als_probe leaves them in d_probe_out
    if ((rc = als_probe(h, test_users, n_test_users, test_items, n_test_items, fresh)) != ALS_OK) return done(rc);
    cudaMemcpyAsync(d_est, h->d_probe_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream);
  }
  int it = 0;
  double value = NAN;
  while (true) {
    if ((rc = als_half_x(h)) != ALS_OK) return done(rc);   // :230
    if ((rc = als_half_y(h)) != ALS_OK) return done(rc);   // :231
    double stat[2] = {NAN, 0.0};
    if (n > 0) {
      if ((rc = als_probe(h, test_users, n_test_users, test_items, n_test_items, fresh)) != ALS_OK) return done(rc);
      stop_rule_kernel<<<1, 32, 0, h->stream>>>(h->d_probe_out, d_est, (int)n, d_est + n);
      h->launches += 1;
      if (cudaMemcpyAsync(stat, d_est + n, sizeof(stat), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
          cudaStreamSynchronize(h->stream) != cudaSuccess)
        return done(fail(h, ALS_E_CUDA, "stop rule: %s", cudaGetErrorString(cudaGetLastError())));
    }
    // a failed row (singular / non-finite) must surface now, not after more iterations
    if ((rc = als_sync(h)) != ALS_OK) return done(rc);
    it++;
    value = stat[0];
    if (iterations_run) *iterations_run = it;
    if (last_convergence_value) *last_convergence_value = value;
    if (max_iterations > 0 && it >= max_iterations) break;              // :242-245
    if (!isfinite(value)) break;                                        // :248-251
    if (!(random_y && it == 1) && value < convergence_threshold) break; // :253-256
  }
  return done(ALS_OK);
}

// ---- fold-in on the resident rows (foldin_dev.cuh) ---------------------------------------------
int als_set_fold_in_state(als_handle* h, int32_t which, const double* qrt, const double* rdiag,
                          const int32_t* perm, double learn_rate) {
  if (!h || (which != 0 && which != 1)) return ALS_E_ARG;
  if (qrt && (!rdiag || !perm)) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  cudaFree(h->fi_qrt[which]); cudaFree(h->fi_rdiag[which]); cudaFree(h->fi_perm[which]);
  h->fi_qrt[which] = nullptr; h->fi_rdiag[which] = nullptr; h->fi_perm[which] = nullptr;
  h->fi_learn_rate = learn_rate;
  if (!qrt) return ALS_OK;  // "model.solver.xtx.compute=false": no solver for this side
  const int k = h->k;
  for (int i = 0; i < k; i++) {
    if (perm[i] < 0 || perm[i] >= k) return fail(h, ALS_E_ARG, "fold-in state: bad permutation");
    if (!(rdiag[i] == rdiag[i]) || rdiag[i] == 0.0) return fail(h, ALS_E_ARG, "fold-in state: singular R");
  }
  CU(h, cudaMalloc(&h->fi_qrt[which], sizeof(double) * (size_t)k * k));
  CU(h, cudaMalloc(&h->fi_rdiag[which], sizeof(double) * (size_t)k));
  CU(h, cudaMalloc(&h->fi_perm[which], sizeof(int) * (size_t)k));
  CU(h, cudaMemcpyAsync(h->fi_qrt[which], qrt, sizeof(double) * (size_t)k * k, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->fi_rdiag[which], rdiag, sizeof(double) * (size_t)k, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->fi_perm[which], perm, sizeof(int) * (size_t)k, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}

int als_fold_in(als_handle* h, const int32_t* users, const int32_t* items, const float* values, int64_t n) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  if (n < 0 || (n > 0 && (!users || !items))) return ALS_E_ARG;
  if (n == 0) return ALS_OK;
  if (h->world > 1) return fail(h, ALS_E_UNSUPPORTED, "fold-in runs on a single-GPU handle (the serving replica)");
  if (!h->fi_qrt[0] && !h->fi_qrt[1]) return fail(h, ALS_E_STATE, "no fold-in solver state (als_set_fold_in_state)");
  for (int64_t e = 0; e < n; e++) {
    if (users[e] < 0 || users[e] >= h->n_users) return fail(h, ALS_E_ARG, "fold-in user %d out of range", users[e]);
    if (items[e] < 0 || items[e] >= h->n_items) return fail(h, ALS_E_ARG, "fold-in item %d out of range", items[e]);
  }
  CU(h, cudaSetDevice(h->device));
  int* d_ui = nullptr;
  float* d_v = nullptr;
  CU(h, cudaMalloc(&d_ui, sizeof(int) * (size_t)(2 * n + 1)));
  if (values && cudaMalloc(&d_v, sizeof(float) * (size_t)n) != cudaSuccess) { cudaFree(d_ui); cudaGetLastError(); return fail(h, ALS_E_OOM, "fold-in scratch"); }
  cudaMemcpyAsync(d_ui, users, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(d_ui + n, items, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (values) cudaMemcpyAsync(d_v, values, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  int* d_st = d_ui + 2 * n;
  cudaMemsetAsync(d_st, 0, sizeof(int), h->stream);
  FoldInSolver sx{h->fi_qrt[0], h->fi_rdiag[0], h->fi_perm[0]}, sy{h->fi_qrt[1], h->fi_rdiag[1], h->fi_perm[1]};
  nvtxRangePushA("als:fold_in");
  switch ((h->k + 31) / 32) {
    case 1: fold_in_kernel<1><<<1, 64, 0, h->stream>>>(h->X, h->Y, h->ks, h->k, d_ui, d_ui + n, d_v, n, sx, sy, h->fi_learn_rate, d_st); break;
    case 2: fold_in_kernel<2><<<1, 64, 0, h->stream>>>(h->X, h->Y, h->ks, h->k, d_ui, d_ui + n, d_v, n, sx, sy, h->fi_learn_rate, d_st); break;
    case 3: fold_in_kernel<3><<<1, 64, 0, h->stream>>>(h->X, h->Y, h->ks, h->k, d_ui, d_ui + n, d_v, n, sx, sy, h->fi_learn_rate, d_st); break;
    default: fold_in_kernel<4><<<1, 64, 0, h->stream>>>(h->X, h->Y, h->ks, h->k, d_ui, d_ui + n, d_v, n, sx, sy, h->fi_learn_rate, d_st); break;
  }
  nvtxRangePop();
  h->launches += 1;
  int st = 0;
  cudaMemcpyAsync(&st, d_st, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_ui); cudaFree(d_v);
  if (e != cudaSuccess) return fail(h, ALS_E_CUDA, "als_fold_in: %s", cudaGetErrorString(e));
  if (st == ALS_E_NONFINITE) return fail(h, ALS_E_NONFINITE, "non-finite fold-in (ServerRecommender.java:872, :889)");
  return ALS_OK;
}

int als_gramian(als_handle* h, int32_t which, double* out) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  if (!out || (which != 0 && which != 1)) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  rc = which == 0 ? launch_gramian(h, h->X, h->n_users) : launch_gramian(h, h->Y, h->n_items);
  if (rc != ALS_OK) return rc;
  CU(h, cudaMemcpy2DAsync(out, sizeof(double) * h->k, h->G, sizeof(double) * h->ks,
                          sizeof(double) * h->k, (size_t)h->k, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}

int als_sync(als_handle* h) {
  if (!h) return ALS_E_ARG;
  CU(h, cudaSetDevice(h->device));
  // With a communicator als_sync is collective: the first error of ANY rank comes back on every
  // rank (otherwise the healthy ranks would block in the next half's collectives forever).
  int global_code = ALS_OK;
  if (h->comm) {
    CU(h, cudaMemcpyAsync(h->d_flag + 1, &h->d_status->code, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    ncclResult_t r = g_nccl.AllReduce(h->d_flag + 1, h->d_flag + 2, 1, ncclInt, ncclMax, h->comm, h->stream);
    if (r != ncclSuccess) return fail(h, ALS_E_NCCL, "ncclAllReduce(status): %s", g_nccl.GetErrorString(r));
    CU(h, cudaMemcpyAsync(&global_code, h->d_flag + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(h, cudaStreamSynchronize(h->stream));
  prof_drain(h);
  if (h->sticky != ALS_OK) return h->sticky;
  DeviceStatus st;
  CU(h, cudaMemcpy(&st, h->d_status, sizeof(st), cudaMemcpyDeviceToHost));
  if (st.code == ALS_OK) {
    if (global_code == ALS_OK) return ALS_OK;
    h->sticky = global_code;
    return fail(h, global_code, "another rank reported %s in this half-iteration",
                global_code == ALS_E_SINGULAR ? "a near-singular row" : "a failure");
  }
  h->sticky = st.code;
  const char* half = st.which == 0 ? "X" : "Y";
  if (st.code == ALS_E_SINGULAR) {
    // apparent rank of the offending W_u (rare path)
    const Csr& R = st.which == 0 ? h->by_user : h->by_item;
    const float* M = st.which == 0 ? h->Y : h->X;
    singular_rank_kernel<<<1, 1, 0, h->stream>>>(R.ptr, R.idx, R.val, st.row - R.row_begin, M, h->G,
                                                 h->ks, h->k, h->cfg.alpha,
                                                 h->cfg.lambda * h->cfg.alpha, h->cfg.reconstruct_r,
                                                 h->cfg.loss_ignores_unspecified, h->d_scratch,
                                                 h->d_rank);
    int rank = 0;
    cudaMemcpyAsync(&rank, h->d_rank, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    h->singular_rank = rank;
    return fail(h, ALS_E_SINGULAR,
                "%s-half row %lld: %d x %d matrix is near-singular (pivot %g <= threshold %g). "
                "Apparent rank: %d",
                half, st.row, h->k, h->k, (double)st.pivot, h->cfg.singularity_threshold, rank);
  }
  return fail(h, st.code, "%s-half row %lld produced a non-finite factor value", half, st.row);
}

const char* als_last_error(const als_handle* h) { return h ? h->err : "null handle"; }
int als_singular_rank(const als_handle* h) { return h ? h->singular_rank : 0; }

int als_get_info(const als_handle* h, als_info* out) {
  if (!h || !out) return ALS_E_ARG;
  memset(out, 0, sizeof(*out));
  out->struct_size = (int32_t)sizeof(als_info);
  out->features = h->k;
  out->padded_features = h->ks;
  out->kernel = h->kernel;
  out->n_users = h->n_users;
  out->n_items = h->n_items;
  out->nnz = h->by_user.nnz;
  out->device_bytes = h->device_bytes;
  out->sm_count = h->sm_count;
  out->world_size = h->world;
  out->rank = h->rank;
  return ALS_OK;
}

int als_profile_enable(als_handle* h, int32_t on) {
  if (!h) return ALS_E_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  prof_drain(h);
  h->profile = on != 0;
  return ALS_OK;
}

int als_get_timings(als_handle* h, als_timings* out, int32_t reset) {
  if (!h || !out) return ALS_E_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  prof_drain(h);
  h->tm.launches = h->launches;
  long long retried[2] = {0, 0};
  cudaMemcpy(retried, h->d_retry_total, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  h->tm.fp64_retry_rows = retried[0];
  h->tm.fp64_resolve_rows = retried[1];
  *out = h->tm;
  if (reset) {
    memset(&h->tm, 0, sizeof(h->tm));
    h->tm.struct_size = (int32_t)sizeof(als_timings);
    h->launches = 0;
    cudaMemsetAsync(h->d_retry_total, 0, 2 * sizeof(long long), h->stream);
  }
  return ALS_OK;
}

// ---- synthetic workload -----------------------------------------------------
int als_synth_interactions(als_handle* h, int64_t n_users, int64_t n_items, int32_t nnz_per_user,
                           uint64_t seed, double neg_fraction) {
  if (!h) return ALS_E_ARG;
  if (nnz_per_user <= 0 || nnz_per_user > n_items) return fail(h, ALS_E_ARG, "bad nnz_per_user");
  if (neg_fraction < 0.0 || neg_fraction > 1.0) return fail(h, ALS_E_ARG, "bad neg_fraction");
  CU(h, cudaSetDevice(h->device));
  int rc = set_dims(h, n_users, n_items);
  if (rc != ALS_OK) return rc;
  long long ub, ue, ib, ie;
  local_block(h, h->n_users, &ub, &ue);
  local_block(h, h->n_items, &ib, &ie);
  Csr& A = h->by_user;
  free_csr(h, &A);
  h->have_by_item = false;
  if ((rc = reset_for_new_interactions(h)) != ALS_OK) return rc;
  A.rows = ue - ub;
  A.nnz = A.rows * nnz_per_user;
  A.row_begin = ub;
  if ((rc = dev_alloc(h, &A.ptr, (size_t)A.rows + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &A.idx, (size_t)A.nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &A.val, (size_t)A.nnz)) != ALS_OK) return rc;
  const unsigned int thr = (unsigned int)(neg_fraction * 16777216.0);
  synth_rows_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(ub, A.rows, n_items, nnz_per_user, seed,
                                                           thr, A.ptr, A.idx, A.val);
  h->launches += 1;
  CU(h, cudaGetLastError());
  if (h->world == 1) return build_transpose(h);
  // Sharded: the by-item orientation of this rank's item block. Every rank re-draws all
  // users (counter-based generator, no communication), keeps its item block, transposes.
  Csr tmp;
  tmp.rows = n_users;
  tmp.row_begin = 0;
  long long* counts = nullptr;
  if ((rc = dev_alloc(h, &counts, (size_t)n_users + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &tmp.ptr, (size_t)n_users + 1)) != ALS_OK) return rc;
  CU(h, cudaMemsetAsync(counts, 0, sizeof(long long) * ((size_t)n_users + 1), h->stream));
  synth_item_block_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(
      n_users, n_items, nnz_per_user, seed, thr, ib, ie, counts, nullptr, nullptr, nullptr);
  {
    void* stmp = nullptr;
    size_t sbytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sbytes, counts, tmp.ptr, (int)(n_users + 1), h->stream);
    CU(h, cudaMalloc(&stmp, sbytes ? sbytes : 1));
    cub::DeviceScan::ExclusiveSum(stmp, sbytes, counts, tmp.ptr, (int)(n_users + 1), h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(stmp);
  }
  CU(h, cudaMemcpy(&tmp.nnz, tmp.ptr + n_users, sizeof(long long), cudaMemcpyDeviceToHost));
  if ((rc = dev_alloc(h, &tmp.idx, (size_t)tmp.nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &tmp.val, (size_t)tmp.nnz)) != ALS_OK) return rc;
  synth_item_block_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(
      n_users, n_items, nnz_per_user, seed, thr, ib, ie, nullptr, tmp.ptr, tmp.idx, tmp.val);
  h->launches += 4;
  CU(h, cudaGetLastError());
  rc = build_transpose_from(h, tmp, ib, ie - ib, &h->by_item);
  dev_free(h, &counts, (size_t)n_users + 1);
  free_csr(h, &tmp);
  if (rc != ALS_OK) return rc;
  h->have_by_item = true;
  return ALS_OK;
}

static unsigned long long gcd_u64(unsigned long long a, unsigned long long b) {
  while (b) { const unsigned long long t = a % b; a = b; b = t; }
  return a;
}

// rows [row_begin, row_begin + rows) of the power-law workload into `out` (see aux_kernels.cuh)
static int powerlaw_block(als_handle* h, long long row_begin, long long rows, int64_t n_items, double mean_nnz,
                          int32_t max_nnz, uint64_t seed, double neg_fraction, Csr* out) {
  Csr& A = *out;
  // E[x] of the truncated law = ln(n_max) n_max / (n_max - 1): scale the draw so the mean is mean_nnz
  const double ex = log((double)max_nnz) * (double)max_nnz / ((double)max_nnz - 1.0);
  const double scale = mean_nnz / ex;
  unsigned long long mul = (mix64(seed ^ 0x2545f4914f6cdd1dULL) % (unsigned long long)n_items) | 1ULL;
  while (gcd_u64(mul, (unsigned long long)n_items) != 1ULL) mul += 2ULL;
  const unsigned long long add = mix64(seed ^ 0x9e3779b97f4a7c15ULL) % (unsigned long long)n_items;
  A.rows = rows;
  A.row_begin = row_begin;
  int rc;
  long long* counts = nullptr;
  if ((rc = dev_alloc(h, &counts, (size_t)rows + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &A.ptr, (size_t)rows + 1)) != ALS_OK) return rc;
  powerlaw_counts_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(row_begin, rows, n_items, scale, max_nnz, seed, counts);
  {
    void* stmp = nullptr;
    size_t sbytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sbytes, counts, A.ptr, (int)(rows + 1), h->stream);
    CU(h, cudaMalloc(&stmp, sbytes ? sbytes : 1));
    cub::DeviceScan::ExclusiveSum(stmp, sbytes, counts, A.ptr, (int)(rows + 1), h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(stmp);
  }
  dev_free(h, &counts, (size_t)rows + 1);
  CU(h, cudaMemcpy(&A.nnz, A.ptr + rows, sizeof(long long), cudaMemcpyDeviceToHost));
  if ((rc = dev_alloc(h, &A.idx, (size_t)A.nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &A.val, (size_t)A.nnz)) != ALS_OK) return rc;
  const unsigned int thr = (unsigned int)(neg_fraction * 16777216.0);
  powerlaw_rows_kernel<<<h->sm_count * 8, 128, 0, h->stream>>>(row_begin, rows, n_items, seed, thr, mul, add, A.ptr,
                                                                A.idx, A.val);
  h->launches += 3;
  CU(h, cudaGetLastError());
  return ALS_OK;
}

// the entries of A whose column lies in [lo, hi), rows kept (A's row order and row_begin)
static int filter_columns(als_handle* h, const Csr& A, long long lo, long long hi, Csr* out) {
  Csr& T = *out;
  T.rows = A.rows;
  T.row_begin = A.row_begin;
  int rc;
  long long* counts = nullptr;
  if ((rc = dev_alloc(h, &counts, (size_t)A.rows + 1)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &T.ptr, (size_t)A.rows + 1)) != ALS_OK) return rc;
  filter_columns_kernel<<<h->sm_count * 8, 128, 0, h->stream>>>(A.ptr, A.rows, A.idx, A.val, (int)lo, (int)hi, counts,
                                                                 nullptr, nullptr, nullptr);
  {
    void* stmp = nullptr;
    size_t sbytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sbytes, counts, T.ptr, (int)(A.rows + 1), h->stream);
    CU(h, cudaMalloc(&stmp, sbytes ? sbytes : 1));
    cub::DeviceScan::ExclusiveSum(stmp, sbytes, counts, T.ptr, (int)(A.rows + 1), h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(stmp);
  }
  dev_free(h, &counts, (size_t)A.rows + 1);
  CU(h, cudaMemcpy(&T.nnz, T.ptr + A.rows, sizeof(long long), cudaMemcpyDeviceToHost));
  if ((rc = dev_alloc(h, &T.idx, (size_t)T.nnz)) != ALS_OK) return rc;
  if ((rc = dev_alloc(h, &T.val, (size_t)T.nnz)) != ALS_OK) return rc;
  filter_columns_kernel<<<h->sm_count * 8, 128, 0, h->stream>>>(A.ptr, A.rows, A.idx, A.val, (int)lo, (int)hi, nullptr,
                                                                 T.ptr, T.idx, T.val);
  h->launches += 3;
  CU(h, cudaGetLastError());
  return ALS_OK;
}

int als_synth_interactions_powerlaw(als_handle* h, int64_t n_users, int64_t n_items, double mean_nnz,
                                    int32_t max_nnz, uint64_t seed, double neg_fraction) {
  if (!h) return ALS_E_ARG;
  if (!(mean_nnz >= 1.0) || max_nnz < 2 || n_items < 1) return fail(h, ALS_E_ARG, "bad power-law parameters");
  if (neg_fraction < 0.0 || neg_fraction > 1.0) return fail(h, ALS_E_ARG, "bad neg_fraction");
  CU(h, cudaSetDevice(h->device));
  int rc = set_dims(h, n_users, n_items);
  if (rc != ALS_OK) return rc;
  // Sharded: every rank draws its own user block (the generator is counter-based: a row depends on
  // its global index only).  The item permutation spreads the popular items over the item blocks.
  long long ub, ue, ib, ie;
  local_block(h, h->n_users, &ub, &ue);
  local_block(h, h->n_items, &ib, &ie);
  free_csr(h, &h->by_user);
  h->have_by_item = false;
  if ((rc = reset_for_new_interactions(h)) != ALS_OK) return rc;
  if ((rc = powerlaw_block(h, ub, ue - ub, n_items, mean_nnz, max_nnz, seed, neg_fraction, &h->by_user)) != ALS_OK) return rc;
  if (h->world == 1) return build_transpose(h);
  // with a communicator the by-item blocks are exchanged on the devices
  if (h->comm) return build_by_item_distributed(h);
  // partition-only handle (no communicator): re-draw every user, keep my item block, transpose
  Csr all, mine;
  if ((rc = powerlaw_block(h, 0, n_users, n_items, mean_nnz, max_nnz, seed, neg_fraction, &all)) == ALS_OK &&
      (rc = filter_columns(h, all, ib, ie, &mine)) == ALS_OK) {
    free_csr(h, &all);
    rc = build_transpose_from(h, mine, ib, ie - ib, &h->by_item);
  }
  free_csr(h, &all);
  free_csr(h, &mine);
  if (rc != ALS_OK) return rc;
  h->have_by_item = true;
  return ALS_OK;
}

int als_synth_y0(als_handle* h, uint64_t seed) {
  int rc = check_ready(h);
  if (rc != ALS_OK) return rc;
  CU(h, cudaSetDevice(h->device));
  synth_y0_kernel<<<(int)((h->n_items + 127) / 128), 128, 0, h->stream>>>(h->Y, h->n_items, h->ks,
                                                                        h->k, seed);
  h->launches += 1;
  CU(h, cudaGetLastError());
  // a fresh build starts from an empty X (AlternatingLeastSquares.java:179)
  CU(h, cudaMemsetAsync(h->X, 0, sizeof(float) * (size_t)h->users_alloc * h->ks, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}

static int get_csr(als_handle* h, const Csr& c, int64_t* ptr, int32_t* idx, float* val) {
  if (!h || !ptr || !idx || !val) return ALS_E_ARG;
  if (!c.ptr) return fail(h, ALS_E_STATE, "interactions not set");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(ptr, c.ptr, sizeof(long long) * ((size_t)c.rows + 1), cudaMemcpyDeviceToHost,
                        h->stream));
  CU(h, cudaMemcpyAsync(idx, c.idx, sizeof(int) * (size_t)c.nnz, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(val, c.val, sizeof(float) * (size_t)c.nnz, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return ALS_OK;
}
int als_get_interactions(als_handle* h, int64_t* row_ptr, int32_t* col_idx, float* val) {
  return h ? get_csr(h, h->by_user, row_ptr, col_idx, val) : ALS_E_ARG;
}
int als_get_interactions_by_column(als_handle* h, int64_t* col_ptr, int32_t* row_idx, float* val) {
  if (!h) return ALS_E_ARG;
  if (!h->have_by_item) return fail(h, ALS_E_STATE, "by-column orientation not built");
  return get_csr(h, h->by_item, col_ptr, row_idx, val);
}

int als_get_interaction_rows(als_handle* h, int32_t by_column, int64_t first_row, int64_t n_rows,
                             int64_t* row_ptr_out, int32_t* idx_out, float* val_out,
                             int64_t capacity) {
  if (!h || !row_ptr_out || n_rows < 0 || first_row < 0 || capacity < 0) return ALS_E_ARG;
  const Csr& c = by_column ? h->by_item : h->by_user;
  if (!c.ptr || (by_column && !h->have_by_item)) return fail(h, ALS_E_STATE, "interactions not set");
  if (first_row + n_rows > c.rows) return fail(h, ALS_E_ARG, "row slice out of range");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(row_ptr_out, c.ptr + first_row, sizeof(long long) * ((size_t)n_rows + 1),
                        cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  const long long e0 = row_ptr_out[0], e1 = row_ptr_out[n_rows];
  if (e1 - e0 > capacity) return fail(h, ALS_E_ARG, "slice has %lld entries > capacity", e1 - e0);
  for (long long r = 0; r <= n_rows; r++) row_ptr_out[r] -= e0;
  if (e1 > e0) {
    if (!idx_out || !val_out) return ALS_E_ARG;
    CU(h, cudaMemcpyAsync(idx_out, c.idx + e0, sizeof(int) * (size_t)(e1 - e0),
                          cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(val_out, c.val + e0, sizeof(float) * (size_t)(e1 - e0),
                          cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  return ALS_OK;
}

// ---- development entry points (not part of include/myrrix_als.h) ------------------------------
// One warp runs the blocked Cholesky (chol_blocked.cuh) on a caller-supplied dense system:
// checks the solver in isolation from the gather / tensor-core / drain pipeline.
#ifdef ALS_SOLVE_PROF
int als_debug_solve_prof(long long* out8) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, als::g_solve_prof, sizeof(long long) * 8);
  long long z[8] = {0};
  cudaMemcpyToSymbol(als::g_solve_prof, z, sizeof(z));
  return 0;
}
#endif
static int g_debug_solve_cycles = 0;
int als_debug_last_solve_cycles(void) { return g_debug_solve_cycles; }
int als_debug_solve_blocked(const float* W, const float* b, int k, float threshold, float* x, int* ok) {
  if (!W || !b || !x || !ok || k < 1 || k > 64) return ALS_E_ARG;
  float *dW = nullptr, *db = nullptr, *dx = nullptr;
  int* dok = nullptr;
  cudaMalloc(&dW, sizeof(float) * k * k); cudaMalloc(&db, sizeof(float) * k);
  cudaMalloc(&dx, sizeof(float) * k); cudaMalloc(&dok, 2 * sizeof(int));
  cudaMemcpy(dW, W, sizeof(float) * k * k, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b, sizeof(float) * k, cudaMemcpyHostToDevice);
  if (k <= 32) debug_solve_blocked_kernel<32><<<1, 32>>>(dW, db, k, threshold, dx, dok);
  else debug_solve_blocked_kernel<64><<<1, 32>>>(dW, db, k, threshold, dx, dok);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(x, dx, sizeof(float) * k, cudaMemcpyDeviceToHost);
  int okc[2] = {0, 0};
  cudaMemcpy(okc, dok, 2 * sizeof(int), cudaMemcpyDeviceToHost);
  ok[0] = okc[0];
  g_debug_solve_cycles = okc[1];
  cudaFree(dW); cudaFree(db); cudaFree(dx); cudaFree(dok);
  return e == cudaSuccess ? ALS_OK : ALS_E_CUDA;
}

#ifdef ALS_DEBUG_SLOT
// debug build only (scripts/v2_debug.py): pick the row whose slot the kernel copies out, read it
int als_debug_set_row(long long row) {
  return cudaMemcpyToSymbol(als::v2::g_debug_row, &row, sizeof(row)) == cudaSuccess ? 0 : 1;
}
int als_debug_get_slot(float* out, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, als::v2::g_debug_slot, sizeof(float) * n) == cudaSuccess ? 0 : 1;
}
#endif

#ifdef ALS_PROFILE_WAITS
// debug build only (scripts/wait_profile.py): read and clear the wait-cycle counters
int als_debug_wait_cycles(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, als::umma::g_wait_cycles, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(als::umma::g_wait_cycles, z, sizeof(z));
  return 0;
}
#endif

// ---- multi-GPU ---------------------------------------------------------------
int als_comm_unique_id_size(void) { return (int)sizeof(ncclUniqueId); }

int als_comm_get_unique_id(void* out_id) {
  if (!out_id) return ALS_E_ARG;
  if (!load_nccl()) return ALS_E_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ALS_E_NCCL;
  memcpy(out_id, &id, sizeof(id));
  return ALS_OK;
}

int als_comm_init(als_handle* h, int32_t rank, int32_t world_size, const void* unique_id) {
  if (!h || world_size < 1 || rank < 0 || rank >= world_size) return ALS_E_ARG;
  if (h->by_user.ptr) return fail(h, ALS_E_STATE, "als_comm_init must precede als_set_interactions");
  if (!unique_id) {
    // Partition-only mode (no communicator): the handle owns block `rank` of `world_size` but
    // never exchanges factors. Lets one GPU check every rank's shard construction in turn.
    h->rank = rank;
    h->world = world_size;
    return ALS_OK;
  }
  if (!load_nccl()) return fail(h, ALS_E_NCCL, "libnccl.so.2 not found");
  CU(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, world_size, id, rank);
  if (r != ncclSuccess) return fail(h, ALS_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  h->rank = rank;
  h->world = world_size;
  return ALS_OK;
}

}  // extern "C"
