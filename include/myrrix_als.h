/*
 * myrrix_als.h -- C ABI of the B200-native ALS factorization core.
 *
 * This is the drop-in boundary for ONE path of myrrix/myrrix-recommender: the
 * implicit-feedback alternating-least-squares build that sits behind
 *   net.myrrix.online.factorizer.MatrixFactorizer
 *       (online/src/net/myrrix/online/factorizer/MatrixFactorizer.java:31-77)
 * implemented in the reference by
 *   net.myrrix.online.factorizer.als.AlternatingLeastSquares
 *       (online/src/net/myrrix/online/factorizer/als/AlternatingLeastSquares.java:66-543)
 * and called from exactly one production site,
 *   DelegateGenerationManager.RefreshCallable.runFactorization
 *       (online-local/src/net/myrrix/online/generation/DelegateGenerationManager.java:388-440).
 *
 * Conventions: one opaque handle; every entry point returns an als_status;
 * no exceptions or longjmp cross the ABI; plain pointers and sizes only; all
 * "host" pointers are caller-owned and only borrowed for the duration of the
 * call; outputs are caller-allocated.  The handle is NOT thread-safe (the
 * reference guarantees at most one factorization in flight per JVM,
 * DelegateGenerationManager.java:125-127, 238-239).
 *
 * Rows/columns are dense 0-based int32 indices; the long-ID -> dense remap
 * (slot-order walk of the FastByIDMap keys) is the shim's job, see
 * INTEGRATION.md.  Factor matrices cross the ABI row-major, `features` floats
 * per row, fp32.
 */
#ifndef MYRRIX_ALS_H_
#define MYRRIX_ALS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYRRIX_ALS_ABI_VERSION 1

typedef enum als_status {
  ALS_OK = 0,
  /* -> SingularMatrixSolverException(apparentRank), must propagate unwrapped so
   * DelegateGenerationManager.java:345-354 can lower model.features and retry. */
  ALS_E_SINGULAR = 1,
  /* a non-finite factor value was produced (GenerationSerializer.java:195-197
   * asserts finiteness; never hand such a model back) */
  ALS_E_NONFINITE = 2,
  ALS_E_CUDA = 3,
  ALS_E_NCCL = 4,
  ALS_E_OOM = 5,
  /* -> IllegalArgumentException / NullPointerException of the constructor
   * preconditions (AlternatingLeastSquares.java:137-141) */
  ALS_E_ARG = 6,
  ALS_E_UNSUPPORTED = 7,
  ALS_E_STATE = 8
} als_status;

/* Which row-update kernel family to use. AUTO picks the tcgen05 path when the
 * padded feature count allows it, else the CUDA-core (SIMT) path. */
typedef enum als_kernel {
  ALS_KERNEL_AUTO = 0,
  ALS_KERNEL_SIMT = 1,
  ALS_KERNEL_TCGEN05 = 2
} als_kernel;

/* Replaces the constructor arguments + JVM system properties the reference reads:
 *   features                  ctor arg `features`   (AlternatingLeastSquares.java:132-147)
 *   alpha                     model.als.alpha       (:71, :506-509) default 1.0
 *   lambda                    model.als.lambda      (:73, :511-514) default 0.1;
 *                             multiplied by alpha and by the row's entry count (:435, :488)
 *   reconstruct_r             model.reconstructRMatrix      (:85-86)
 *   loss_ignores_unspecified  model.lossIgnoresUnspecified  (:90-91)
 *   singularity_threshold     common.matrix.singularityThreshold
 *                             (common/.../math/LinearSystemSolver.java:33-34) default 1e-5
 */
typedef struct als_config {
  int32_t struct_size; /* = sizeof(als_config), for ABI evolution */
  int32_t features;
  double alpha;
  double lambda;
  int32_t reconstruct_r;
  int32_t loss_ignores_unspecified;
  double singularity_threshold;
  int32_t device; /* CUDA device ordinal */
  int32_t kernel; /* als_kernel */
} als_config;

typedef struct als_handle als_handle;

/* Fills *cfg with the reference defaults (features = 30, MatrixFactorizer.java:34). */
int als_config_default(als_config *cfg);

/* new AlternatingLeastSquares(...): AlternatingLeastSquares.java:132-147. */
int als_create(const als_config *cfg, als_handle **out);
int als_destroy(als_handle *h);

/* Run all library work on a caller-provided CUDA stream (cudaStream_t as void*),
 * e.g. the caller's current stream so its own events bracket the kernels.
 * NULL restores the handle's private stream. */
int als_set_stream(als_handle *h, void *cuda_stream);

/* RbyRow (ctor arg, AlternatingLeastSquares.java:132): CSR by user, HOST pointers.
 * row_ptr[n_users+1] int64, col_idx[nnz] int32 in [0,n_items), val[nnz] fp32.
 * Duplicates already summed (FastByIDFloatMap.increment, FastByIDFloatMap.java:129-138)
 * and |val| < 1e-4 already pruned (InputFilesReader.java:202-211) by the caller.
 * The by-item orientation (RbyColumn) is built on the device unless
 * als_set_interactions_by_column is also called. Rows with no entries are "not in
 * the map": their factor row is left untouched by a half-iteration. */
int als_set_interactions(als_handle *h, int64_t n_users, int64_t n_items, const int64_t *row_ptr,
                         const int32_t *col_idx, const float *val);

/* RbyColumn (ctor arg): CSR by item, HOST pointers; optional (see above). Must describe
 * the same matrix as als_set_interactions. */
int als_set_interactions_by_column(als_handle *h, const int64_t *col_ptr, const int32_t *row_idx,
                                   const float *val);

/* Same as als_set_interactions but the three arrays are DEVICE pointers on the
 * handle's device; they are copied. */
int als_set_interactions_device(als_handle *h, int64_t n_users, int64_t n_items,
                                const int64_t *d_row_ptr, const int32_t *d_col_idx,
                                const float *d_val);

/* Rows that are KEYS of RbyRow (which = 0) / RbyColumn (which = 1) but hold no entries: the
 * reference reaches them because addWorkers walks every map entry
 * (AlternatingLeastSquares.java:391-410) and removeSmall empties maps without removing them
 * (InputFilesReader.java:202-211); it solves W_u = G, b_u = 0, i.e. writes the zero vector (or
 * throws the singular exception when G is singular).  rows[n]: dense GLOBAL indices, HOST pointer;
 * every listed row must have no entries.  Call after the interactions are set; a later
 * als_set_interactions* forgets the lists.  Rows with no entries that are NOT listed stay
 * "not in the map" (stale rows: carried through untouched, still counted in M^T M). */
int als_set_present_empty_rows(als_handle *h, int32_t which, const int32_t *rows, int64_t n);

/* setPreviousY (MatrixFactorizer.java:60-64, AlternatingLeastSquares.java:171-174, 304-308):
 * complete initial Y, n_items x features, HOST pointer. Rows of items that have no
 * interactions ("stale" rows) are carried through unchanged and still count in Y^T Y
 * (MatrixUtils.java:219-239 walks every row in the map). */
int als_set_y(als_handle *h, const float *y);
/* Warm X (only needed for the convergence probe's iteration-0 estimates). */
int als_set_x(als_handle *h, const float *x);

/* iterateXFromY / iterateYFromX (AlternatingLeastSquares.java:340-389): Gramian of the
 * opposite factor + every row's W_u, b_u and solve (Worker.call, :432-504). */
int als_half_x(als_handle *h);
int als_half_y(als_handle *h);
/* n_iterations x (half_x; half_y) -- the body of the while(true) loop, :227-229. */
int als_iterate(als_handle *h, int32_t n_iterations);

/* Convergence probe (AlternatingLeastSquares.java:217-221, 231-237): out[i*n_items+j] =
 * SimpleVectorMath.dot(X[users[i]], Y[items[j]]) with fp32-rounded products summed in
 * fp64 (SimpleVectorMath.java:34-41). The stop rule itself stays in the caller. */
int als_probe(als_handle *h, const int32_t *users, int32_t n_users, const int32_t *items,
              int32_t n_items, double *out);

/* getX / getY (MatrixFactorizer.java:68-76): copy out, HOST pointers, rows x features. */
int als_get_x(als_handle *h, float *out);
int als_get_y(als_handle *h, float *out);

/* Rows [first_row, first_row + n_rows) of X (which = 0) or Y (which = 1), n_rows x features fp32,
 * HOST pointer: a rank of a sharded build reads back just the block it owns.  Call als_sync
 * first (with a communicator the peers' rows arrive asynchronously). */
int als_get_factor_block(als_handle *h, int32_t which, int64_t first_row, int64_t n_rows, float *out);

/* Selected rows of X (which = 0) or Y (which = 1): out[i] = factor[rows[i]], n x features fp32,
 * HOST pointers (the per-key lookups of getX().get(id) without copying the whole matrix). */
int als_get_rows(als_handle *h, int32_t which, const int32_t *rows, int32_t n, float *out);

/* M^T M of the current X (which=0) or Y (which=1), features x features fp64, HOST
 * pointer: MatrixUtils.transposeTimesSelf (MatrixUtils.java:219-239). */
int als_gramian(als_handle *h, int32_t which, double *out);

/* Block until all queued work is done; returns the first deferred error
 * (ALS_E_SINGULAR / ALS_E_NONFINITE / ALS_E_CUDA).  With a communicator (als_comm_init) the
 * call is COLLECTIVE: every rank must make it, and every rank gets the first error of any rank. */
int als_sync(als_handle *h);

/* Message for the last non-OK status on this handle (never NULL). */
const char *als_last_error(const als_handle *h);
/* SingularMatrixSolverException.getApparentRank() for the last ALS_E_SINGULAR
 * (CommonsMathLinearSystemSolver.java:46-54: RRQR getRank(0.01)); 0 if none. */
int als_singular_rank(const als_handle *h);

/* ---- introspection / measurement (no reference counterpart) ---------------- */

typedef struct als_info {
  int32_t struct_size;
  int32_t features;
  int32_t padded_features; /* row stride of the device factor matrices */
  int32_t kernel;          /* als_kernel actually selected */
  int64_t n_users, n_items, nnz;
  int64_t device_bytes; /* bytes of HBM currently held by this handle */
  int32_t sm_count;
  int32_t world_size, rank;
} als_info;
int als_get_info(const als_handle *h, als_info *out);

/* Per-kernel device timing with CUDA events on the launching stream. When enabled,
 * every half-iteration records events around its Gramian and row-update launches. */
typedef struct als_timings {
  int32_t struct_size;
  int32_t launches;         /* kernels launched since last reset */
  double gramian_ms;        /* summed */
  double update_x_ms;       /* row-update kernel over users, summed */
  double update_y_ms;       /* row-update kernel over items, summed */
  double exchange_ms;       /* multi-GPU factor exchange, summed */
  int32_t n_half_x, n_half_y;
  int64_t fp64_retry_rows;  /* rows the fp32 tensor-core path handed to the fp64 kernels */
  int64_t fp64_resolve_rows; /* of those: solved in fp64 from the stashed data term, without gathering the row again */
} als_timings;
int als_profile_enable(als_handle *h, int32_t on);
int als_get_timings(als_handle *h, als_timings *out, int32_t reset);

/* ---- synthetic workload, generated on the device (bench/test support) ------- */
/* Counter-based generator keyed (seed,row,j) per SURVEY.md 8(d): each user draws
 * nnz_per_user distinct items uniformly, strength uniform in {1..5}; a fraction
 * neg_fraction of strengths is negated. Fills the handle's interactions (both
 * orientations) without touching the host. */
int als_synth_interactions(als_handle *h, int64_t n_users, int64_t n_items, int32_t nnz_per_user,
                           uint64_t seed, double neg_fraction);
/* Power-law variant: entries per user from a truncated power law (density
 * ~ x^-2 on [1, max_nnz], scaled to a mean of about mean_nnz, at least one), item popularity
 * Zipf(s = 1) over a pseudo-random permutation of the items, no item twice per user; strengths as
 * above.  Sharded handles draw their own user block (the by-item blocks are exchanged on the devices). */
int als_synth_interactions_powerlaw(als_handle *h, int64_t n_users, int64_t n_items, double mean_nnz,
                                    int32_t max_nnz, uint64_t seed, double neg_fraction);
/* Y0: every row k i.i.d. N(0,1) normalised to unit L2 norm in fp32 (the distribution of
 * RandomUtils.doRandomUnitVector, common/.../random/RandomUtils.java:88-100). */
int als_synth_y0(als_handle *h, uint64_t seed);
/* Copy the handle's by-user CSR back to HOST buffers (so a host-side checker can see
 * exactly what the device generated). */
int als_get_interactions(als_handle *h, int64_t *row_ptr, int32_t *col_idx, float *val);
int als_get_interactions_by_column(als_handle *h, int64_t *col_ptr, int32_t *row_idx, float *val);
/* Copy rows [first_row, first_row+n_rows) of one orientation (by_column = 0: by user,
 * 1: by item) to HOST buffers; row_ptr_out[n_rows+1] is rebased to start at 0 and at most
 * `capacity` entries are written to idx_out/val_out (ALS_E_ARG if the slice is larger). */
int als_get_interaction_rows(als_handle *h, int32_t by_column, int64_t first_row, int64_t n_rows,
                             int64_t *row_ptr_out, int32_t *idx_out, float *val_out,
                             int64_t capacity);

/* AlternatingLeastSquares.call's iteration loop (AlternatingLeastSquares.java:206-257) in one
 * call: X<-Y, Y<-X, then the convergence statistic -- the DoubleWeightedMean of |new - old| estimate
 * weighted by max(0, new) over test_users x test_items (:232-240), evaluated on the device -- until
 * the iteration limit (max_iterations > 0, :242-245), a non-finite statistic (:248-251) or
 * statistic < convergence_threshold, except after iteration 1 of a build that started from a random
 * Y (random_y != 0, :253-256).  The test IDs are the caller's choice (chooseAboutNFromStream walks
 * the caller's maps, :206-213); x_is_empty != 0: first ever build, estimates start at 0 (:215-223).
 * Outputs may be NULL.  Errors of a row update (ALS_E_SINGULAR, ...) come back from this call. */
int als_call(als_handle *h, const int32_t *test_users, int32_t n_test_users,
             const int32_t *test_items, int32_t n_test_items, int32_t max_iterations,
             double convergence_threshold, int32_t random_y, int32_t x_is_empty,
             int32_t *iterations_run, double *last_convergence_value);

/* ---- the live model between builds ------------------------- */
/* Generation.recomputeState (online/src/net/myrrix/online/generation/Generation.java:132-158) for a
 * model that stays in HBM: als_gramian gives X'X / Y'Y from the resident factors, libmyrrix_foldin.so
 * (myrrix_foldin.h) applies the infNorm guard and factorises them once per generation, and this call
 * keeps a copy of the factors (foldin_export_solver) on the device.  which: 0 = X'X solver,
 * 1 = Y'Y solver; qrt == NULL: no solver for that side (model.solver.xtx.compute=false).
 * learn_rate: model.foldin.learningRate. */
int als_set_fold_in_state(als_handle *h, int32_t which, const double *qrt, const double *rdiag,
                          const int32_t *perm, double learn_rate);
/* ServerRecommender.updateFeatures (online/.../ServerRecommender.java:865-907) for n writes
 * (user, item, value; values == NULL: all 1.0), applied in order to the resident rows of X and Y:
 * each write sees what the previous ones left, like the reference's one-call-at-a-time stream.
 * als_recommend and als_get_rows see the updated rows at once. */
int als_fold_in(als_handle *h, const int32_t *users, const int32_t *items, const float *values,
                int64_t n);

/* ---- top-N scoring on the resident model ------------------- */
/* Replaces, for dense indices, ServerRecommender.recommend / recommendToMany
 * (online/src/net/myrrix/online/ServerRecommender.java:355-441) -> multithreadedTopN (:443-509)
 * -> RecommendIterator.next (online/.../RecommendIterator.java:68-110) -> TopN
 * (common/src/net/myrrix/common/TopN.java:55-131): every item row of Y is scored as
 * (float) (sum over the query's users of SimpleVectorMath.dot(y_i, x_u) / n_users), bit-identical
 * to the reference's arithmetic, and the best `how_many` come back ordered by value descending,
 * equal values by ascending item (ByValueAscComparator reversed).  Filtered, like the reference:
 * consider_known_items == 0 removes the intersection of the known items (the users' rows of R)
 * of those users that have any (:396-425); `exclude` lists further items to skip (the tag IDs /
 * IDRescorer.isFiltered of the caller; the callbacks themselves stay in the host language).
 * out_items / out_values: how_many entries, *out_count of them valid.  1 <= how_many <= 128,
 * 1 <= n_users <= 16.  ALS_E_NONFINITE: "Bad recommendation value" (RecommendIterator.java:99). */
int als_recommend(als_handle *h, const int32_t *users, int32_t n_users, int32_t how_many,
                  int32_t consider_known_items, const int32_t *exclude, int32_t n_exclude,
                  int32_t *out_items, float *out_values, int32_t *out_count);
/* One single-user query per entry of `users` (the loop of AllRecommendations.call,
 * web-common/src/net/myrrix/web/AllRecommendations.java:77-117): four queries share each pass
 * over Y.  out_items / out_values: [n_queries][how_many], out_counts: [n_queries]. */
int als_recommend_batch(als_handle *h, const int32_t *users, int64_t n_queries, int32_t how_many,
                        int32_t consider_known_items, int32_t *out_items, float *out_values,
                        int32_t *out_counts);
/* The same scoring for caller-supplied feature vectors (features: host, [n_vectors][features]):
 * which = 1 scores the item rows (recommendToAnonymous, ServerRecommender.java:511-560, after the
 * caller's fold-in), which = 0 the user rows. */
int als_top_n(als_handle *h, int32_t which, const float *features, int32_t n_vectors,
              const int32_t *exclude, int32_t n_exclude, int32_t how_many, int32_t *out_ids,
              float *out_values, int32_t *out_count);

/* ---- multi-GPU: one process per GPU, user/item ranges sharded over ranks ----- */
/* Size in bytes of the opaque NCCL unique id, and creation of one (rank 0). */
int als_comm_unique_id_size(void);
int als_comm_get_unique_id(void *out_id);
/* Join a communicator: every rank calls with the same id. After this the handle
 * updates only its own contiguous block of users (X half) / items (Y half) and
 * all-gathers the fresh block over NVLink after each half. unique_id == NULL selects a
 * partition-only mode (no communicator, no exchange) used to test shard construction. */
int als_comm_init(als_handle *h, int32_t rank, int32_t world_size, const void *unique_id);

int als_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MYRRIX_ALS_H_ */
