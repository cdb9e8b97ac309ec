/* myrrix_foldin.h -- C ABI of the fold-in math that follows a model build
 * (libmyrrix_foldin.so; plain C++, no CUDA: k x k fp64 work per online write).
 *
 * Replaces, for the path "new (user, item, strength) datum -> updated factor rows":
 *   Generation.recomputeState / recomputeSolver   online/src/net/myrrix/online/generation/Generation.java:132-158
 *       (X'X and Y'Y, the infNorm < 1 ill-conditioning guard, MatrixUtils.getSolver)
 *   CommonsMathLinearSystemSolver.getSolver       common/src/net/myrrix/common/math/CommonsMathLinearSystemSolver.java:36-55
 *   CommonsMathSolver.solveFToD                   common/src/net/myrrix/common/math/CommonsMathSolver.java:46-57
 *   ServerRecommender.updateFeatures              online/src/net/myrrix/online/ServerRecommender.java:865-907
 *   ServerRecommender.foldInWeight                ...ServerRecommender.java:981-994
 *   ServerRecommender.buildAnonymousUserFeatures  ...ServerRecommender.java:561-608 (the arithmetic; the
 *       ID lookups and locks stay with the caller)
 * The two Gramians come from the ALS handle (als_gramian in myrrix_als.h: computed on the GPU
 * from the resident factors right after a build) or from any k x k row-major fp64 array.
 * The solver is commons-math3 3.2's RRQRDecomposition (third party, pom.xml:81), restated:
 * Householder QR with column pivoting, |R_jj| <= threshold => singular, rank = getRank(0.01).
 * One handle per generation; its methods are read-only after creation (thread-safe).
 */
#ifndef MYRRIX_FOLDIN_H
#define MYRRIX_FOLDIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct foldin_handle foldin_handle;

enum foldin_status {
  FOLDIN_OK = 0,
  FOLDIN_E_ARG = 1,
  FOLDIN_E_ILL_CONDITIONED = 2, /* IllConditionedSolverException: infNorm(M'M) < 1 (Generation.java:147-151) */
  FOLDIN_E_SINGULAR = 3,        /* SingularMatrixSolverException(apparentRank) */
  FOLDIN_E_NOT_READY = 4,       /* the solver asked for was not given (NotReadyException / null solver) */
  FOLDIN_E_NONFINITE = 5,       /* Preconditions.checkState(isFinite(...)) */
  FOLDIN_E_OOM = 6
};

/* xtx / yty: k x k row-major fp64, either may be NULL (model.solver.xtx.compute=false etc.).
 * singularity_threshold: common.matrix.singularityThreshold (1e-5). learn_rate:
 * model.foldin.learningRate (1.0). On FOLDIN_E_SINGULAR / ILL_CONDITIONED *which_failed is
 * 0 (X'X) or 1 (Y'Y) and *apparent_rank the reported rank (singular only). */
int foldin_create(int32_t features, const double* xtx, const double* yty, double singularity_threshold,
                  double learn_rate, foldin_handle** out, int32_t* which_failed, int32_t* apparent_rank);
void foldin_destroy(foldin_handle* h);

/* foldInWeight(estimate, value), learn rate included. */
double foldin_weight(const foldin_handle* h, double estimate, float value);
/* x = (M'M)^-1 b  (Solver.solveFToD); which: 0 = X'X, 1 = Y'Y. */
int foldin_solve(const foldin_handle* h, int32_t which, const float* b, double* x);
/* updateFeatures: both rows are updated in place from their values on entry. */
int foldin_update_features(const foldin_handle* h, float* user_features, float* item_features, float value);
/* n writes applied one after the other (each sees the rows the previous ones left): X [n_users][k]
 * and Y [n_items][k] row-major, dense row indices; values may be NULL (all 1.0). The reference
 * does this one setPreference call at a time (ServerRecommender.java:735-760, bulk ingest). */
int foldin_update_many(const foldin_handle* h, float* X, float* Y, const int32_t* users,
                       const int32_t* items, const float* values, int64_t n);
/* The factors of one solver, for a device-side copy (als_set_fold_in_state in myrrix_als.h):
 * qrt [k][k] (row `minor` = Householder vector / R column of step `minor`), rdiag [k] (R's
 * diagonal), perm [k] (perm[j] = original column now at position j). */
int foldin_export_solver(const foldin_handle* h, int32_t which, double* qrt, double* rdiag, int32_t* perm);
/* buildAnonymousUserFeatures over the n item rows that were found (row-major n x k); values may
 * be NULL (all 1.0). out: k floats. */
int foldin_anonymous_user(const foldin_handle* h, const float* item_rows, const float* values, int32_t n,
                          float* out);

#ifdef __cplusplus
}
#endif
#endif
