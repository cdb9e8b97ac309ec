/*
 * als_oracle.c -- CPU restatement of the reference's ALS factorization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object, and only as the checker or
 * the timed CPU baseline -- never as a fallback for the CUDA library.
 *
 * The reference (myrrix/myrrix-recommender) is Java 6; there is no JVM in this
 * image, so the reference itself cannot be run here.  This file restates, line
 * by line and in fp64 exactly as the Java does, the arithmetic of:
 *
 *   online/src/net/myrrix/online/factorizer/als/AlternatingLeastSquares.java
 *       :177-262  call()  (driver, stop rule)
 *       :340-389  iterateXFromY / iterateYFromX
 *       :432-504  Worker.call()  (per-row W, b, solve)
 *       :524-539  partialTransposeTimesSelf (LOSS_IGNORES_UNSPECIFIED)
 *   common/src/net/myrrix/common/math/MatrixUtils.java:219-239  transposeTimesSelf
 *   common/src/net/myrrix/common/math/CommonsMathLinearSystemSolver.java:36-55
 *   common/src/net/myrrix/common/math/CommonsMathSolver.java:36-44 (fp64 -> fp32 cast)
 *   common/src/net/myrrix/common/math/LinearSystemSolver.java:33-34 (threshold 1e-5)
 *   common/src/net/myrrix/common/math/SimpleVectorMath.java:34-41 (dot)
 *   common/src/net/myrrix/common/stats/DoubleWeightedMean.java:73-81
 *
 * Third-party arithmetic NOT under /root/reference: org.apache.commons:commons-math3:3.2
 * (pom.xml:81) RRQRDecomposition / QRDecomposition.Solver.  Its published algorithm
 * (Householder QR on the transposed copy with column pivoting on the largest
 * squared 2-norm, |rDiag| <= threshold => singular, solve = reflect b, back
 * substitute, un-permute; getRank(dropThreshold) via Frobenius norms of trailing
 * blocks of R) is restated in rrqr_* below from the 3.2 sources as recalled.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against every
 * golden vector the reference's own tests hold for this path
 * (AlternatingLeastSquaresTest.java:39-78, NegativeInputTest.java:37-80,
 * MatrixUtilsTest.java:63-73, SimpleVectorMathTest.java:29-37), all at k=2/3.
 * For k in {16,32,64,128} the reference pins nothing; parity there is
 * "CUDA vs this restated oracle".
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fPIC -shared -pthread).
 * -ffp-contract=off: Java never fuses a*b+c, so neither may the compiler here.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_E_SINGULAR 1
#define ORACLE_E_ARG 6
#define ORACLE_E_OOM 5

/* AlternatingLeastSquares.java:77 */
#define WORK_UNIT_SIZE 100

/* ------------------------------------------------------------------------- */
/* MatrixUtils.transposeTimesSelf (MatrixUtils.java:219-239).
 * `rowValue * vector[col]` is float*float -> rounded to fp32, THEN widened and
 * added into the fp64 entry (Array2DRowRealMatrix.addToEntry).  Full k x k. */
void oracle_transpose_times_self(const float *M, int64_t n_rows, int k, double *G) {
  memset(G, 0, sizeof(double) * (size_t)k * (size_t)k);
  for (int64_t r = 0; r < n_rows; r++) {
    const float *vector = M + r * (int64_t)k;
    for (int row = 0; row < k; row++) {
      float rowValue = vector[row];
      double *Grow = G + (size_t)row * (size_t)k;
      for (int col = 0; col < k; col++) {
        float prod = rowValue * vector[col];
        Grow[col] += (double)prod;
      }
    }
  }
}

/* Same, restricted to a subset of rows given by index (used when only rows
 * present in the factor map count, and by partialTransposeTimesSelf,
 * AlternatingLeastSquares.java:524-539). */
static void transpose_times_self_subset(const float *M, const int32_t *idx, int64_t n, int k,
                                        double *G) {
  memset(G, 0, sizeof(double) * (size_t)k * (size_t)k);
  for (int64_t t = 0; t < n; t++) {
    const float *vector = M + (int64_t)idx[t] * (int64_t)k;
    for (int row = 0; row < k; row++) {
      float rowValue = vector[row];
      double *Grow = G + (size_t)row * (size_t)k;
      for (int col = 0; col < k; col++) {
        float prod = rowValue * vector[col];
        Grow[col] += (double)prod;
      }
    }
  }
}

/* SimpleVectorMath.dot (SimpleVectorMath.java:34-41): float*float product
 * rounded to fp32, summed in fp64. */
double oracle_dot(const float *x, const float *y, int n) {
  double dot = 0.0;
  for (int i = 0; i < n; i++) {
    float p = x[i] * y[i];
    dot += (double)p;
  }
  return dot;
}

/* SimpleVectorMath.norm(float[]) (SimpleVectorMath.java:46-52) */
double oracle_norm(const float *x, int n) {
  double total = 0.0;
  for (int i = 0; i < n; i++) {
    float p = x[i] * x[i];
    total += (double)p;
  }
  return sqrt(total);
}

/* ------------------------------------------------------------------------- */
/* commons-math3 3.2 RRQRDecomposition, restated.  qrt is the TRANSPOSE of the
 * input: qrt[col][row], stored here as k pointers so column swaps are pointer
 * swaps exactly as in the Java (`double[] tmp1 = qrt[minor]; ...`). */
typedef struct {
  int k;
  double *store;  /* k*k */
  double **qrt;   /* k column pointers */
  double *rDiag;  /* k */
  int *p;         /* k */
  double *y;      /* k scratch */
  double *x;      /* k scratch */
} rrqr_t;

static int rrqr_alloc(rrqr_t *q, int k) {
  q->k = k;
  q->store = (double *)malloc(sizeof(double) * (size_t)k * (size_t)k);
  q->qrt = (double **)malloc(sizeof(double *) * (size_t)k);
  q->rDiag = (double *)malloc(sizeof(double) * (size_t)k);
  q->p = (int *)malloc(sizeof(int) * (size_t)k);
  q->y = (double *)malloc(sizeof(double) * (size_t)k);
  q->x = (double *)malloc(sizeof(double) * (size_t)k);
  return (q->store && q->qrt && q->rDiag && q->p && q->y && q->x) ? 0 : -1;
}

static void rrqr_free(rrqr_t *q) {
  free(q->store); free(q->qrt); free(q->rDiag); free(q->p); free(q->y); free(q->x);
}

/* QRDecomposition.performHouseholderReflection (commons-math3 3.2) */
static void householder(rrqr_t *q, int minor) {
  int k = q->k;
  double *qrtMinor = q->qrt[minor];
  double xNormSqr = 0.0;
  for (int row = minor; row < k; row++) {
    double c = qrtMinor[row];
    xNormSqr += c * c;
  }
  double a = (qrtMinor[minor] > 0) ? -sqrt(xNormSqr) : sqrt(xNormSqr);
  q->rDiag[minor] = a;
  if (a != 0.0) {
    qrtMinor[minor] -= a;
    for (int col = minor + 1; col < k; col++) {
      double *qrtCol = q->qrt[col];
      double alpha = 0.0;
      for (int row = minor; row < k; row++) {
        alpha -= qrtCol[row] * qrtMinor[row];
      }
      alpha /= a * qrtMinor[minor];
      for (int row = minor; row < k; row++) {
        qrtCol[row] -= alpha * qrtMinor[row];
      }
    }
  }
}

/* RRQRDecomposition(matrix, threshold): transpose copy, pivot on the column of
 * greatest squared 2-norm (3.2 takes the norm over the WHOLE column, j = 0..),
 * swap, record permutation, reflect. W is row-major k x k. */
static void rrqr_decompose(rrqr_t *q, const double *W) {
  int k = q->k;
  for (int c = 0; c < k; c++) {
    q->qrt[c] = q->store + (size_t)c * (size_t)k;
    for (int r = 0; r < k; r++) {
      q->qrt[c][r] = W[(size_t)r * (size_t)k + (size_t)c];
    }
    q->p[c] = c;
  }
  for (int minor = 0; minor < k; minor++) {
    double l2NormSquaredMax = 0.0;
    int l2NormSquaredMaxIndex = minor;
    for (int i = minor; i < k; i++) {
      double l2NormSquared = 0.0;
      const double *col = q->qrt[i];
      for (int j = 0; j < k; j++) {
        l2NormSquared += col[j] * col[j];
      }
      if (l2NormSquared > l2NormSquaredMax) {
        l2NormSquaredMax = l2NormSquared;
        l2NormSquaredMaxIndex = i;
      }
    }
    if (l2NormSquaredMaxIndex != minor) {
      double *tmp1 = q->qrt[minor];
      q->qrt[minor] = q->qrt[l2NormSquaredMaxIndex];
      q->qrt[l2NormSquaredMaxIndex] = tmp1;
      int tmp2 = q->p[minor];
      q->p[minor] = q->p[l2NormSquaredMaxIndex];
      q->p[l2NormSquaredMaxIndex] = tmp2;
    }
    householder(q, minor);
  }
}

/* QRDecomposition.Solver.isNonSingular: every |rDiag| > threshold */
static int rrqr_is_nonsingular(const rrqr_t *q, double threshold) {
  for (int i = 0; i < q->k; i++) {
    if (fabs(q->rDiag[i]) <= threshold) return 0;
  }
  return 1;
}

/* RRQRDecomposition.getRank(dropThreshold) */
static int rrqr_rank(const rrqr_t *q, double dropThreshold) {
  int k = q->k;
  /* R[row][col] = rDiag[row] (row==col), qrt[col][row] (col>row), else 0 */
  double *sq = (double *)calloc((size_t)k + 1, sizeof(double)); /* sq[s] = ||R[s:,s:]||_F^2 */
  if (!sq) return 0;
  for (int s = k - 1; s >= 0; s--) {
    double acc = q->rDiag[s] * q->rDiag[s];
    for (int col = s + 1; col < k; col++) {
      double v = q->qrt[col][s];
      acc += v * v;
    }
    sq[s] = sq[s + 1] + acc;
  }
  int rank = 1;
  double lastNorm = sqrt(sq[0]);
  double rNorm = lastNorm;
  while (rank < k) {
    double thisNorm = sqrt(sq[rank]);
    if (thisNorm == 0 || (thisNorm / lastNorm) * rNorm < dropThreshold) break;
    lastNorm = thisNorm;
    rank++;
  }
  free(sq);
  return rank;
}

/* QRDecomposition.Solver.solve + RRQR un-permute (x_out[p[i]] = x[i]), then
 * CommonsMathSolver.solveDToF's (float) cast (CommonsMathSolver.java:36-44). */
static void rrqr_solve_to_float(rrqr_t *q, const double *b, float *out) {
  int k = q->k;
  double *y = q->y, *x = q->x;
  memcpy(y, b, sizeof(double) * (size_t)k);
  for (int minor = 0; minor < k; minor++) {
    const double *qrtMinor = q->qrt[minor];
    double dotProduct = 0.0;
    for (int row = minor; row < k; row++) dotProduct += y[row] * qrtMinor[row];
    dotProduct /= q->rDiag[minor] * qrtMinor[minor];
    for (int row = minor; row < k; row++) y[row] += dotProduct * qrtMinor[row];
  }
  for (int row = k - 1; row >= 0; --row) {
    y[row] /= q->rDiag[row];
    double yRow = y[row];
    const double *qrtRow = q->qrt[row];
    x[row] = yRow;
    for (int i = 0; i < row; i++) y[i] -= yRow * qrtRow[i];
  }
  for (int i = 0; i < k; i++) out[q->p[i]] = (float)x[i];
}

/* Stand-alone solve for tests: W row-major k x k (destroyed: no), b fp64, out fp32.
 * Returns ORACLE_OK or ORACLE_E_SINGULAR with *apparent_rank = getRank(0.01)
 * (CommonsMathLinearSystemSolver.java:41-54). */
int oracle_solve(const double *W, const double *b, int k, double threshold, float *out,
                 int *apparent_rank) {
  rrqr_t q;
  if (rrqr_alloc(&q, k) != 0) { rrqr_free(&q); return ORACLE_E_OOM; }
  rrqr_decompose(&q, W);
  int rc = ORACLE_OK;
  if (!rrqr_is_nonsingular(&q, threshold)) {
    if (apparent_rank) *apparent_rank = rrqr_rank(&q, 0.01);
    rc = ORACLE_E_SINGULAR;
  } else {
    rrqr_solve_to_float(&q, b, out);
  }
  rrqr_free(&q);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* One half-iteration: for every row u of R (CSR) with at least one entry,
 * out[u] = solve(W_u, b_u).  Worker.call, AlternatingLeastSquares.java:432-504.
 * Rows with no entries are not in the reference's map and are left untouched. */
typedef struct {
  const int64_t *row_ptr;
  const int32_t *col_idx;
  const float *val;
  const float *M;   /* opposite factor, n_other x k */
  const double *G;  /* M^T M, k x k */
  float *out;       /* n_rows x k */
  int64_t n_rows;
  int k;
  double alpha, lambda;
  int reconstruct_r, loss_ignores_unspecified;
  double threshold;
  /* optional: present[u] != 0 marks a row that is a KEY of the reference's map even though it
   * has no entries (InputFilesReader.removeSmall, :202-211, empties maps without removing
   * them).  addWorkers walks every map entry (ALS.java:391-410), so such a row is solved with
   * Wu = YTY (lambda * 0 on the diagonal) and an all-zero right-hand side. NULL = none. */
  const uint8_t *present;
  /* work queue */
  int64_t next_unit;
  int64_t n_units;
  pthread_mutex_t mu;
  int status;
  int apparent_rank;
} half_job_t;

static void *half_worker(void *arg) {
  half_job_t *job = (half_job_t *)arg;
  int k = job->k;
  size_t kk = (size_t)k * (size_t)k;
  double *Wu = (double *)malloc(sizeof(double) * kk);
  double *YTCupu = (double *)malloc(sizeof(double) * (size_t)k);
  rrqr_t q;
  int alloc_ok = (rrqr_alloc(&q, k) == 0) && Wu && YTCupu;
  double alpha = job->alpha;
  double lambda = job->lambda * alpha; /* ALS.java:435 */
  for (;;) {
    pthread_mutex_lock(&job->mu);
    int64_t unit = job->next_unit++;
    int stop = (job->status != ORACLE_OK);
    pthread_mutex_unlock(&job->mu);
    if (stop || unit >= job->n_units) break;
    if (!alloc_ok) {
      pthread_mutex_lock(&job->mu); job->status = ORACLE_E_OOM; pthread_mutex_unlock(&job->mu);
      break;
    }
    int64_t u0 = unit * WORK_UNIT_SIZE;
    int64_t u1 = u0 + WORK_UNIT_SIZE;
    if (u1 > job->n_rows) u1 = job->n_rows;
    for (int64_t u = u0; u < u1; u++) {
      int64_t e0 = job->row_ptr[u], e1 = job->row_ptr[u + 1];
      int64_t nu = e1 - e0; /* ru.size() */
      if (nu == 0 && !(job->present && job->present[u])) continue; /* not a key of the map */
      /* Wu = YTY.copy() or the partial variant (ALS.java:447-450) */
      if (job->loss_ignores_unspecified) {
        transpose_times_self_subset(job->M, job->col_idx + e0, nu, k, Wu);
      } else {
        memcpy(Wu, job->G, sizeof(double) * kk);
      }
      memset(YTCupu, 0, sizeof(double) * (size_t)k);
      for (int64_t e = e0; e < e1; e++) {
        double xu = (double)job->val[e];
        const float *vector = job->M + (int64_t)job->col_idx[e] * (int64_t)k;
        if (job->reconstruct_r) { /* ALS.java:466-469 */
          for (int row = 0; row < k; row++) YTCupu[row] += xu * (double)vector[row];
        } else { /* ALS.java:470-483 */
          double cu = 1.0 + alpha * fabs(xu);
          for (int row = 0; row < k; row++) {
            float vectorAtRow = vector[row];
            double rowValue = (double)vectorAtRow * (cu - 1.0);
            double *WuDataRow = Wu + (size_t)row * (size_t)k;
            for (int col = 0; col < k; col++) {
              WuDataRow[col] += rowValue * (double)vector[col];
            }
            if (xu > 0.0) {
              YTCupu[row] += (double)vectorAtRow * cu;
            }
          }
        }
      }
      double lambdaTimesCount = lambda * (double)nu; /* ALS.java:488-492 */
      for (int x = 0; x < k; x++) Wu[(size_t)x * (size_t)k + (size_t)x] += lambdaTimesCount;
      /* MatrixUtils.getSolver(Wu).solveDToF(YTCupu), ALS.java:494 */
      rrqr_decompose(&q, Wu);
      if (!rrqr_is_nonsingular(&q, job->threshold)) {
        int rank = rrqr_rank(&q, 0.01);
        pthread_mutex_lock(&job->mu);
        if (job->status == ORACLE_OK) { job->status = ORACLE_E_SINGULAR; job->apparent_rank = rank; }
        pthread_mutex_unlock(&job->mu);
        goto done;
      }
      rrqr_solve_to_float(&q, YTCupu, job->out + u * (int64_t)k);
    }
  }
done:
  rrqr_free(&q);
  free(Wu);
  free(YTCupu);
  return NULL;
}

int oracle_als_half_p(const int64_t *row_ptr, const int32_t *col_idx, const float *val,
                      int64_t n_rows, const float *M, const double *G, int k, double alpha,
                      double lambda, int reconstruct_r, int loss_ignores_unspecified,
                      double threshold, int n_threads, float *out, int *apparent_rank,
                      const uint8_t *present);

int oracle_als_half(const int64_t *row_ptr, const int32_t *col_idx, const float *val,
                    int64_t n_rows, const float *M, const double *G, int k, double alpha,
                    double lambda, int reconstruct_r, int loss_ignores_unspecified,
                    double threshold, int n_threads, float *out, int *apparent_rank) {
  return oracle_als_half_p(row_ptr, col_idx, val, n_rows, M, G, k, alpha, lambda, reconstruct_r,
                           loss_ignores_unspecified, threshold, n_threads, out, apparent_rank, NULL);
}

int oracle_als_half_p(const int64_t *row_ptr, const int32_t *col_idx, const float *val,
                      int64_t n_rows, const float *M, const double *G, int k, double alpha,
                      double lambda, int reconstruct_r, int loss_ignores_unspecified,
                      double threshold, int n_threads, float *out, int *apparent_rank,
                      const uint8_t *present) {
  if (k <= 0 || n_rows < 0) return ORACLE_E_ARG;
  half_job_t job;
  memset(&job, 0, sizeof(job));
  job.present = present;
  job.row_ptr = row_ptr; job.col_idx = col_idx; job.val = val; job.M = M; job.G = G;
  job.out = out; job.n_rows = n_rows; job.k = k; job.alpha = alpha; job.lambda = lambda;
  job.reconstruct_r = reconstruct_r; job.loss_ignores_unspecified = loss_ignores_unspecified;
  job.threshold = threshold;
  job.n_units = (n_rows + WORK_UNIT_SIZE - 1) / WORK_UNIT_SIZE;
  job.status = ORACLE_OK;
  pthread_mutex_init(&job.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  if (n_threads == 1) {
    half_worker(&job);
  } else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, half_worker, &job);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  pthread_mutex_destroy(&job.mu);
  if (job.status == ORACLE_E_SINGULAR && apparent_rank) *apparent_rank = job.apparent_rank;
  return job.status;
}

/* ------------------------------------------------------------------------- */
/* DoubleWeightedMean.increment (DoubleWeightedMean.java:73-81) */
typedef struct { double totalWeight, mean; } dwm_t;
static void dwm_increment(dwm_t *m, double datum, double weight) {
  double oldTotalWeight = m->totalWeight;
  m->totalWeight += weight;
  if (oldTotalWeight <= 0) {
    m->mean = datum;
  } else {
    m->mean = m->mean * oldTotalWeight / m->totalWeight + datum * weight / m->totalWeight;
  }
}

/* The convergence probe body of call() (ALS.java:230-238): updates estimates in
 * place and returns the weighted mean of |new-old| with weight max(0,new). */
double oracle_convergence_probe(const float *X, const float *Y, int k, const int32_t *test_users,
                                int n_tu, const int32_t *test_items, int n_ti,
                                double *estimates) {
  dwm_t m = {0.0, NAN};
  for (int i = 0; i < n_tu; i++) {
    for (int j = 0; j < n_ti; j++) {
      double newValue = oracle_dot(X + (int64_t)test_users[i] * k, Y + (int64_t)test_items[j] * k, k);
      double oldValue = estimates[(size_t)i * (size_t)n_ti + (size_t)j];
      estimates[(size_t)i * (size_t)n_ti + (size_t)j] = newValue;
      double w = newValue > 0.0 ? newValue : 0.0; /* FastMath.max(0.0, newValue) */
      dwm_increment(&m, fabs(newValue - oldValue), w);
    }
  }
  return m.mean;
}

/* Full driver: AlternatingLeastSquares.call (ALS.java:177-262) with Y0 supplied
 * through setPreviousY (same k => adopted in place, :304-308), i.e. randomY=false.
 *   R   : CSR by user (row_ptr/col_idx/val), n_users rows
 *   RT  : CSR by item  (the reference is handed both orientations, ALS.java:132-136)
 *   X   : n_users x k, in: zeros (first build => estimates start at 0, :215-223), out: result
 *   Y   : n_items x k, in: Y0, out: result
 *   random_y: 1 mirrors the "don't converge after 1 iteration" guard (:253)
 *   test_users/test_items: the convergence sample (all rows when <= 100, RandomUtils.java:207-211)
 * Returns status; *iterations_run = number of completed iterations. */
int oracle_als_run_p(const int64_t *r_ptr, const int32_t *r_idx, const float *r_val, int64_t n_users,
                     const int64_t *rt_ptr, const int32_t *rt_idx, const float *rt_val,
                     int64_t n_items, int k, double alpha, double lambda, int reconstruct_r,
                     int loss_ignores_unspecified, double threshold, double convergence_threshold,
                     int max_iterations, int random_y, const int32_t *test_users, int n_tu,
                     const int32_t *test_items, int n_ti, int n_threads, float *X, float *Y,
                     int *iterations_run, int *apparent_rank, double *last_convergence_value,
                     const uint8_t *present_users, const uint8_t *present_items);

int oracle_als_run(const int64_t *r_ptr, const int32_t *r_idx, const float *r_val, int64_t n_users,
                   const int64_t *rt_ptr, const int32_t *rt_idx, const float *rt_val,
                   int64_t n_items, int k, double alpha, double lambda, int reconstruct_r,
                   int loss_ignores_unspecified, double threshold, double convergence_threshold,
                   int max_iterations, int random_y, const int32_t *test_users, int n_tu,
                   const int32_t *test_items, int n_ti, int n_threads, float *X, float *Y,
                   int *iterations_run, int *apparent_rank, double *last_convergence_value) {
  return oracle_als_run_p(r_ptr, r_idx, r_val, n_users, rt_ptr, rt_idx, rt_val, n_items, k, alpha,
                          lambda, reconstruct_r, loss_ignores_unspecified, threshold,
                          convergence_threshold, max_iterations, random_y, test_users, n_tu,
                          test_items, n_ti, n_threads, X, Y, iterations_run, apparent_rank,
                          last_convergence_value, NULL, NULL);
}

/* Same, with the keys-without-entries masks of RbyRow / RbyColumn (see half_job_t.present). */
int oracle_als_run_p(const int64_t *r_ptr, const int32_t *r_idx, const float *r_val, int64_t n_users,
                     const int64_t *rt_ptr, const int32_t *rt_idx, const float *rt_val,
                     int64_t n_items, int k, double alpha, double lambda, int reconstruct_r,
                     int loss_ignores_unspecified, double threshold, double convergence_threshold,
                     int max_iterations, int random_y, const int32_t *test_users, int n_tu,
                     const int32_t *test_items, int n_ti, int n_threads, float *X, float *Y,
                     int *iterations_run, int *apparent_rank, double *last_convergence_value,
                     const uint8_t *present_users, const uint8_t *present_items) {
  size_t kk = (size_t)k * (size_t)k;
  double *G = (double *)malloc(sizeof(double) * kk);
  double *estimates = (double *)calloc((size_t)n_tu * (size_t)n_ti + 1, sizeof(double));
  /* rows "in the map": Y starts with every row present (Y0 complete); X starts
   * empty and gains exactly the rows of RbyRow with entries. transposeTimesSelf
   * runs over rows in the map; absent X rows are zero here and add nothing. */
  int rc = ORACLE_OK;
  int iterationNumber = 0;
  if (!G || !estimates) { rc = ORACLE_E_OOM; goto out; }
  for (;;) {
    oracle_transpose_times_self(Y, n_items, k, G);
    rc = oracle_als_half_p(r_ptr, r_idx, r_val, n_users, Y, G, k, alpha, lambda, reconstruct_r,
                           loss_ignores_unspecified, threshold, n_threads, X, apparent_rank,
                           present_users);
    if (rc != ORACLE_OK) break;
    oracle_transpose_times_self(X, n_users, k, G);
    rc = oracle_als_half_p(rt_ptr, rt_idx, rt_val, n_items, X, G, k, alpha, lambda, reconstruct_r,
                           loss_ignores_unspecified, threshold, n_threads, Y, apparent_rank,
                           present_items);
    if (rc != ORACLE_OK) break;
    double conv = oracle_convergence_probe(X, Y, k, test_users, n_tu, test_items, n_ti, estimates);
    if (last_convergence_value) *last_convergence_value = conv;
    iterationNumber++;
    if (max_iterations > 0 && iterationNumber >= max_iterations) break;
    if (!isfinite(conv)) break;
    if (!(random_y && iterationNumber == 1) && conv < convergence_threshold) break;
  }
out:
  if (iterations_run) *iterations_run = iterationNumber;
  free(G);
  free(estimates);
  return rc;
}
