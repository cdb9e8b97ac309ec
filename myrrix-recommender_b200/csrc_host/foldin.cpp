// foldin.cpp -- the fold-in math behind include/myrrix_foldin.h (host C++, fp64).
//
// One online write costs two k x k triangular solves against factorisations computed once
// per generation, so this stays on the host: the GPU's part is the two Gramians
// (als_gramian), which are k x k reductions over the factors already resident in HBM.
#include "../../include/myrrix_foldin.h"

#include <math.h>
#include <string.h>

#include <new>
#include <vector>

namespace {

// RRQRDecomposition(M, threshold) of commons-math3 3.2 for a square k x k matrix: Householder
// reflections on the transposed copy, before each step the remaining column of largest norm is
// swapped in; solve = apply the reflectors to b, back-substitute against R, undo the permutation.
struct Rrqr {
  int k = 0;
  std::vector<double> qrt;    // [k][k]: column `minor` of the matrix is row `minor` here
  std::vector<double> rdiag;  // R's diagonal
  std::vector<int> perm;      // perm[j] = original column now at position j
  double threshold = 0;

  void decompose(const double* M, int k_, double thr) {
    k = k_;
    threshold = thr;
    qrt.assign((size_t)k * k, 0.0);
    rdiag.assign(k, 0.0);
    perm.resize(k);
    for (int i = 0; i < k; i++) {
      perm[i] = i;
      for (int j = 0; j < k; j++) qrt[(size_t)j * k + i] = M[(size_t)i * k + j];  // transpose
    }
    for (int minor = 0; minor < k; minor++) {
      // pivot: remaining column with the largest squared norm -- taken over the WHOLE stored
      // column, reflected part included, as RRQRDecomposition 3.2 does
      double best = 0;
      int piv = minor;
      for (int c = minor; c < k; c++) {
        double n2 = 0;
        for (int r = 0; r < k; r++) n2 += qrt[(size_t)c * k + r] * qrt[(size_t)c * k + r];
        if (n2 > best) { best = n2; piv = c; }
      }
      if (piv != minor) {
        for (int r = 0; r < k; r++) std::swap(qrt[(size_t)minor * k + r], qrt[(size_t)piv * k + r]);
        std::swap(perm[minor], perm[piv]);
      }
      double* col = &qrt[(size_t)minor * k];
      double x2 = 0;
      for (int r = minor; r < k; r++) x2 += col[r] * col[r];
      const double a = (col[minor] > 0) ? -sqrt(x2) : sqrt(x2);
      rdiag[minor] = a;
      if (a != 0.0) {
        col[minor] -= a;  // v = x - a e
        for (int c = minor + 1; c < k; c++) {
          double* other = &qrt[(size_t)c * k];
          double alpha = 0;
          for (int r = minor; r < k; r++) alpha -= other[r] * col[r];
          alpha /= a * col[minor];
          for (int r = minor; r < k; r++) other[r] -= alpha * col[r];
        }
      }
    }
  }
  bool nonsingular() const {
    for (int i = 0; i < k; i++)
      if (fabs(rdiag[i]) <= threshold) return false;
    return true;
  }
  // getRank(dropThreshold): ||R_rr|| of the trailing blocks relative to the previous one
  int rank(double drop) const {
    auto frob_tail = [&](int from) {  // Frobenius norm of R[from:, from:]
      double s = 0;
      for (int i = from; i < k; i++) {
        s += rdiag[i] * rdiag[i];
        for (int c = i + 1; c < k; c++) s += qrt[(size_t)c * k + i] * qrt[(size_t)c * k + i];
      }
      return sqrt(s);
    };
    int r = 1;
    double last = frob_tail(0), rn = last;
    while (r < k) {
      const double next = frob_tail(r);
      if (next == 0 || (next / last) * rn < drop) break;
      last = next;
      r++;
    }
    return r;
  }
  // y, z: caller-provided scratch of k doubles each
  void solve(const double* b, double* x, double* y, double* z) const {
    memcpy(y, b, sizeof(double) * k);
    for (int minor = 0; minor < k; minor++) {  // y = Q' b
      const double* col = &qrt[(size_t)minor * k];
      double dot = 0;
      for (int r = minor; r < k; r++) dot += y[r] * col[r];
      dot /= rdiag[minor] * col[minor];
      for (int r = minor; r < k; r++) y[r] += dot * col[r];
    }
    for (int row = k - 1; row >= 0; row--) {  // R z = y
      y[row] /= rdiag[row];
      const double yr = y[row];
      z[row] = yr;
      const double* col = &qrt[(size_t)row * k];
      for (int i = 0; i < row; i++) y[i] -= yr * col[i];
    }
    for (int j = 0; j < k; j++) x[perm[j]] = z[j];  // undo the column permutation
  }
};

}  // namespace

struct foldin_handle {
  int k = 0;
  double learn_rate = 1.0;
  bool have[2] = {false, false};
  Rrqr solver[2];
};

namespace {

// recomputeSolver (Generation.java:141-158)
int build_solver(const double* M, int k, double threshold, Rrqr* out, int* apparent_rank) {
  double inf_norm = 0;  // RealMatrix.getNorm(): maximum absolute row sum
  for (int i = 0; i < k; i++) {
    double s = 0;
    for (int j = 0; j < k; j++) {
      if (!isfinite(M[(size_t)i * k + j])) return FOLDIN_E_NONFINITE;
      s += fabs(M[(size_t)i * k + j]);
    }
    if (s > inf_norm) inf_norm = s;
  }
  if (inf_norm < 1.0) return FOLDIN_E_ILL_CONDITIONED;
  out->decompose(M, k, threshold);
  if (!out->nonsingular()) {
    if (apparent_rank) *apparent_rank = out->rank(0.01);
    return FOLDIN_E_SINGULAR;
  }
  return FOLDIN_OK;
}

double fold_in_weight(double learn_rate, double estimate, float value) {
  double w;
  if (value > 0.0f && estimate < 1.0) {
    const double multiplier = 1.0 - fmax(0.0, estimate);
    w = (1.0 - 1.0 / (1.0 + (double)value)) * multiplier;
  } else if (value < 0.0f && estimate > 0.0) {
    const double multiplier = -fmin(1.0, estimate);
    w = (1.0 - 1.0 / (1.0 - (double)value)) * multiplier;
  } else {
    w = 0.0;
  }
  return learn_rate * w;
}

}  // namespace

extern "C" {

int foldin_create(int32_t features, const double* xtx, const double* yty, double singularity_threshold,
                  double learn_rate, foldin_handle** out, int32_t* which_failed, int32_t* apparent_rank) {
  if (!out || features <= 0 || !(singularity_threshold >= 0)) return FOLDIN_E_ARG;
  foldin_handle* h = new (std::nothrow) foldin_handle();
  if (!h) return FOLDIN_E_OOM;
  h->k = features;
  h->learn_rate = learn_rate;
  const double* M[2] = {xtx, yty};
  for (int w = 0; w < 2; w++) {
    if (!M[w]) continue;
    int rank = 0;
    const int rc = build_solver(M[w], features, singularity_threshold, &h->solver[w], &rank);
    if (rc != FOLDIN_OK) {
      if (which_failed) *which_failed = w;
      if (apparent_rank) *apparent_rank = rank;
      delete h;
      return rc;
    }
    h->have[w] = true;
  }
  *out = h;
  return FOLDIN_OK;
}

void foldin_destroy(foldin_handle* h) { delete h; }

double foldin_weight(const foldin_handle* h, double estimate, float value) {
  return fold_in_weight(h ? h->learn_rate : 1.0, estimate, value);
}

int foldin_solve(const foldin_handle* h, int32_t which, const float* b, double* x) {
  if (!h || !b || !x || (which != 0 && which != 1)) return FOLDIN_E_ARG;
  if (!h->have[which]) return FOLDIN_E_NOT_READY;
  thread_local std::vector<double> scratch;  // no allocation per online write
  if (scratch.size() < 3 * (size_t)h->k) scratch.resize(3 * (size_t)h->k);
  double* bd = scratch.data();
  for (int i = 0; i < h->k; i++) bd[i] = (double)b[i];
  h->solver[which].solve(bd, x, bd + h->k, bd + 2 * h->k);
  return FOLDIN_OK;
}

int foldin_update_features(const foldin_handle* h, float* user, float* item, float value) {
  if (!h) return FOLDIN_E_ARG;
  if (!user || !item) return FOLDIN_OK;  // (:866-868)
  const int k = h->k;
  double est = 0;  // SimpleVectorMath.dot: fp32 products, fp64 sum (SimpleVectorMath.java:34-41)
  for (int i = 0; i < k; i++) est += (double)(user[i] * item[i]);
  if (!isfinite(est)) return FOLDIN_E_NONFINITE;
  const double w = fold_in_weight(h->learn_rate, est, value);
  if (w == 0.0) return FOLDIN_OK;
  thread_local std::vector<double> folds;
  if (folds.size() < 2 * (size_t)k) folds.resize(2 * (size_t)k);
  double* item_fold = folds.data();
  double* user_fold = folds.data() + k;
  // both solves read the rows as they were on entry (:876-884)
  if (h->have[0]) foldin_solve(h, 0, user, item_fold);
  if (h->have[1]) foldin_solve(h, 1, item, user_fold);
  if (h->have[0]) {
    for (int i = 0; i < k; i++) {
      const double delta = w * item_fold[i];
      if (!isfinite(delta)) return FOLDIN_E_NONFINITE;
      item[i] += (float)delta;
    }
  }
  if (h->have[1]) {
    for (int i = 0; i < k; i++) {
      const double delta = w * user_fold[i];
      if (!isfinite(delta)) return FOLDIN_E_NONFINITE;
      user[i] += (float)delta;
    }
  }
  return FOLDIN_OK;
}

int foldin_update_many(const foldin_handle* h, float* X, float* Y, const int32_t* users,
                       const int32_t* items, const float* values, int64_t n) {
  if (!h || !X || !Y || !users || !items || n < 0) return FOLDIN_E_ARG;
  for (int64_t e = 0; e < n; e++) {
    const int rc = foldin_update_features(h, X + (size_t)users[e] * h->k, Y + (size_t)items[e] * h->k,
                                          values ? values[e] : 1.0f);
    if (rc != FOLDIN_OK) return rc;
  }
  return FOLDIN_OK;
}

int foldin_export_solver(const foldin_handle* h, int32_t which, double* qrt, double* rdiag, int32_t* perm) {
  if (!h || !qrt || !rdiag || !perm || (which != 0 && which != 1)) return FOLDIN_E_ARG;
  if (!h->have[which]) return FOLDIN_E_NOT_READY;
  const Rrqr& s = h->solver[which];
  memcpy(qrt, s.qrt.data(), sizeof(double) * (size_t)h->k * h->k);
  memcpy(rdiag, s.rdiag.data(), sizeof(double) * (size_t)h->k);
  for (int i = 0; i < h->k; i++) perm[i] = s.perm[i];
  return FOLDIN_OK;
}

int foldin_anonymous_user(const foldin_handle* h, const float* item_rows, const float* values, int32_t n,
                          float* out) {
  if (!h || !out || n < 0 || (n > 0 && !item_rows)) return FOLDIN_E_ARG;
  if (!h->have[1]) return FOLDIN_E_NOT_READY;  // (:570-573)
  const int k = h->k;
  memset(out, 0, sizeof(float) * k);
  std::vector<double> fold(k);
  for (int j = 0; j < n; j++) {
    foldin_solve(h, 1, item_rows + (size_t)j * k, fold.data());
    const double w = fold_in_weight(h->learn_rate, 0.0, values ? values[j] : 1.0f);  // (:596)
    if (w != 0.0)
      for (int i = 0; i < k; i++) out[i] += (float)(w * fold[i]);
  }
  return FOLDIN_OK;
}

}  // extern "C"
