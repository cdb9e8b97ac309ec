/*
 * A plain C caller of include/myrrix_als.h, LINKED against libmyrrix_als.so (no ctypes, no
 * Python): the reference's own golden test AlternatingLeastSquaresTest.testALS
 * (online/test/net/myrrix/online/factorizer/als/AlternatingLeastSquaresTest.java:39-57) driven the
 * way the JNI shim drives the library -- set both orientations, set Y0, iterate with the stop rule
 * of AlternatingLeastSquares.call (:227-257) on als_probe, read X and Y back, compare X Y^T with
 * the 25 pinned values.  golden_data.h is generated from tests/golden/reference_goldens.json by
 * tests/test_c_link.py.
 *
 * exit 0 = parity within 1e-4 (the north_star bar); 77 = no CUDA device (als_create said so: the
 * library has no CPU fallback); anything else = failure.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "golden_data.h" /* N_USERS, N_ITEMS, FEATURES, THRESHOLD, MAX_ITER, R[][], Y0[][], PRODUCT[][] */
#include "myrrix_als.h"

#define CHECK(expr)                                                                   \
  do {                                                                                \
    int rc_ = (expr);                                                                 \
    if (rc_ != ALS_OK) {                                                              \
      fprintf(stderr, "%s -> %d (%s)\n", #expr, rc_, h ? als_last_error(h) : "?");    \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(void) {
  als_config cfg;
  als_handle *h = NULL;
  als_config_default(&cfg);
  cfg.features = FEATURES;
  int rc = als_create(&cfg, &h);
  if (rc == ALS_E_CUDA) {
    printf("no CUDA device: als_create returned ALS_E_CUDA (no CPU fallback)\n");
    if (h) als_destroy(h);
    return 77;
  }
  if (rc != ALS_OK) {
    fprintf(stderr, "als_create -> %d\n", rc);
    return 1;
  }
  /* dense R -> CSR by user and by item (MatrixUtils.addTo order, AlternatingLeastSquaresTest.java:92-105) */
  int64_t rp[N_USERS + 1], cp[N_ITEMS + 1];
  int32_t ci[N_USERS * N_ITEMS], ri[N_USERS * N_ITEMS];
  float rv[N_USERS * N_ITEMS], cv[N_USERS * N_ITEMS];
  int64_t e = 0;
  for (int u = 0; u < N_USERS; u++) {
    rp[u] = e;
    for (int i = 0; i < N_ITEMS; i++)
      if (R[u][i] != 0.0f) { ci[e] = i; rv[e] = R[u][i]; e++; }
  }
  rp[N_USERS] = e;
  e = 0;
  for (int i = 0; i < N_ITEMS; i++) {
    cp[i] = e;
    for (int u = 0; u < N_USERS; u++)
      if (R[u][i] != 0.0f) { ri[e] = u; cv[e] = R[u][i]; e++; }
  }
  cp[N_ITEMS] = e;
  CHECK(als_set_interactions(h, N_USERS, N_ITEMS, rp, ci, rv));
  CHECK(als_set_interactions_by_column(h, cp, ri, cv));
  CHECK(als_set_y(h, &Y0[0][0]));

  int32_t users[N_USERS], items[N_ITEMS];
  for (int u = 0; u < N_USERS; u++) users[u] = u;
  for (int i = 0; i < N_ITEMS; i++) items[i] = i;
  double est[N_USERS * N_ITEMS] = {0}, fresh[N_USERS * N_ITEMS];
  int it = 0;
  for (;;) {
    CHECK(als_half_x(h));
    CHECK(als_half_y(h));
    CHECK(als_sync(h));
    CHECK(als_probe(h, users, N_USERS, items, N_ITEMS, fresh));
    /* DoubleWeightedMean.increment (common/.../stats/DoubleWeightedMean.java:73-81) */
    double total = 0.0, mean = NAN;
    for (int j = 0; j < N_USERS * N_ITEMS; j++) {
      double datum = fabs(fresh[j] - est[j]), w = fresh[j] > 0.0 ? fresh[j] : 0.0, old = total;
      total += w;
      mean = (old <= 0.0) ? datum : mean * old / total + datum * w / total;
      est[j] = fresh[j];
    }
    it++;
    if (MAX_ITER > 0 && it >= MAX_ITER) break;
    if (!isfinite(mean)) break;
    if (mean < THRESHOLD) break; /* previousY given: randomY == false, ALS.java:253 */
  }
  float X[N_USERS][FEATURES], Y[N_ITEMS][FEATURES];
  CHECK(als_get_x(h, &X[0][0]));
  CHECK(als_get_y(h, &Y[0][0]));
  CHECK(als_destroy(h));
  double worst = 0.0, scale = 0.0;
  for (int u = 0; u < N_USERS; u++)
    for (int i = 0; i < N_ITEMS; i++) {
      double p = 0.0;
      for (int f = 0; f < FEATURES; f++) p += (double)X[u][f] * (double)Y[i][f];
      if (fabs(p - PRODUCT[u][i]) > worst) worst = fabs(p - PRODUCT[u][i]);
      if (fabs(PRODUCT[u][i]) > scale) scale = fabs(PRODUCT[u][i]);
    }
  printf("iterations %d, max |X Y^T - golden| / max |golden| = %.3g\n", it, worst / scale);
  return (worst <= 1e-4 * scale && abs(it - 28) <= 1) ? 0 : 2;
}
