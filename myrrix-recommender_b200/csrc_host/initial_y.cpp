// initial_y.cpp -- libmyrrix_init.so (include/myrrix_init.h): MersenneTwister stream, random unit
// vectors, constructInitialY.  Citations in the header.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/myrrix_init.h"

namespace {
constexpr int N = 624, M = 397;
constexpr int kMaxFarFrom = 100000;  // MAX_FAR_FROM_VECTORS (AlternatingLeastSquares.java:81)
}  // namespace

struct myrrix_rng {
  uint32_t mt[N];
  int mti;
  double next_gaussian;  // cached second value of the pair; NaN: none
};

namespace {

void seed_int(myrrix_rng* r, uint32_t seed) {  // MersenneTwister.setSeed(int)
  r->mt[0] = seed;
  for (int i = 1; i < N; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = N;
  r->next_gaussian = NAN;
}

void seed_array(myrrix_rng* r, const uint32_t* key, int len) {  // MersenneTwister.setSeed(int[])
  seed_int(r, 19650218u);
  int i = 1, j = 0;
  for (int k = (N > len ? N : len); k != 0; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    i++;
    j++;
    if (i >= N) { r->mt[0] = r->mt[N - 1]; i = 1; }
    if (j >= len) j = 0;
  }
  for (int k = N - 1; k != 0; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= N) { r->mt[0] = r->mt[N - 1]; i = 1; }
  }
  r->mt[0] = 0x80000000u;
  r->mti = N;
  r->next_gaussian = NAN;
}

uint32_t next32(myrrix_rng* r) {
  if (r->mti >= N) {
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    int kk = 0;
    for (; kk < N - M; kk++) {
      const uint32_t y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < N - 1; kk++) {
      const uint32_t y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
      r->mt[kk] = r->mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    const uint32_t y = (r->mt[N - 1] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
    r->mt[N - 1] = r->mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    r->mti = 0;
  }
  uint32_t y = r->mt[r->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

inline uint32_t next_bits(myrrix_rng* r, int bits) { return next32(r) >> (32 - bits); }

double next_double(myrrix_rng* r) {  // BitsStreamGenerator.nextDouble
  const uint64_t high = (uint64_t)next_bits(r, 26) << 26;
  const uint32_t low = next_bits(r, 26);
  return (double)(high | low) * 0x1.0p-52;
}

double next_gaussian(myrrix_rng* r) {  // BitsStreamGenerator.nextGaussian
  if (isnan(r->next_gaussian)) {
    const double x = next_double(r);
    const double y = next_double(r);
    const double alpha = 2 * M_PI * x;
    const double rad = sqrt(-2 * log(y));
    const double random = rad * cos(alpha);
    r->next_gaussian = rad * sin(alpha);
    return random;
  }
  const double random = r->next_gaussian;
  r->next_gaussian = NAN;
  return random;
}

int32_t next_int(myrrix_rng* r, int32_t n) {  // BitsStreamGenerator.nextInt(int)
  if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next_bits(r, 31)) >> 31);
  int32_t bits, val;
  do {
    bits = (int32_t)next_bits(r, 31);
    val = bits % n;
  } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);  // Java int overflow wraps
  return val;
}

void do_random_unit_vector(myrrix_rng* r, float* v, int dims) {  // RandomUtils.java:88-100
  double total = 0.0;
  for (int i = 0; i < dims; i++) {
    const double d = next_gaussian(r);
    v[i] = (float)d;
    total += d * d;
  }
  const float normalization = (float)sqrt(total);
  for (int i = 0; i < dims; i++) v[i] /= normalization;
}

double dot(const float* x, const float* y, int n) {  // SimpleVectorMath.dot (fp32 products, fp64 sum)
  double d = 0.0;
  for (int i = 0; i < n; i++) {
    const float p = x[i] * y[i];
    d += (double)p;
  }
  return d;
}

void normalize(float* x, int n) {  // SimpleVectorMath.normalize (SimpleVectorMath.java:81-86)
  double total = 0.0;
  for (int i = 0; i < n; i++) {
    const float p = x[i] * x[i];
    total += (double)p;
  }
  const float norm = (float)sqrt(total);
  for (int i = 0; i < n; i++) x[i] /= norm;
}

// far_from: pointers to the vectors, in list order
void far_from_ptrs(myrrix_rng* r, int dims, const std::vector<const float*>& far, float* out) {
  const int size = (int)far.size();
  const int num_samples = size < 100 ? size : 100;
  bool accepted = false;
  while (!accepted) {
    do_random_unit_vector(r, out, dims);
    double smallest = INFINITY;
    for (int s = 0; s < num_samples; s++) {
      const float* other = far[size == num_samples ? s : next_int(r, size)];
      const double dist2 = 2.0 - 2.0 * dot(out, other, dims);
      if (isfinite(dist2) && dist2 < smallest) smallest = dist2;
    }
    if (isfinite(smallest) && !(dims == 1 && smallest == 0.0)) {
      const double accept_probability = smallest / 4.0;
      accepted = next_double(r) < accept_probability;
    } else {
      accepted = true;
    }
  }
}

}  // namespace

extern "C" {

myrrix_rng* myrrix_rng_create(int64_t seed) {
  myrrix_rng* r = new (std::nothrow) myrrix_rng;
  if (!r) return nullptr;
  const uint32_t key[2] = {(uint32_t)((uint64_t)seed >> 32), (uint32_t)((uint64_t)seed & 0xffffffffu)};
  seed_array(r, key, 2);  // MersenneTwister.setSeed(long)
  return r;
}

myrrix_rng* myrrix_rng_create_by_array(const int32_t* key, int32_t n) {
  if (!key || n <= 0) return nullptr;
  myrrix_rng* r = new (std::nothrow) myrrix_rng;
  if (!r) return nullptr;
  seed_array(r, reinterpret_cast<const uint32_t*>(key), n);
  return r;
}

void myrrix_rng_destroy(myrrix_rng* r) { delete r; }

uint32_t myrrix_rng_next_bits(myrrix_rng* r, int32_t bits) { return (bits < 1 || bits > 32) ? 0u : next_bits(r, bits); }
double myrrix_rng_next_double(myrrix_rng* r) { return next_double(r); }
double myrrix_rng_next_gaussian(myrrix_rng* r) { return next_gaussian(r); }
int32_t myrrix_rng_next_int(myrrix_rng* r, int32_t n) { return n > 0 ? next_int(r, n) : -1; }

int myrrix_random_unit_vector(myrrix_rng* r, int32_t dimensions, float* out) {
  if (!r || !out || dimensions <= 0) return MYRRIX_INIT_E_ARG;
  do_random_unit_vector(r, out, dimensions);
  return MYRRIX_INIT_OK;
}

int myrrix_random_unit_vector_far_from(myrrix_rng* r, int32_t dimensions, const float* far_from, int64_t n_far,
                                       float* out) {
  if (!r || !out || dimensions <= 0 || n_far < 0 || (n_far > 0 && !far_from) || n_far > 0x7fffffff)
    return MYRRIX_INIT_E_ARG;
  std::vector<const float*> far((size_t)n_far);
  for (int64_t i = 0; i < n_far; i++) far[(size_t)i] = far_from + i * dimensions;
  far_from_ptrs(r, dimensions, far, out);
  return MYRRIX_INIT_OK;
}

int myrrix_construct_initial_y(myrrix_rng* r, int32_t features, int64_t n_rows, int32_t prev_features,
                               const float* prev, const int64_t* prev_order, int64_t n_prev,
                               const int64_t* column_order, int64_t n_columns, float* y, uint8_t* has_vector_out) {
  if (!r || !y || features <= 0 || n_rows < 0 || n_prev < 0 || n_columns < 0 || prev_features < 0) return MYRRIX_INIT_E_ARG;
  if (n_prev > 0 && (!prev || !prev_order || prev_features <= 0)) return MYRRIX_INIT_E_ARG;
  if (n_columns > 0 && !column_order) return MYRRIX_INIT_E_ARG;
  std::vector<uint8_t> has;
  try {
    has.assign((size_t)n_rows, 0);
  } catch (...) {
    return MYRRIX_INIT_E_OOM;
  }
  for (int64_t i = 0; i < n_prev; i++)
    if (prev_order[i] < 0 || prev_order[i] >= n_rows) return MYRRIX_INIT_E_ARG;
  for (int64_t i = 0; i < n_columns; i++)
    if (column_order[i] < 0 || column_order[i] >= n_rows) return MYRRIX_INIT_E_ARG;
  // previousY -> randomY (:268-309), in the map's iteration order
  for (int64_t i = 0; i < n_prev; i++) {
    const int64_t row = prev_order[i];
    const float* old = prev + row * prev_features;
    float* v = y + row * features;
    if (prev_features > features) {          // project down + normalise (:277-287)
      memcpy(v, old, sizeof(float) * (size_t)features);
      normalize(v, features);
    } else if (prev_features < features) {   // subspace + N(0,1) + normalise (:289-302)
      memcpy(v, old, sizeof(float) * (size_t)prev_features);
      for (int f = prev_features; f < features; f++) v[f] = (float)next_gaussian(r);
      normalize(v, features);
    } else {                                 // adopted as it is (:304-308)
      memcpy(v, old, sizeof(float) * (size_t)features);
    }
    has[(size_t)row] = 1;
  }
  // recentVectors: the first MAX_FAR_FROM_VECTORS entries of randomY (:311-317)
  std::vector<const float*> recent;
  try {
    recent.reserve((size_t)((n_prev + n_columns) < kMaxFarFrom ? (n_prev + n_columns) : kMaxFarFrom));
  } catch (...) {
    return MYRRIX_INIT_E_OOM;
  }
  for (int64_t i = 0; i < n_prev && (int)recent.size() < kMaxFarFrom; i++) recent.push_back(y + prev_order[i] * features);
  // every key of RbyColumn without a vector (:318-332)
  for (int64_t i = 0; i < n_columns; i++) {
    const int64_t row = column_order[i];
    if (has[(size_t)row]) continue;
    float* v = y + row * features;
    far_from_ptrs(r, features, recent, v);
    has[(size_t)row] = 1;
    if ((int)recent.size() < kMaxFarFrom) recent.push_back(v);
  }
  if (has_vector_out) memcpy(has_vector_out, has.data(), (size_t)n_rows);
  return MYRRIX_INIT_OK;
}

}  // extern "C"
