/*
 * myrrix_init.h -- C ABI of the cold-start path: constructInitialY and the random stream behind it
 * (SURVEY.md 8a A6, 8f N4).  Host C++ (libmyrrix_init.so, plain g++; loads without CUDA): the
 * rejection sampling is a serial, order-dependent walk of ONE random stream -- O(items * 100 * k)
 * flops, once per build -- so it stays on the host like in the reference and hands a complete Y0
 * to als_set_y.
 *
 * Replaces
 *   AlternatingLeastSquares.constructInitialY  (online/src/net/myrrix/online/factorizer/als/
 *       AlternatingLeastSquares.java:264-335)
 *   RandomUtils.randomUnitVector / randomUnitVectorFarFrom  (common/src/net/myrrix/common/random/
 *       RandomUtils.java:82-140)
 *   RandomManager.getRandom  (common/.../random/RandomManager.java:41-54: new MersenneTwister(),
 *       or MersenneTwister(1234567890L) under useTestSeed)
 * and, from commons-math3 3.2 (absent from /root/reference, pom.xml:81; restated from the published
 * algorithm): MersenneTwister(long) = MT19937 init_by_array({high 32 bits, low 32 bits}),
 * BitsStreamGenerator.nextDouble ((next(26) << 26 | next(26)) * 2^-52), nextGaussian (Box-Muller
 * pair: r cos(2 pi x), r sin(2 pi x), r = sqrt(-2 log y), the second value cached), nextInt(n)
 * (power of two: (n * next(31)) >> 31; else rejection on bits - val + (n - 1) < 0).
 * log / cos / sin are libm's here and FastMath's there (both < 1 ulp): the stream of raw 32-bit
 * words is bit-exact (pinned against the published MT19937 vectors), a Gaussian may differ in its
 * last fp64 bit before it is narrowed to fp32.
 */
#ifndef MYRRIX_INIT_H_
#define MYRRIX_INIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct myrrix_rng myrrix_rng;

enum myrrix_init_status { MYRRIX_INIT_OK = 0, MYRRIX_INIT_E_ARG = 1, MYRRIX_INIT_E_OOM = 2 };

/* new MersenneTwister(seed) */
myrrix_rng* myrrix_rng_create(int64_t seed);
/* new MersenneTwister(int[] key): MT19937 init_by_array */
myrrix_rng* myrrix_rng_create_by_array(const int32_t* key, int32_t n);
void myrrix_rng_destroy(myrrix_rng* r);
/* RandomGenerator.next(bits) of the raw stream, nextDouble, nextGaussian, nextInt(n) (n > 0) */
uint32_t myrrix_rng_next_bits(myrrix_rng* r, int32_t bits);
double myrrix_rng_next_double(myrrix_rng* r);
double myrrix_rng_next_gaussian(myrrix_rng* r);
int32_t myrrix_rng_next_int(myrrix_rng* r, int32_t n);

/* RandomUtils.randomUnitVector (RandomUtils.java:82-100) */
int myrrix_random_unit_vector(myrrix_rng* r, int32_t dimensions, float* out);
/* RandomUtils.randomUnitVectorFarFrom (:110-140); far_from: [n_far][dimensions] row-major */
int myrrix_random_unit_vector_far_from(myrrix_rng* r, int32_t dimensions, const float* far_from,
                                       int64_t n_far, float* out);

/*
 * constructInitialY on dense item rows.  y: [n_rows][features] row-major, in/out.
 *   prev_features            feature count of previousY (0: none / empty -> "start from scratch")
 *   prev                     [n_rows][prev_features]: previousY's vector of every row that has one
 *   prev_order[n_prev]       dense rows of previousY's entries in the order the caller's map
 *                            iterates them (FastByIDMap slot order; drives the random stream when
 *                            the feature count grew, :289-302, and the far-from list, :311-317)
 *   column_order[n_columns]  dense rows of RbyColumn's keys in keySetIterator() order (:318-332)
 * On return y holds: previous vectors as they are (same feature count, :304-308), truncated and
 * normalised (:277-287) or padded with N(0,1) and normalised (:289-302); every row of column_order
 * without a previous vector gets randomUnitVectorFarFrom(the first <= 100000 vectors, :311-327).
 * has_vector_out[n_rows] (may be NULL): 1 for rows that now have a vector.
 */
int myrrix_construct_initial_y(myrrix_rng* r, int32_t features, int64_t n_rows, int32_t prev_features,
                               const float* prev, const int64_t* prev_order, int64_t n_prev,
                               const int64_t* column_order, int64_t n_columns, float* y,
                               uint8_t* has_vector_out);

#ifdef __cplusplus
}
#endif
#endif /* MYRRIX_INIT_H_ */
