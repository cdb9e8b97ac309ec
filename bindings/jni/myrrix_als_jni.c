/*
 * JNI stub between CudaAlternatingLeastSquares.java and the C ABI (include/myrrix_als.h).
 * NOT compiled here (no jni.h in this image).  Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       bindings/jni/myrrix_als_jni.c -o libmyrrix_als_jni.so -L<dir of libmyrrix_als.so> -lmyrrix_als
 * Every function is a 1:1 forward; large arrays are native memory passed by address (no copy here).
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "myrrix_als.h"

#define H(x) ((als_handle *)(intptr_t)(x))
#define CLS(name) Java_net_myrrix_online_factorizer_als_CudaAlternatingLeastSquares_##name

JNIEXPORT jlong JNICALL CLS(nCreate)(JNIEnv *env, jclass c, jint features, jdouble alpha, jdouble lambda,
                                     jboolean reconstructR, jboolean lossIgnoresUnspecified,
                                     jdouble singularityThreshold, jint device) {
  als_config cfg;
  als_handle *h = NULL;
  als_config_default(&cfg);
  cfg.features = features;
  cfg.alpha = alpha;
  cfg.lambda = lambda;
  cfg.reconstruct_r = reconstructR;
  cfg.loss_ignores_unspecified = lossIgnoresUnspecified;
  cfg.singularity_threshold = singularityThreshold;
  cfg.device = device;
  if (als_create(&cfg, &h) != ALS_OK) {
    jclass ex = (*env)->FindClass(env, "java/lang/IllegalStateException");
    (*env)->ThrowNew(env, ex, h ? als_last_error(h) : "als_create failed (no CUDA device?)");
    if (h) als_destroy(h);
    return 0;
  }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL CLS(nDestroy)(JNIEnv *env, jclass c, jlong h) { als_destroy(H(h)); }

/* Large arrays live in native memory owned by the Java side (nAlloc / nFree) and are passed by
 * address: no 2 GiB ByteBuffer limit, no copy in this stub. */
#define P(T, a) ((T *)(intptr_t)(a))

JNIEXPORT jlong JNICALL CLS(nAlloc)(JNIEnv *env, jclass c, jlong bytes) {
  return (jlong)(intptr_t)malloc((size_t)bytes);
}
JNIEXPORT void JNICALL CLS(nFree)(JNIEnv *env, jclass c, jlong address) { free(P(void, address)); }
JNIEXPORT jobject JNICALL CLS(nWindow)(JNIEnv *env, jclass c, jlong address, jlong offset, jint length) {
  return (*env)->NewDirectByteBuffer(env, P(char, address) + offset, (jlong)length);
}

JNIEXPORT jint JNICALL CLS(nSetInteractions)(JNIEnv *env, jclass c, jlong h, jlong nUsers, jlong nItems,
                                             jlong rowPtr, jlong colIdx, jlong val) {
  return als_set_interactions(H(h), nUsers, nItems, P(const int64_t, rowPtr), P(const int32_t, colIdx),
                              P(const float, val));
}

JNIEXPORT jint JNICALL CLS(nSetInteractionsByColumn)(JNIEnv *env, jclass c, jlong h, jlong colPtr,
                                                     jlong rowIdx, jlong val) {
  return als_set_interactions_by_column(H(h), P(const int64_t, colPtr), P(const int32_t, rowIdx),
                                        P(const float, val));
}

JNIEXPORT jint JNICALL CLS(nSetPresentEmptyRows)(JNIEnv *env, jclass c, jlong h, jint which, jintArray rows) {
  jint n = (*env)->GetArrayLength(env, rows);
  jint *r = (*env)->GetIntArrayElements(env, rows, NULL);
  int rc = als_set_present_empty_rows(H(h), which, (const int32_t *)r, n);
  (*env)->ReleaseIntArrayElements(env, rows, r, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nSetY)(JNIEnv *env, jclass c, jlong h, jlong y) {
  return als_set_y(H(h), P(const float, y));
}

JNIEXPORT jint JNICALL CLS(nHalfX)(JNIEnv *env, jclass c, jlong h) { return als_half_x(H(h)); }
JNIEXPORT jint JNICALL CLS(nHalfY)(JNIEnv *env, jclass c, jlong h) { return als_half_y(H(h)); }
JNIEXPORT jint JNICALL CLS(nSync)(JNIEnv *env, jclass c, jlong h) { return als_sync(H(h)); }

JNIEXPORT jint JNICALL CLS(nProbe)(JNIEnv *env, jclass c, jlong h, jintArray users, jintArray items,
                                   jdoubleArray out) {
  jint nu = (*env)->GetArrayLength(env, users), ni = (*env)->GetArrayLength(env, items);
  jint *u = (*env)->GetIntArrayElements(env, users, NULL);
  jint *i = (*env)->GetIntArrayElements(env, items, NULL);
  jdouble *o = (*env)->GetDoubleArrayElements(env, out, NULL);
  int rc = als_probe(H(h), (const int32_t *)u, nu, (const int32_t *)i, ni, o);
  (*env)->ReleaseIntArrayElements(env, users, u, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, items, i, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, out, o, 0);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nCall)(JNIEnv *env, jclass c, jlong h, jintArray users, jintArray items,
                                  jint maxIterations, jdouble threshold, jboolean randomY, jboolean xIsEmpty,
                                  jintArray iterationsRun, jdoubleArray lastValue) {
  jint nu = (*env)->GetArrayLength(env, users), ni = (*env)->GetArrayLength(env, items);
  jint *u = (*env)->GetIntArrayElements(env, users, NULL);
  jint *i = (*env)->GetIntArrayElements(env, items, NULL);
  int32_t it = 0;
  double value = 0.0;
  int rc = als_call(H(h), (const int32_t *)u, nu, (const int32_t *)i, ni, maxIterations, threshold,
                    randomY ? 1 : 0, xIsEmpty ? 1 : 0, &it, &value);
  (*env)->ReleaseIntArrayElements(env, users, u, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, items, i, JNI_ABORT);
  jint jit = it;
  (*env)->SetIntArrayRegion(env, iterationsRun, 0, 1, &jit);
  (*env)->SetDoubleArrayRegion(env, lastValue, 0, 1, &value);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nGetX)(JNIEnv *env, jclass c, jlong h, jlong out) {
  return als_get_x(H(h), P(float, out));
}
JNIEXPORT jint JNICALL CLS(nGetY)(JNIEnv *env, jclass c, jlong h, jlong out) {
  return als_get_y(H(h), P(float, out));
}
JNIEXPORT jstring JNICALL CLS(nLastError)(JNIEnv *env, jclass c, jlong h) {
  return (*env)->NewStringUTF(env, als_last_error(H(h)));
}
JNIEXPORT jint JNICALL CLS(nSingularRank)(JNIEnv *env, jclass c, jlong h) { return als_singular_rank(H(h)); }
