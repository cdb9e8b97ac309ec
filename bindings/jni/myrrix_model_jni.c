/*
 * JNI stub between CudaResidentModel.java and the resident-model entry points of the C ABI
 * (include/myrrix_als.h).  NOT compiled here (no jni.h in this image); build like myrrix_als_jni.c.
 * Every function is a 1:1 forward.
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "myrrix_als.h"

#define H(x) ((als_handle *)(intptr_t)(x))
#define CLS(name) Java_net_myrrix_online_CudaResidentModel_##name

static jint *ints(JNIEnv *env, jintArray a, jint *n) {
  *n = a ? (*env)->GetArrayLength(env, a) : 0;
  return *n ? (*env)->GetIntArrayElements(env, a, NULL) : NULL;
}
static void done_ints(JNIEnv *env, jintArray a, jint *p, jint mode) {
  if (p) (*env)->ReleaseIntArrayElements(env, a, p, mode);
}

JNIEXPORT jint JNICALL CLS(nRecommend)(JNIEnv *env, jclass c, jlong h, jintArray users, jint howMany,
                                       jboolean considerKnownItems, jintArray exclude, jintArray outItems,
                                       jfloatArray outValues, jintArray outCount) {
  jint nu, ne, no, nc;
  jint *u = ints(env, users, &nu), *e = ints(env, exclude, &ne);
  jint *oi = ints(env, outItems, &no), *oc = ints(env, outCount, &nc);
  jfloat *ov = (*env)->GetFloatArrayElements(env, outValues, NULL);
  int rc = als_recommend(H(h), (const int32_t *)u, nu, howMany, considerKnownItems ? 1 : 0, (const int32_t *)e, ne,
                         (int32_t *)oi, ov, (int32_t *)oc);
  done_ints(env, users, u, JNI_ABORT);
  done_ints(env, exclude, e, JNI_ABORT);
  done_ints(env, outItems, oi, 0);
  done_ints(env, outCount, oc, 0);
  (*env)->ReleaseFloatArrayElements(env, outValues, ov, 0);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nRecommendBatch)(JNIEnv *env, jclass c, jlong h, jintArray users, jint howMany,
                                            jboolean considerKnownItems, jintArray outItems,
                                            jfloatArray outValues, jintArray outCounts) {
  jint nu, no, nc;
  jint *u = ints(env, users, &nu), *oi = ints(env, outItems, &no), *oc = ints(env, outCounts, &nc);
  jfloat *ov = (*env)->GetFloatArrayElements(env, outValues, NULL);
  int rc = als_recommend_batch(H(h), (const int32_t *)u, nu, howMany, considerKnownItems ? 1 : 0, (int32_t *)oi, ov,
                               (int32_t *)oc);
  done_ints(env, users, u, JNI_ABORT);
  done_ints(env, outItems, oi, 0);
  done_ints(env, outCounts, oc, 0);
  (*env)->ReleaseFloatArrayElements(env, outValues, ov, 0);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nTopN)(JNIEnv *env, jclass c, jlong h, jint which, jfloatArray features, jint nVectors,
                                  jintArray exclude, jint howMany, jintArray outIDs, jfloatArray outValues,
                                  jintArray outCount) {
  jint ne, no, nc;
  jint *e = ints(env, exclude, &ne), *oi = ints(env, outIDs, &no), *oc = ints(env, outCount, &nc);
  jfloat *f = (*env)->GetFloatArrayElements(env, features, NULL);
  jfloat *ov = (*env)->GetFloatArrayElements(env, outValues, NULL);
  int rc = als_top_n(H(h), which, f, nVectors, (const int32_t *)e, ne, howMany, (int32_t *)oi, ov, (int32_t *)oc);
  done_ints(env, exclude, e, JNI_ABORT);
  done_ints(env, outIDs, oi, 0);
  done_ints(env, outCount, oc, 0);
  (*env)->ReleaseFloatArrayElements(env, features, f, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, outValues, ov, 0);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nSetFoldInState)(JNIEnv *env, jclass c, jlong h, jint which, jdoubleArray qrt,
                                            jdoubleArray rdiag, jintArray perm, jdouble learnRate) {
  if (!qrt) return als_set_fold_in_state(H(h), which, NULL, NULL, NULL, learnRate);
  jint np;
  jdouble *q = (*env)->GetDoubleArrayElements(env, qrt, NULL);
  jdouble *r = (*env)->GetDoubleArrayElements(env, rdiag, NULL);
  jint *p = ints(env, perm, &np);
  int rc = als_set_fold_in_state(H(h), which, q, r, (const int32_t *)p, learnRate);
  (*env)->ReleaseDoubleArrayElements(env, qrt, q, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, rdiag, r, JNI_ABORT);
  done_ints(env, perm, p, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL CLS(nFoldIn)(JNIEnv *env, jclass c, jlong h, jintArray users, jintArray items,
                                    jfloatArray values) {
  jint nu, ni;
  jint *u = ints(env, users, &nu), *i = ints(env, items, &ni);
  jfloat *v = values ? (*env)->GetFloatArrayElements(env, values, NULL) : NULL;
  int rc = (nu == ni) ? als_fold_in(H(h), (const int32_t *)u, (const int32_t *)i, v, nu) : ALS_E_ARG;
  done_ints(env, users, u, JNI_ABORT);
  done_ints(env, items, i, JNI_ABORT);
  if (v) (*env)->ReleaseFloatArrayElements(env, values, v, JNI_ABORT);
  return rc;
}

JNIEXPORT jstring JNICALL CLS(nLastError)(JNIEnv *env, jclass c, jlong h) {
  return (*env)->NewStringUTF(env, als_last_error(H(h)));
}
