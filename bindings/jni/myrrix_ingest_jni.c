/*
 * JNI stub between NativeInputFilesReader.java and the C ABI (include/myrrix_ingest.h).
 * NOT compiled here (no jni.h in this image).  Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       bindings/jni/myrrix_ingest_jni.c -o libmyrrix_ingest_jni.so -L<dir> -lmyrrix_ingest
 * Every function is a 1:1 forward; buffers are direct ByteBuffers so no copy happens here.
 */
#include <jni.h>
#include <stdint.h>

#include "myrrix_ingest.h"

#define H(x) ((ingest_handle *)(intptr_t)(x))
#define CLS(name) Java_net_myrrix_online_generation_NativeInputFilesReader_##name
#define BUF(b) ((*env)->GetDirectBufferAddress(env, (b)))

JNIEXPORT jlong JNICALL CLS(nCreate)(JNIEnv *env, jclass c, jfloat zeroThreshold) {
  ingest_handle *h = NULL;
  if (ingest_create(zeroThreshold, &h) != INGEST_OK) return 0;
  return (jlong)(intptr_t)h;
}
JNIEXPORT void JNICALL CLS(nDestroy)(JNIEnv *env, jclass c, jlong h) { ingest_destroy(H(h)); }
JNIEXPORT jint JNICALL CLS(nAddFile)(JNIEnv *env, jclass c, jlong h, jobject bytes, jlong len) {
  return ingest_add_file(H(h), (const char *)BUF(bytes), (size_t)len);
}
JNIEXPORT jint JNICALL CLS(nFinish)(JNIEnv *env, jclass c, jlong h) { return ingest_finish(H(h)); }
JNIEXPORT jlong JNICALL CLS(nCount)(JNIEnv *env, jclass c, jlong h, jint kind) { return ingest_count(H(h), kind); }
JNIEXPORT jint JNICALL CLS(nGetIds)(JNIEnv *env, jclass c, jlong h, jint which, jobject out) {
  return ingest_get_ids(H(h), which, (int64_t *)BUF(out));
}
JNIEXPORT jint JNICALL CLS(nGetCsr)(JNIEnv *env, jclass c, jlong h, jobject rowPtr, jobject colIdx, jobject val) {
  return ingest_get_csr(H(h), (int64_t *)BUF(rowPtr), (int32_t *)BUF(colIdx), (float *)BUF(val));
}
JNIEXPORT jint JNICALL CLS(nGetKnown)(JNIEnv *env, jclass c, jlong h, jobject rowPtr, jobject colIdx) {
  return ingest_get_known(H(h), (int64_t *)BUF(rowPtr), (int32_t *)BUF(colIdx));
}
JNIEXPORT jint JNICALL CLS(nGetTags)(JNIEnv *env, jclass c, jlong h, jint which, jobject out) {
  return ingest_get_tags(H(h), which, (int64_t *)BUF(out));
}
JNIEXPORT jstring JNICALL CLS(nLastError)(JNIEnv *env, jclass c, jlong h) {
  return (*env)->NewStringUTF(env, ingest_last_error(H(h)));
}
