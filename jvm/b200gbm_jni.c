/*
 * JNI shim: com.microsoft.ml.lightgbm.lightgbmlibJNI  ->  libb200gbm.so  (SOURCE ONLY — this image has no JDK / jni.h).
 *
 * The reference loads `_lightgbm` and `_lightgbm_swig` from the lightgbmlib jar
 * (lightgbm/src/main/scala/com/microsoft/ml/spark/lightgbm/LightGBMUtils.scala:38-41) and calls the SWIG-generated
 * class `lightgbmlib`, whose static methods forward to `lightgbmlibJNI` natives taking raw C pointers as jlong
 * (SWIGTYPE_p_* wrappers, read back with SwigPtrWrapper.getCPtrValue, lightgbm/src/main/scala/com/microsoft/lightgbm/SWIG.scala:8-13).
 * Because libb200gbm.so keeps LightGBM's C names and signatures, the shim is mechanical: one native per C function,
 * pointers as jlong, strings via GetStringUTFChars.  Build (on a box with a JDK):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include b200gbm_jni.c -L../mmlspark_b200/lib -lb200gbm -o lib_lightgbm_swig.so
 * and ship libb200gbm.so under the resource name lib_lightgbm.so so NativeLoader
 * (core/src/main/scala/com/microsoft/ml/spark/core/env/NativeLoader.java:47-68) needs no change.
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include "b200gbm_c_api.h"

#define JNI_FN(name) JNIEXPORT jint JNICALL Java_com_microsoft_ml_lightgbm_lightgbmlibJNI_##name
#define P(x) ((void*)(intptr_t)(x))

JNIEXPORT jstring JNICALL Java_com_microsoft_ml_lightgbm_lightgbmlibJNI_LGBM_1GetLastError(JNIEnv* env, jclass cls) {
  (void)cls;
  return (*env)->NewStringUTF(env, LGBM_GetLastError());
}
JNI_FN(LGBM_1NetworkInit)(JNIEnv* env, jclass cls, jstring machines, jint port, jint timeout, jint n) {
  (void)cls;
  const char* m = (*env)->GetStringUTFChars(env, machines, NULL);
  int rc = LGBM_NetworkInit(m, port, timeout, n);
  (*env)->ReleaseStringUTFChars(env, machines, m);
  return rc;
}
JNI_FN(LGBM_1NetworkFree)(JNIEnv* env, jclass cls) { (void)env; (void)cls; return LGBM_NetworkFree(); }

/* data / out are SWIG pointer wrappers: double_to_voidp_ptr(new_doubleArray(n)), voidpp_handle() */
JNI_FN(LGBM_1DatasetCreateFromMat)(JNIEnv* env, jclass cls, jlong data, jint data_type, jint nrow, jint ncol, jint row_major,
                                    jstring params, jlong reference, jlong out) {
  (void)cls;
  const char* p = (*env)->GetStringUTFChars(env, params, NULL);
  int rc = LGBM_DatasetCreateFromMat(P(data), data_type, nrow, ncol, row_major, p, P(reference), (DatasetHandle*)P(out));
  (*env)->ReleaseStringUTFChars(env, params, p);
  return rc;
}
JNI_FN(LGBM_1DatasetSetField)(JNIEnv* env, jclass cls, jlong h, jstring name, jlong data, jint n, jint type) {
  (void)cls;
  const char* s = (*env)->GetStringUTFChars(env, name, NULL);
  int rc = LGBM_DatasetSetField(P(h), s, P(data), n, type);
  (*env)->ReleaseStringUTFChars(env, name, s);
  return rc;
}
JNI_FN(LGBM_1DatasetGetNumData)(JNIEnv* env, jclass cls, jlong h, jlong out) { (void)env; (void)cls; return LGBM_DatasetGetNumData(P(h), (int*)P(out)); }
JNI_FN(LGBM_1DatasetGetNumFeature)(JNIEnv* env, jclass cls, jlong h, jlong out) { (void)env; (void)cls; return LGBM_DatasetGetNumFeature(P(h), (int*)P(out)); }
JNI_FN(LGBM_1DatasetFree)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; return LGBM_DatasetFree(P(h)); }

JNI_FN(LGBM_1BoosterCreate)(JNIEnv* env, jclass cls, jlong train, jstring params, jlong out) {
  (void)cls;
  const char* p = (*env)->GetStringUTFChars(env, params, NULL);
  int rc = LGBM_BoosterCreate(P(train), p, (BoosterHandle*)P(out));
  (*env)->ReleaseStringUTFChars(env, params, p);
  return rc;
}
JNI_FN(LGBM_1BoosterLoadModelFromString)(JNIEnv* env, jclass cls, jstring model, jlong out_iters, jlong out) {
  (void)cls;
  const char* m = (*env)->GetStringUTFChars(env, model, NULL);
  int rc = LGBM_BoosterLoadModelFromString(m, (int*)P(out_iters), (BoosterHandle*)P(out));
  (*env)->ReleaseStringUTFChars(env, model, m);
  return rc;
}
JNI_FN(LGBM_1BoosterMerge)(JNIEnv* env, jclass cls, jlong h, jlong other) { (void)env; (void)cls; return LGBM_BoosterMerge(P(h), P(other)); }
JNI_FN(LGBM_1BoosterAddValidData)(JNIEnv* env, jclass cls, jlong h, jlong v) { (void)env; (void)cls; return LGBM_BoosterAddValidData(P(h), P(v)); }
JNI_FN(LGBM_1BoosterFree)(JNIEnv* env, jclass cls, jlong h) { (void)env; (void)cls; return LGBM_BoosterFree(P(h)); }
/* the hot call: one boosting iteration on the GPU */
JNI_FN(LGBM_1BoosterUpdateOneIter)(JNIEnv* env, jclass cls, jlong h, jlong is_finished) {
  (void)env; (void)cls;
  return LGBM_BoosterUpdateOneIter(P(h), (int*)P(is_finished));
}
JNI_FN(LGBM_1BoosterUpdateOneIterCustom)(JNIEnv* env, jclass cls, jlong h, jlong grad, jlong hess, jlong is_finished) {
  (void)env; (void)cls;
  return LGBM_BoosterUpdateOneIterCustom(P(h), (const float*)P(grad), (const float*)P(hess), (int*)P(is_finished));
}
JNI_FN(LGBM_1BoosterResetParameter)(JNIEnv* env, jclass cls, jlong h, jstring params) {
  (void)cls;
  const char* p = (*env)->GetStringUTFChars(env, params, NULL);
  int rc = LGBM_BoosterResetParameter(P(h), p);
  (*env)->ReleaseStringUTFChars(env, params, p);
  return rc;
}
JNI_FN(LGBM_1BoosterGetEval)(JNIEnv* env, jclass cls, jlong h, jint idx, jlong out_len, jlong out) {
  (void)env; (void)cls;
  return LGBM_BoosterGetEval(P(h), idx, (int*)P(out_len), (double*)P(out));
}
/* SWIG convenience: returns the model text, retrying with the reported length (LightGBMBooster.scala:269-274) */
JNIEXPORT jstring JNICALL Java_com_microsoft_ml_lightgbm_lightgbmlibJNI_LGBM_1BoosterSaveModelToStringSWIG(
    JNIEnv* env, jclass cls, jlong h, jint start_iter, jint num_iter, jint imp_type, jlong buffer_len, jlong out_len) {
  (void)cls;
  int64_t* n = (int64_t*)P(out_len);
  char* buf = (char*)malloc((size_t)buffer_len);
  if (LGBM_BoosterSaveModelToString(P(h), start_iter, num_iter, imp_type, buffer_len, n, buf) != 0) { free(buf); return NULL; }
  if (*n > buffer_len) {
    free(buf);
    buf = (char*)malloc((size_t)*n);
    if (LGBM_BoosterSaveModelToString(P(h), start_iter, num_iter, imp_type, *n, n, buf) != 0) { free(buf); return NULL; }
  }
  jstring s = (*env)->NewStringUTF(env, buf);
  free(buf);
  return s;
}
/* ... the remaining natives (GetPredict, NumClasses, FeatureImportance, PredictForMatSingle, PredictForCSRSingle, DumpModel,
 * the new_/delete_/_getitem/_setitem array helpers and the three ChunkedArray classes -> B200GBM_ChunkedArray*) follow the same
 * one-line pattern; see include/b200gbm_c_api.h for the full list and INTEGRATION.md for the mapping table. */
