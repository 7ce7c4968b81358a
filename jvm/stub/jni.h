/* Minimal stand-in for the JDK's <jni.h>: ONLY for type-checking and link-checking jvm/b200gbm_jni.c in an image without a JDK
 * (tests/test_capi_cpu.py::test_jni_shim_compiles_and_links).  The layout of JNINativeInterface_ here is NOT the real one; a real
 * build must use $JAVA_HOME/include/jni.h. */
#ifndef B200GBM_STUB_JNI_H_
#define B200GBM_STUB_JNI_H_
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef int32_t jsize;
typedef uint8_t jboolean;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jclass (*FindClass)(JNIEnv*, const char*);
  jobjectArray (*NewObjectArray)(JNIEnv*, jsize, jclass, jobject);
  void (*SetObjectArrayElement)(JNIEnv*, jobjectArray, jsize, jobject);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  jdouble* (*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*);
  void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
};
#define JNI_ABORT 2
#endif
