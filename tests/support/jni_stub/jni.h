// Minimal stand-in for <jni.h> (types and the JNIEnv members integration/jni/sgp_jni.cpp uses): lets the CPU test suite
// syntax-check the JNI glue in an image without a JDK.  Not a JNI implementation.
#pragma once
#include <cstdint>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef double jdouble; typedef int32_t jsize; typedef unsigned char jboolean;
struct _jobject {}; typedef _jobject* jobject; typedef jobject jclass; typedef jobject jarray; typedef jarray jintArray;
typedef jarray jdoubleArray; typedef jarray jlongArray; typedef jobject jstring;
struct JNIEnv {
  jclass FindClass(const char*); void ExceptionClear(); jint ThrowNew(jclass, const char*);
  jsize GetArrayLength(jarray); jint* GetIntArrayElements(jintArray, jboolean*); jdouble* GetDoubleArrayElements(jdoubleArray, jboolean*);
  jlong* GetLongArrayElements(jlongArray, jboolean*);
  void ReleaseIntArrayElements(jintArray, jint*, jint); void ReleaseDoubleArrayElements(jdoubleArray, jdouble*, jint);
  void ReleaseLongArrayElements(jlongArray, jlong*, jint);
  jdoubleArray NewDoubleArray(jsize); void SetDoubleArrayRegion(jdoubleArray, jsize, jsize, const jdouble*);
  jlongArray NewLongArray(jsize); void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong*);
  void* GetPrimitiveArrayCritical(jarray, jboolean*); void ReleasePrimitiveArrayCritical(jarray, void*, jint);
};
