// JNI glue: org.apache.spark.ml.commons.NativeProjectedProcess  ->  the C-ABI of include/sgp.h.
// Not compiled in this repository's build (there is no JDK in the image): `__has_include(<jni.h>)` guards it so that
// `g++ -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude integration/jni/sgp_jni.cpp
//      -L spark_gp_b200 -lsgp -o libsgp_jni.so` builds it where a JDK exists.
// Arrays: Get<Primitive>ArrayElements copies/pins on the JVM side; the native context owns every device buffer.
#if __has_include(<jni.h>)
#include <jni.h>

#include <vector>

#include "../../include/sgp.h"

namespace {
void throw_for(JNIEnv* env, sgp_ctx* ctx, int rc) {
  const char* cls = "java/lang/RuntimeException";
  if (rc == SGP_E_NOT_PD) cls = "org/apache/spark/ml/commons/ProjectedGaussianProcessHelper$NotPositiveDefiniteException";
  else if (rc == SGP_E_BADARG) cls = "java/lang/IllegalArgumentException";
  else if (rc == SGP_E_STATE) cls = "org/apache/spark/ml/commons/kernel/TrainingVectorsNotInitializedException";
  else if (rc == SGP_E_SINGULAR) cls = "breeze/linalg/MatrixSingularException";
  env->ThrowNew(env->FindClass(cls), sgp_last_error(ctx));
}
}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_create(JNIEnv* env, jclass, jint device) {
  sgp_ctx* ctx = nullptr;
  const int rc = sgp_ctx_create(&ctx, device);
  if (rc != SGP_OK) { throw_for(env, nullptr, rc); return 0; }
  return reinterpret_cast<jlong>(ctx);
}

JNIEXPORT void JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_destroy(JNIEnv*, jclass, jlong h) {
  sgp_ctx_destroy(reinterpret_cast<sgp_ctx*>(h));
}

// terms: parallel arrays describing the flattened kernel (type, scale, sigma) + concatenated ARD betas (d each)
JNIEXPORT void JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_begin(
    JNIEnv* env, jclass, jlong h, jintArray types, jdoubleArray scales, jdoubleArray sigmas, jdoubleArray betas,
    jdoubleArray activeSet, jint m, jint d) {
  sgp_ctx* ctx = reinterpret_cast<sgp_ctx*>(h);
  const jsize nt = env->GetArrayLength(types);
  jint* ty = env->GetIntArrayElements(types, nullptr);
  jdouble* sc = env->GetDoubleArrayElements(scales, nullptr);
  jdouble* sg = env->GetDoubleArrayElements(sigmas, nullptr);
  jdouble* be = env->GetDoubleArrayElements(betas, nullptr);
  jdouble* z = env->GetDoubleArrayElements(activeSet, nullptr);
  std::vector<sgp_kernel_term> terms(nt);
  int ard = 0;
  for (jsize t = 0; t < nt; ++t) {
    terms[t].type = ty[t]; terms[t].reserved = 0; terms[t].scale = sc[t]; terms[t].sigma = sg[t];
    terms[t].beta = (ty[t] == SGP_TERM_ARD) ? be + (ard++) * d : nullptr;
  }
  sgp_kernel_desc desc{static_cast<int32_t>(nt), 0, terms.data()};
  const int rc = sgp_stats_begin(ctx, &desc, z, m, d);
  env->ReleaseIntArrayElements(types, ty, JNI_ABORT);
  env->ReleaseDoubleArrayElements(scales, sc, JNI_ABORT);
  env->ReleaseDoubleArrayElements(sigmas, sg, JNI_ABORT);
  env->ReleaseDoubleArrayElements(betas, be, JNI_ABORT);
  env->ReleaseDoubleArrayElements(activeSet, z, JNI_ABORT);
  if (rc != SGP_OK) throw_for(env, ctx, rc);
}

// one partition's points, packed row-major by the Scala side: X (n*d doubles), y (n doubles)
JNIEXPORT void JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_accumulate(
    JNIEnv* env, jclass, jlong h, jdoubleArray X, jdoubleArray y, jlong n) {
  sgp_ctx* ctx = reinterpret_cast<sgp_ctx*>(h);
  jdouble* x = static_cast<jdouble*>(env->GetPrimitiveArrayCritical(X, nullptr));
  jdouble* yy = static_cast<jdouble*>(env->GetPrimitiveArrayCritical(y, nullptr));
  const int rc = sgp_stats_accumulate(ctx, x, /*x_is_f32=*/0, yy, n);   // returns after the last H2D copy
  env->ReleasePrimitiveArrayCritical(y, yy, JNI_ABORT);
  env->ReleasePrimitiveArrayCritical(X, x, JNI_ABORT);
  if (rc != SGP_OK) throw_for(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_finish(
    JNIEnv* env, jclass, jlong h, jdoubleArray G, jdoubleArray b) {
  sgp_ctx* ctx = reinterpret_cast<sgp_ctx*>(h);
  jdouble* g = env->GetDoubleArrayElements(G, nullptr);
  jdouble* bb = env->GetDoubleArrayElements(b, nullptr);
  const int rc = sgp_stats_finish(ctx, g, bb);
  env->ReleaseDoubleArrayElements(G, g, 0);
  env->ReleaseDoubleArrayElements(b, bb, 0);
  if (rc != SGP_OK) throw_for(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_org_apache_spark_ml_commons_NativeProjectedProcess_magic(
    JNIEnv* env, jclass, jlong h, jdoubleArray G, jdoubleArray b, jdoubleArray magicVector, jdoubleArray magicMatrix) {
  sgp_ctx* ctx = reinterpret_cast<sgp_ctx*>(h);
  jdouble* g = env->GetDoubleArrayElements(G, nullptr);
  jdouble* bb = env->GetDoubleArrayElements(b, nullptr);
  jdouble* mv = env->GetDoubleArrayElements(magicVector, nullptr);
  jdouble* mm = env->GetDoubleArrayElements(magicMatrix, nullptr);
  const int rc = sgp_magic(ctx, g, bb, mv, mm);
  env->ReleaseDoubleArrayElements(G, g, JNI_ABORT);
  env->ReleaseDoubleArrayElements(b, bb, JNI_ABORT);
  env->ReleaseDoubleArrayElements(magicVector, mv, 0);
  env->ReleaseDoubleArrayElements(magicMatrix, mm, 0);
  if (rc != SGP_OK) throw_for(env, ctx, rc);
}

}  // extern "C"
#endif  // __has_include(<jni.h>)
