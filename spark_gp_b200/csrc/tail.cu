// The m x m tail (fp64, one GPU) and block prediction.
//
//   commons/ProjectedGaussianProcessHelper.scala:49-65   getMagicVector + assertSymPositiveDefinite
//   commons/GaussianProcessCommons.scala:118-126         GaussianProjectedProcessRawPredictor.predict
//
// The reference's dense linear algebra is LAPACK through Breeze (dsyevd for the PD check, dgesv for `\`,
// dgetrf+dgetri for `inv`); here the same factorizations run on the device through cuSOLVER (plain library
// factorizations of an m x m matrix -- not the hot path).  Same algorithm choices as the reference on
// purpose: LU (not Cholesky), explicit inverses, PD check = "any eigenvalue < 0".
#include "sgp_internal.h"

namespace sgp {

#define SGP_SOLVER(c, expr)                                                                    \
  do {                                                                                         \
    cusolverStatus_t st_ = (expr);                                                             \
    if (st_ != CUSOLVER_STATUS_SUCCESS)                                                        \
      return fail((c), SGP_E_CUDA, std::string(#expr) + ": cusolver status " + std::to_string((int)st_)); \
  } while (0)

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};
}  // namespace

int run_tail(Ctx* c, double* magic_vector, double* magic_matrix) {
  const int m = c->m;
  const size_t mm = static_cast<size_t>(m) * m;
  cudaStream_t s = c->stream;
  SGP_SOLVER(c, cusolverDnSetStream(c->solver, s));

  DevBuf Kmm, A, T1, T2, W, ipiv, info, work;
  SGP_CUDA(c, Kmm.alloc(mm * 8));
  SGP_CUDA(c, A.alloc(mm * 8));
  SGP_CUDA(c, T1.alloc(mm * 8));
  SGP_CUDA(c, T2.alloc(mm * 8));
  SGP_CUDA(c, W.alloc(static_cast<size_t>(m) * 8));
  SGP_CUDA(c, ipiv.alloc(static_cast<size_t>(m) * sizeof(int)));
  SGP_CUDA(c, info.alloc(sizeof(int)));

  const double wn = c->kf.eye_sum;                                   // kernel.whiteNoiseVar
  const double* G = c->dGb;
  const double* b = c->dGb + mm;

  SGP_CUDA(c, launch_kmm_build(Kmm.as<double>(), c->dZs, nullptr, c->kf, m, c->m_pad, c->dpad, s));   // PGPH:54
  SGP_CUDA(c, launch_axpby_diag(A.as<double>(), Kmm.as<double>(), G, wn, m, s));                      // PGPH:55-56
  c->launches += 2;

  int lwork_eig = 0, lwork_lu = 0;
  SGP_SOLVER(c, cusolverDnDsyevd_bufferSize(c->solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m,
                                            T1.as<double>(), m, W.as<double>(), &lwork_eig));
  SGP_SOLVER(c, cusolverDnDgetrf_bufferSize(c->solver, m, m, A.as<double>(), m, &lwork_lu));
  const int lwork = lwork_eig > lwork_lu ? lwork_eig : lwork_lu;
  SGP_CUDA(c, work.alloc(static_cast<size_t>(lwork) * 8));

  // ---- assertSymPositiveDefinite (PGPH:62-65): any eigenvalue < 0 -> NotPositiveDefiniteException ----
  SGP_CUDA(c, cudaMemcpyAsync(T1.p, A.p, mm * 8, cudaMemcpyDeviceToDevice, s));
  SGP_SOLVER(c, cusolverDnDsyevd(c->solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m, T1.as<double>(),
                                 m, W.as<double>(), work.as<double>(), lwork, info.as<int>()));
  {
    std::vector<double> ev(m);
    int hinfo = 0;
    SGP_CUDA(c, cudaMemcpyAsync(ev.data(), W.p, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaStreamSynchronize(s));
    if (hinfo != 0) return fail(c, SGP_E_CUDA, "dsyevd did not converge, info=" + std::to_string(hinfo));
    for (int i = 0; i < m; ++i)
      if (ev[i] < 0.0 || ev[i] != ev[i])
        return fail(c, SGP_E_NOT_PD,
                    "Some matrix which is supposed to be positive definite is not. This probably happened due "
                    "to `sigma2` parameter being too small. Try to gradually increase it.");
  }

  auto lu_inverse = [&](double* M, double* inv_out, double* rhs_vec) -> int {
    // M is overwritten by its LU factors.  inv_out = M^-1 ; optionally rhs_vec <- M^-1 rhs_vec.
    int hinfo = 0;
    SGP_SOLVER(c, cusolverDnDgetrf(c->solver, m, m, M, m, work.as<double>(), ipiv.as<int>(), info.as<int>()));
    SGP_CUDA(c, cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaStreamSynchronize(s));
    if (hinfo > 0) return fail(c, SGP_E_SINGULAR, "matrix is singular (LU pivot " + std::to_string(hinfo) + ")");
    if (hinfo < 0) return fail(c, SGP_E_CUDA, "dgetrf bad argument " + std::to_string(-hinfo));
    if (rhs_vec)
      SGP_SOLVER(c, cusolverDnDgetrs(c->solver, CUBLAS_OP_N, m, 1, M, m, ipiv.as<int>(), rhs_vec, m, info.as<int>()));
    SGP_CUDA(c, launch_set_identity(inv_out, m, s));
    c->launches += 1;
    SGP_SOLVER(c, cusolverDnDgetrs(c->solver, CUBLAS_OP_N, m, m, M, m, ipiv.as<int>(), inv_out, m, info.as<int>()));
    return SGP_OK;
  };

  // magicVector = A \ b ; inv(A)                                                    PGPH:59
  SGP_CUDA(c, cudaMemcpyAsync(c->dMagicVec, b, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToDevice, s));
  int rc = lu_inverse(A.as<double>(), T1.as<double>(), c->dMagicVec);
  if (rc != SGP_OK) return rc;
  // inv(K_mm)
  rc = lu_inverse(Kmm.as<double>(), T2.as<double>(), nullptr);
  if (rc != SGP_OK) return rc;
  // magicMatrix = inv(A) * whiteNoiseVar - inv(K_mm)
  SGP_CUDA(c, launch_magic_matrix(c->dMagicMat, T1.as<double>(), T2.as<double>(), wn, m, s));
  c->launches += 1;

  if (magic_vector)
    SGP_CUDA(c, cudaMemcpyAsync(magic_vector, c->dMagicVec, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToHost, s));
  if (magic_matrix) SGP_CUDA(c, cudaMemcpyAsync(magic_matrix, c->dMagicMat, mm * 8, cudaMemcpyDeviceToHost, s));
  SGP_CUDA(c, cudaStreamSynchronize(s));
  c->has_magic = true;
  return SGP_OK;
}

int run_predict(Ctx* c, const double* X, long long n, double* mean_out, double* var_out) {
  const int m = c->m, d = c->d;
  cudaStream_t s = c->stream;
  const long long chunk = 32768;
  DevBuf dX, dK, dW, dMean, dVar;
  const long long cn = n < chunk ? n : chunk;
  SGP_CUDA(c, dX.alloc(static_cast<size_t>(cn) * d * 8));
  SGP_CUDA(c, dK.alloc(static_cast<size_t>(cn) * m * 8));
  SGP_CUDA(c, dW.alloc(static_cast<size_t>(cn) * m * 8));
  SGP_CUDA(c, dMean.alloc(static_cast<size_t>(cn) * 8));
  SGP_CUDA(c, dVar.alloc(static_cast<size_t>(cn) * 8));
  if (cublasSetStream(c->blas, s) != CUBLAS_STATUS_SUCCESS) return fail(c, SGP_E_CUDA, "cublasSetStream");
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    const long long rn = (n - r0 < chunk) ? (n - r0) : chunk;
    SGP_CUDA(c, cudaMemcpyAsync(dX.p, X + static_cast<size_t>(r0) * d, static_cast<size_t>(rn) * d * 8,
                                cudaMemcpyHostToDevice, s));
    SGP_CUDA(c, launch_cross_kernel(dK.as<double>(), dX.as<double>(), c->dZs, c->dBeta, c->kf, rn, d, c->dpad, m,
                                    c->m_pad, s));
    c->launches += 1;
    if (var_out) {
      // W (rn x m, row-major) = K * magicMatrix  ==  column-major  W^T = M * K^T     (plain library GEMM)
      const double one = 1.0, zero = 0.0;
      cublasStatus_t st = cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, m, static_cast<int>(rn), m, &one,
                                      c->dMagicMat, m, dK.as<double>(), m, &zero, dW.as<double>(), m);
      if (st != CUBLAS_STATUS_SUCCESS) return fail(c, SGP_E_CUDA, "cublasDgemm status " + std::to_string((int)st));
    }
    SGP_CUDA(c, launch_predict_finish(dMean.as<double>(), var_out ? dVar.as<double>() : nullptr, dK.as<double>(),
                                      var_out ? dW.as<double>() : nullptr, c->dMagicVec, c->kf.self_kernel, rn, m,
                                      s));
    c->launches += 1;
    SGP_CUDA(c, cudaMemcpyAsync(mean_out + r0, dMean.p, static_cast<size_t>(rn) * 8, cudaMemcpyDeviceToHost, s));
    if (var_out)
      SGP_CUDA(c, cudaMemcpyAsync(var_out + r0, dVar.p, static_cast<size_t>(rn) * 8, cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaStreamSynchronize(s));
  }
  return SGP_OK;
}

}  // namespace sgp
