// The m x m tail (fp64, one GPU) and block prediction.
//
//   commons/ProjectedGaussianProcessHelper.scala:49-65   getMagicVector + assertSymPositiveDefinite
//   commons/GaussianProcessCommons.scala:118-126         GaussianProjectedProcessRawPredictor.predict
//
// The reference's dense linear algebra is LAPACK through Breeze (dsyevd for the PD check, dgesv for `\`,
// dgetrf+dgetri for `inv`), all on one driver thread.  Here the tail runs on the device with cuSOLVER / cuBLAS
// (plain library factorizations of an m x m matrix -- not the hot path) and is organised for latency:
//
//   fast path (every well-posed model):  A = whiteNoiseVar K_mm + G is symmetric.  A successful Cholesky
//     factorization proves that every eigenvalue is > 0, i.e. the reference's check "no eigenvalue < 0" (PGPH:63)
//     passes -- without the O(10 m^3) dsyevd that dominated round 1's tail (26.5 ms at m = 1000).  The same factors
//     give  magicVector = A \ b  and  inv(A)  (dpotrs); K_mm (SPD: a kernel matrix plus the Eye terms) is inverted the
//     same way on a SECOND stream with its own cuSOLVER handle, concurrently with the A chain.  For SPD matrices the
//     Cholesky and LU solutions agree to rounding (both backward stable; the reference's LU pivots are a no-op
//     on a diagonally dominant SPD matrix up to rounding) -- the parity tests (mean / variance <= 1e-5 vs the oracle's
//     LAPACK LU) are unchanged.
//   slow path (Cholesky of A or K_mm breaks down):  exactly the reference's sequence -- dsyevd, throw iff an
//     eigenvalue is < 0 (a semi-definite matrix passes, PGPH:63 is a strict comparison), then LU solves/inverses
//     (dgetrf / dgetrs), SGP_E_SINGULAR on a zero pivot like Breeze's MatrixSingularException.
//
// All device workspaces live in the context and are reused across calls (cudaMalloc/cudaFree synchronise the device
// and cost more than the factorizations at m = 1000).
#include "sgp_internal.h"

namespace sgp {

#define SGP_SOLVER(c, expr)                                                                    \
  do {                                                                                         \
    cusolverStatus_t st_ = (expr);                                                             \
    if (st_ != CUSOLVER_STATUS_SUCCESS)                                                        \
      return fail((c), SGP_E_CUDA, std::string(#expr) + ": cusolver status " + std::to_string((int)st_)); \
  } while (0)

// Grow-only device scratch owned by the context.
int ctx_scratch(Ctx* c, DevScratch& s, size_t bytes) {
  if (bytes <= s.cap && s.p) return SGP_OK;
  if (s.p) {
    SGP_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->tail_stream) SGP_CUDA(c, cudaStreamSynchronize(c->tail_stream));
    cudaFree(s.p);
    s.p = nullptr; s.cap = 0;
  }
  SGP_CUDA(c, cudaMalloc(&s.p, bytes ? bytes : 8));
  s.cap = bytes ? bytes : 8;
  return SGP_OK;
}

static const char* kNotPdMsg =
    "Some matrix which is supposed to be positive definite is not. This probably happened due "
    "to `sigma2` parameter being too small. Try to gradually increase it.";

// The reference's own sequence on one matrix (slow path): M is overwritten by its LU factors.
// inv_out = M^-1 ; optionally rhs_vec <- M^-1 rhs_vec.
static int lu_inverse(Ctx* c, cusolverDnHandle_t h, cudaStream_t s, int m, double* M, double* inv_out, double* rhs_vec,
                      double* work, int* ipiv, int* info) {
  int hinfo = 0;
  SGP_SOLVER(c, cusolverDnDgetrf(h, m, m, M, m, work, ipiv, info));
  SGP_CUDA(c, cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
  SGP_CUDA(c, cudaStreamSynchronize(s));
  if (hinfo > 0) return fail(c, SGP_E_SINGULAR, "matrix is singular (LU pivot " + std::to_string(hinfo) + ")");
  if (hinfo < 0) return fail(c, SGP_E_CUDA, "dgetrf bad argument " + std::to_string(-hinfo));
  if (rhs_vec) SGP_SOLVER(c, cusolverDnDgetrs(h, CUBLAS_OP_N, m, 1, M, m, ipiv, rhs_vec, m, info));
  SGP_CUDA(c, launch_set_identity(inv_out, m, s));
  c->launches += 1;
  SGP_SOLVER(c, cusolverDnDgetrs(h, CUBLAS_OP_N, m, m, M, m, ipiv, inv_out, m, info));
  return SGP_OK;
}

int run_tail(Ctx* c, double* magic_vector, double* magic_matrix) {
  const int m = c->m;
  const size_t mm = static_cast<size_t>(m) * m;
  cudaStream_t s = c->stream, s2 = c->tail_stream;
  SGP_SOLVER(c, cusolverDnSetStream(c->solver, s));
  SGP_SOLVER(c, cusolverDnSetStream(c->solver2, s2));

  // ---- persistent workspaces -------------------------------------------------------------------------------
  int lw_potrf = 0, lw_eig = 0, lw_lu = 0;
  SGP_SOLVER(c, cusolverDnDpotrf_bufferSize(c->solver, CUBLAS_FILL_MODE_LOWER, m, c->dGb, m, &lw_potrf));
  SGP_SOLVER(c, cusolverDnDgetrf_bufferSize(c->solver, m, m, c->dGb, m, &lw_lu));
  int lwork = lw_potrf > lw_lu ? lw_potrf : lw_lu;
  // layout of the one scratch block: [Kmm | A | invA | invK | Acopy] (mm doubles each) [work | work2] [ipiv] [info x 4]
  const size_t need = (5 * mm + 2 * static_cast<size_t>(lwork)) * 8 + static_cast<size_t>(m) * sizeof(int) + 64;
  int rc = ctx_scratch(c, c->tail_ws, need);
  if (rc != SGP_OK) return rc;
  double* Kmm = static_cast<double*>(c->tail_ws.p);
  double* A = Kmm + mm;
  double* invA = A + mm;
  double* invK = invA + mm;
  double* Acopy = invK + mm;
  double* work = Acopy + mm;
  double* work2 = work + lwork;
  int* ipiv = reinterpret_cast<int*>(work2 + lwork);
  int* info = ipiv + m;                                               // info[0]: A chain, info[1]: K_mm chain

  const double wn = c->kf.eye_sum;                                   // kernel.whiteNoiseVar
  const double* G = c->dGb;
  const double* b = c->dGb + mm;

  // K_mm = trainingKernel (PGPH:54) on the main stream; both chains need it
  SGP_CUDA(c, launch_kmm_build(Kmm, c->dZs, nullptr, c->kf, m, c->m_pad, c->dpad, s));
  SGP_CUDA(c, launch_axpby_diag(A, Kmm, G, wn, m, s));                                                 // PGPH:55-56
  c->launches += 2;
  SGP_CUDA(c, cudaMemcpyAsync(Acopy, A, mm * 8, cudaMemcpyDeviceToDevice, s));     // kept for the slow path
  SGP_CUDA(c, cudaEventRecord(c->tail_fork, s));
  SGP_CUDA(c, cudaStreamWaitEvent(s2, c->tail_fork, 0));

  // ---- chain 2 (second stream): inv(K_mm) by Cholesky ------------------------------------------------------
  SGP_CUDA(c, cudaMemcpyAsync(invA, Kmm, mm * 8, cudaMemcpyDeviceToDevice, s2));   // invA doubles as K_mm's factor buffer
  SGP_SOLVER(c, cusolverDnDpotrf(c->solver2, CUBLAS_FILL_MODE_LOWER, m, invA, m, work2, lwork, info + 1));
  SGP_CUDA(c, launch_set_identity(invK, m, s2));
  SGP_SOLVER(c, cusolverDnDpotrs(c->solver2, CUBLAS_FILL_MODE_LOWER, m, m, invA, m, invK, m, info + 2));
  SGP_CUDA(c, cudaEventRecord(c->tail_join, s2));

  // ---- chain 1 (main stream): Cholesky of A = PD check; magicVector = A \ b -----------------------------------
  SGP_SOLVER(c, cusolverDnDpotrf(c->solver, CUBLAS_FILL_MODE_LOWER, m, A, m, work, lwork, info));
  SGP_CUDA(c, cudaMemcpyAsync(c->dMagicVec, b, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToDevice, s));
  SGP_SOLVER(c, cusolverDnDpotrs(c->solver, CUBLAS_FILL_MODE_LOWER, m, 1, A, m, c->dMagicVec, m, info + 3));
  c->launches += 1;
  // inv(A) goes into c->dMagicMat (finished in place below): invA is still K_mm's factor buffer on the other stream
  SGP_CUDA(c, launch_set_identity(c->dMagicMat, m, s));
  SGP_SOLVER(c, cusolverDnDpotrs(c->solver, CUBLAS_FILL_MODE_LOWER, m, m, A, m, c->dMagicMat, m, info + 3));
  c->launches += 1;

  int hinfo[2] = {0, 0};
  SGP_CUDA(c, cudaStreamWaitEvent(s, c->tail_join, 0));
  SGP_CUDA(c, cudaMemcpyAsync(hinfo, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  SGP_CUDA(c, cudaStreamSynchronize(s));
  if (hinfo[0] < 0 || hinfo[1] < 0) return fail(c, SGP_E_CUDA, "dpotrf bad argument");

  bool a_done = (hinfo[0] == 0), k_done = (hinfo[1] == 0);
  c->tail_fast = a_done && k_done;
  c->has_magic_run = true;
  if (!a_done) {
    // ---- slow path for A: the reference's literal check, then LU ---------------------------------------------
    int lw = 0;
    SGP_SOLVER(c, cusolverDnDsyevd_bufferSize(c->solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m, A, m,
                                              invA, &lw_eig));
    lw = lw_eig > lw_lu ? lw_eig : lw_lu;
    rc = ctx_scratch(c, c->tail_ws2, (static_cast<size_t>(lw) + m) * 8);
    if (rc != SGP_OK) return rc;
    double* wk = static_cast<double*>(c->tail_ws2.p);
    double* W = wk + lw;
    SGP_CUDA(c, cudaMemcpyAsync(A, Acopy, mm * 8, cudaMemcpyDeviceToDevice, s));
    SGP_SOLVER(c, cusolverDnDsyevd(c->solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m, A, m, W, wk, lw, info));
    std::vector<double> ev(m);
    int einfo = 0;
    SGP_CUDA(c, cudaMemcpyAsync(ev.data(), W, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaMemcpyAsync(&einfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaStreamSynchronize(s));
    if (einfo != 0) return fail(c, SGP_E_CUDA, "dsyevd did not converge, info=" + std::to_string(einfo));
    for (int i = 0; i < m; ++i)
      if (ev[i] < 0.0 || ev[i] != ev[i]) return fail(c, SGP_E_NOT_PD, kNotPdMsg);        // PGPH:62-65
    // semi-definite (or PD only to rounding): carry on with LU like the reference                      PGPH:59
    SGP_CUDA(c, cudaMemcpyAsync(A, Acopy, mm * 8, cudaMemcpyDeviceToDevice, s));
    SGP_CUDA(c, cudaMemcpyAsync(c->dMagicVec, b, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToDevice, s));
    rc = lu_inverse(c, c->solver, s, m, A, c->dMagicMat, c->dMagicVec, wk, ipiv, info);
    if (rc != SGP_OK) return rc;
  }
  if (!k_done) {
    // Cholesky of K_mm broke down (e.g. a kernel without any Eye term on duplicated active points): LU like the reference
    rc = ctx_scratch(c, c->tail_ws2, (static_cast<size_t>(lw_lu) + m) * 8);
    if (rc != SGP_OK) return rc;
    rc = lu_inverse(c, c->solver, s, m, Kmm, invK, nullptr, static_cast<double*>(c->tail_ws2.p), ipiv, info);
    if (rc != SGP_OK) return rc;
  }
  // magicMatrix = inv(A) * whiteNoiseVar - inv(K_mm)                                                    PGPH:59
  SGP_CUDA(c, launch_magic_matrix(c->dMagicMat, c->dMagicMat, invK, wn, m, s));
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
  const long long cn = n < chunk ? n : chunk;
  // one grow-only block: [X | K | W | mean | var]
  const size_t xb = static_cast<size_t>(cn) * d, kb = static_cast<size_t>(cn) * m;
  int rc = ctx_scratch(c, c->predict_ws, (xb + 2 * kb + 2 * static_cast<size_t>(cn)) * 8);
  if (rc != SGP_OK) return rc;
  double* dX = static_cast<double*>(c->predict_ws.p);
  double* dK = dX + xb;
  double* dW = dK + kb;
  double* dMean = dW + kb;
  double* dVar = dMean + cn;
  if (cublasSetStream(c->blas, s) != CUBLAS_STATUS_SUCCESS) return fail(c, SGP_E_CUDA, "cublasSetStream");
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    const long long rn = (n - r0 < chunk) ? (n - r0) : chunk;
    SGP_CUDA(c, cudaMemcpyAsync(dX, X + static_cast<size_t>(r0) * d, static_cast<size_t>(rn) * d * 8,
                                cudaMemcpyHostToDevice, s));
    SGP_CUDA(c, launch_cross_kernel(dK, dX, c->dZs, c->dBeta, c->kf, rn, d, c->dpad, m, c->m_pad, s));
    c->launches += 1;
    if (var_out) {
      // W (rn x m, row-major) = K * magicMatrix  ==  column-major  W^T = M * K^T     (plain library GEMM)
      const double one = 1.0, zero = 0.0;
      cublasStatus_t st = cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, m, static_cast<int>(rn), m, &one,
                                      c->dMagicMat, m, dK, m, &zero, dW, m);
      if (st != CUBLAS_STATUS_SUCCESS) return fail(c, SGP_E_CUDA, "cublasDgemm status " + std::to_string((int)st));
    }
    SGP_CUDA(c, launch_predict_finish(dMean, var_out ? dVar : nullptr, dK, var_out ? dW : nullptr, c->dMagicVec,
                                      c->kf.self_kernel, rn, m, s));
    c->launches += 1;
    SGP_CUDA(c, cudaMemcpyAsync(mean_out + r0, dMean, static_cast<size_t>(rn) * 8, cudaMemcpyDeviceToHost, s));
    if (var_out) SGP_CUDA(c, cudaMemcpyAsync(var_out + r0, dVar, static_cast<size_t>(rn) * 8, cudaMemcpyDeviceToHost, s));
    SGP_CUDA(c, cudaStreamSynchronize(s));
  }
  return SGP_OK;
}

}  // namespace sgp
