// Batched per-expert BCM objective: negative log marginal likelihood + gradient (fp64), one CTA per expert.
//
// Replaces, for all experts of a rank at once, the body of the hyper-parameter objective
//   regression/GaussianProcessRegression.scala:55-68   likelihoodAndGradient
//   commons/util/logDetAndInv.scala:36-63              (LU-based log|det| and inverse)
//   kernel/*.scala  trainingKernelAndDerivative        (ARDRBFKernel.scala:61-79, RBFKernel.scala:56-64,
//                                                       ScalarTimesKernel.scala:50-54,93-97, SumOfKernels.scala:50-55)
// summed over experts like the treeAggregate of commons/GaussianProcessCommons.scala:73-78.
//
//   K   = sum_t scale_t k_t(X_e, X_e) + eye_sum I                      (n_e x n_e, n_e ~ 100)
//   nll = 1/2 y^T K^-1 y + 1/2 log|det K|      (no n/2 log 2pi term -- GPR:61)
//   g_i = -1/2 sum_ab dK_i[a,b] (alpha_a alpha_b - K^-1[a,b]),   alpha = K^-1 y          (GPR:63-66)
//
// The reference factors K with LU; K is symmetric positive definite here (the sigma2 Eye term is always present,
// GPC:18), so Cholesky gives the same log-determinant and inverse (up to rounding) at a third of the work.  A
// non-positive pivot is reported (the reference would return a negative-determinant "logdet" of |det| instead).
#include "expert_common.cuh"

namespace sgp {
namespace {

constexpr int MAX_HYPERS = 72;      // 1 scale + 64 ARD betas + a few more

struct NllParams {
  const double* X;          // packed expert-major: expert e owns rows off[e] .. off[e+1]-1 (row-major, d columns)
  const double* y;
  const long long* off;     // [E+1]
  int n_max;                // largest expert
  int x_in_smem;            // the expert's rows are staged in shared memory
  HyperView hv;
  double* out;              // [E][1 + n_hypers]   per-expert (nll, gradient)
  int* flags;               // bit 0: a pivot was not positive
};

__global__ void __launch_bounds__(EX_THREADS, 2) bcm_nll_kernel(const NllParams p) {
  extern __shared__ double sm[];
  const long long e = blockIdx.x;
  const long long r0 = p.off[e];
  const int n = static_cast<int>(p.off[e + 1] - r0);
  const int ld = p.n_max + 1;                 // padded leading dimension
  double* K = sm;                             // [n_max][ld]   K -> L -> L^-1 -> K^-1 (lower triangle)
  double* yv = K + static_cast<size_t>(p.n_max) * ld;      // [n_max]
  double* alpha = yv + p.n_max;               // [n_max]
  double* rowbuf = alpha + p.n_max;           // [n_max]
  double* red = rowbuf + p.n_max;             // [8]
  double* sums = red + 8;                     // [EX_SUMS]
  double* Xs = sums + EX_SUMS;                // [n_max][d|1] staged copy of the expert's rows (if it fits)
  const int tid = threadIdx.x;
  int xld;
  const double* Xe = ex_stage_rows(p.X + static_cast<size_t>(r0) * p.hv.d, n, p.hv.d, p.x_in_smem, Xs, xld);
  for (int i = tid; i < n; i += EX_THREADS) yv[i] = p.y[r0 + i];
  __syncthreads();
  ex_build_kernel<false>(p.hv, Xe, xld, n, K, ld);
  __syncthreads();
  bool bad = false;
  const double logdet = 2.0 * ex_cholesky(K, n, ld, rowbuf, red, bad);
  if (bad && tid == 0) atomicOr(p.flags, 1);
  ex_invert_lower(K, n, ld, rowbuf);
  ex_ltl_inplace(K, n, ld, rowbuf);           // K now holds K^-1 (lower)
  // ---- alpha = K^-1 y ;  nll ----------------------------------------------------------------------------------------
  for (int a = tid; a < n; a += EX_THREADS) {
    double s = 0.0;
    for (int b = 0; b < n; ++b) s = fma((b <= a) ? K[a * ld + b] : K[b * ld + a], yv[b], s);
    alpha[a] = s;
  }
  __syncthreads();
  double part = 0.0;
  for (int a = tid; a < n; a += EX_THREADS) part += yv[a] * alpha[a];
  const double yay = ex_block_sum(part, red);
  double* out = p.out + static_cast<size_t>(e) * (1 + p.hv.n_hypers);
  if (tid == 0) out[0] = 0.5 * yay + 0.5 * logdet;
  // ---- gradient: -1/2 sum_ab dK_i[a,b] (alpha_a alpha_b - K^-1[a,b])   (GPR:63-66) -----------------------------------
  ex_descriptor_gradient(p.hv, Xe, xld, n, [&](int a, int b) { return alpha[a] * alpha[b] - K[a * ld + b]; }, -0.5,
                         out + 1, sums, red);
}

// Sum the per-expert rows: one CTA per column, thread t adds experts t, t+256, ... and the 256 partials are combined by
// a fixed tree -- the order never depends on timing, so the result is deterministic.
__global__ void __launch_bounds__(EX_THREADS) rows_reduce_kernel(double* __restrict__ total,
                                                                   const double* __restrict__ per_expert, long long E,
                                                                   int width) {
  __shared__ double part[EX_THREADS];
  const int c = blockIdx.x;
  double s = 0.0;
  for (long long e = threadIdx.x; e < E; e += EX_THREADS) s += per_expert[e * width + c];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = EX_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[c] = part[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// General path: experts of ANY size (GaussianProcessParams.scala:36 puts no bound on datasetSizeForExpert) and kernel
// matrices that are not positive definite.  Exactly the reference's arithmetic: LU with partial pivoting
// (commons/util/logDetAndInv.scala:58-63 -> LAPACK dgetrf), log|det| = sum log|u_ii| with the SIGN IGNORED (GPR:59 drops
// it), inverse from the same factors (dgetri).  The kernel matrices live in global memory ([E][n_max][n_max], experts
// smaller than n_max padded with an identity block); the factorizations are cuBLAS' batched LU / inverse (plain library
// calls on small dense matrices), the build / objective / gradient kernels are ours.
// ---------------------------------------------------------------------------------------------------------------------
struct GenParams {
  const double* X; const double* y; const long long* off;
  int n_max, x_in_smem;
  HyperView hv;
  double* A;                // [E][n_max][n_max]  K, then its LU factors
  const double* Ainv;       // [E][n_max][n_max]  inverse
  const int* info;          // [E] dgetrf info (> 0: exactly singular)
  double* out;              // [E][1 + n_hypers]
  int* flags;               // bit 1: some expert's matrix is singular (MatrixSingularException)
};

__global__ void __launch_bounds__(EX_THREADS) gen_build_kernel(const GenParams p, double** Aptr, double** Cptr, double* Ainv) {
  const long long e = blockIdx.x;
  const long long r0 = p.off[e];
  const int n = static_cast<int>(p.off[e + 1] - r0), ld = p.n_max;
  double* K = p.A + static_cast<size_t>(e) * ld * ld;
  if (threadIdx.x == 0) { Aptr[e] = K; Cptr[e] = Ainv + static_cast<size_t>(e) * ld * ld; }
  const double* Xe = p.X + static_cast<size_t>(r0) * p.hv.d;
  ex_build_kernel<true>(p.hv, Xe, p.hv.d, n, K, ld);
  for (int idx = threadIdx.x; idx < ld * ld; idx += EX_THREADS) {      // identity padding (does not change det or alpha)
    const int a = idx / ld, b = idx % ld;
    if (a >= n || b >= n) K[idx] = (a == b) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(EX_THREADS) gen_objective_kernel(const GenParams p) {
  extern __shared__ double sm[];
  const long long e = blockIdx.x;
  const long long r0 = p.off[e];
  const int n = static_cast<int>(p.off[e + 1] - r0), ld = p.n_max;
  double* yv = sm;                       // [n_max]
  double* alpha = yv + p.n_max;          // [n_max]
  double* red = alpha + p.n_max;         // [8]
  double* sums = red + 8;                // [EX_SUMS]
  double* Xs = sums + EX_SUMS;
  const int tid = threadIdx.x;
  const double* LU = p.A + static_cast<size_t>(e) * ld * ld;
  const double* Ki = p.Ainv + static_cast<size_t>(e) * ld * ld;
  int xld;
  const double* Xe = ex_stage_rows(p.X + static_cast<size_t>(r0) * p.hv.d, n, p.hv.d, p.x_in_smem, Xs, xld);
  for (int i = tid; i < n; i += EX_THREADS) yv[i] = p.y[r0 + i];
  __syncthreads();
  // log|det| = sum log|u_ii| (LU2logdet, logDetAndInv.scala:36-51; cuBLAS stores the factors column-major in place:
  // the diagonal is the same either way)
  double part = 0.0;
  for (int i = tid; i < n; i += EX_THREADS) part += log(fabs(LU[static_cast<size_t>(i) * ld + i]));
  const double logdet = ex_block_sum(part, red);
  if (tid == 0 && p.info[e] > 0) atomicOr(p.flags, 2);
  // alpha = K^-1 y  (the inverse of a symmetric matrix: row-/column-major reads agree up to rounding; use the average)
  for (int a = tid; a < n; a += EX_THREADS) {
    double s = 0.0;
    for (int b = 0; b < n; ++b)
      s = fma(0.5 * (Ki[static_cast<size_t>(a) * ld + b] + Ki[static_cast<size_t>(b) * ld + a]), yv[b], s);
    alpha[a] = s;
  }
  __syncthreads();
  part = 0.0;
  for (int a = tid; a < n; a += EX_THREADS) part += yv[a] * alpha[a];
  const double yay = ex_block_sum(part, red);
  double* out = p.out + static_cast<size_t>(e) * (1 + p.hv.n_hypers);
  if (tid == 0) out[0] = 0.5 * yay + 0.5 * logdet;                                            // GPR:61
  ex_descriptor_gradient(p.hv, Xe, xld, n,
                         [&](int a, int b) {
                           return alpha[a] * alpha[b] - 0.5 * (Ki[static_cast<size_t>(a) * ld + b] + Ki[static_cast<size_t>(b) * ld + a]);
                         },
                         -0.5, out + 1, sums, red);                                           // GPR:63-66
}

}  // namespace

size_t bcm_general_workspace_bytes(long long E, int n_max) {
  const size_t mat = static_cast<size_t>(E) * n_max * n_max * sizeof(double);
  return 2 * mat + static_cast<size_t>(E) * (2 * sizeof(double*) + sizeof(int) * (1 + static_cast<size_t>(n_max))) + 256;
}

// ws: bcm_general_workspace_bytes(E, n_max) bytes of device memory.  Returns cudaSuccess / a CUDA error; cuBLAS failures
// are reported through *blas_status.
cudaError_t launch_bcm_nll_general(cublasHandle_t blas, int* blas_status, void* ws, const double* dX, const double* dy,
                                   const long long* dOff, long long E, int d, int n_max, const KernelFlat& kf,
                                   const double* dBeta, int n_hypers, const int* dKind, const int* dTerm, const int* dDim,
                                   const double* dCoef, const double* dValue, int any_ard, double* dPerExpert,
                                   double* dTotal, int* dFlags, cudaStream_t s) {
  *blas_status = 0;
  const size_t mat = static_cast<size_t>(E) * n_max * n_max;
  double* A = static_cast<double*>(ws);
  double* Ainv = A + mat;
  double** Aptr = reinterpret_cast<double**>(Ainv + mat);
  double** Cptr = Aptr + E;
  int* info = reinterpret_cast<int*>(Cptr + E);
  int* piv = info + E;
  GenParams p{};
  p.X = dX; p.y = dy; p.off = dOff; p.n_max = n_max;
  p.hv = make_hyper_view(d, kf, dBeta, n_hypers, dKind, dTerm, dDim, dCoef, dValue, any_ard);
  p.A = A; p.Ainv = Ainv; p.info = info; p.out = dPerExpert; p.flags = dFlags;
  gen_build_kernel<<<static_cast<unsigned>(E), EX_THREADS, 0, s>>>(p, Aptr, Cptr, Ainv);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  cublasSetStream(blas, s);
  // K is symmetric, so cuBLAS' column-major view of the row-major buffer is the same matrix
  cublasStatus_t st = cublasDgetrfBatched(blas, n_max, Aptr, n_max, piv, info, static_cast<int>(E));
  if (st == CUBLAS_STATUS_SUCCESS)
    st = cublasDgetriBatched(blas, n_max, Aptr, n_max, piv, Cptr, n_max, info + 0, static_cast<int>(E));
  if (st != CUBLAS_STATUS_SUCCESS) { *blas_status = static_cast<int>(st); return cudaSuccess; }
  size_t smem = sizeof(double) * (2 * static_cast<size_t>(n_max) + 8 + EX_SUMS);
  const size_t with_x = smem + sizeof(double) * static_cast<size_t>(n_max) * (d | 1);
  p.x_in_smem = (with_x <= 200 * 1024) ? 1 : 0;
  if (p.x_in_smem) smem = with_x;
  e = cudaFuncSetAttribute(gen_objective_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  gen_objective_kernel<<<static_cast<unsigned>(E), EX_THREADS, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_rows_reduce(dTotal, dPerExpert, E, 1 + n_hypers, s);
}

size_t bcm_nll_smem_bytes(int n_max) {
  return sizeof(double) * (static_cast<size_t>(n_max) * (n_max + 1) + 3 * static_cast<size_t>(n_max) + 8 + EX_SUMS);
}
int bcm_nll_max_hypers() { return MAX_HYPERS; }

cudaError_t launch_rows_reduce(double* dTotal, const double* dPerExpert, long long E, int width, cudaStream_t s) {
  rows_reduce_kernel<<<width, EX_THREADS, 0, s>>>(dTotal, dPerExpert, E, width);
  return cudaGetLastError();
}

HyperView make_hyper_view(int d, const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind, const int* dTerm,
                          const int* dDim, const double* dCoef, const double* dValue, int any_ard) {
  HyperView hv{};
  hv.d = d; hv.n_terms = kf.n_terms;
  for (int t = 0; t < kMaxTerms; ++t) hv.scale[t] = kf.scale[t];
  hv.beta = dBeta; hv.eye_sum = kf.eye_sum;
  hv.n_hypers = n_hypers; hv.any_ard = any_ard;
  hv.h_kind = dKind; hv.h_term = dTerm; hv.h_dim = dDim; hv.h_coef = dCoef; hv.h_value = dValue;
  return hv;
}

cudaError_t launch_bcm_nll(const double* dX, const double* dy, const long long* dOff, long long E, int d, int n_max,
                           const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind, const int* dTerm,
                           const int* dDim, const double* dCoef, const double* dValue, int any_ard,
                           double* dPerExpert, double* dTotal, int* dFlags, cudaStream_t s) {
  NllParams p{};
  p.X = dX; p.y = dy; p.off = dOff; p.n_max = n_max;
  p.hv = make_hyper_view(d, kf, dBeta, n_hypers, dKind, dTerm, dDim, dCoef, dValue, any_ard);
  p.out = dPerExpert; p.flags = dFlags;
  // stage the expert's rows in shared memory when two CTAs per SM still fit (113 KB each)
  size_t smem = bcm_nll_smem_bytes(n_max);
  const size_t with_x = smem + sizeof(double) * static_cast<size_t>(n_max) * (d | 1);
  p.x_in_smem = (with_x <= 113 * 1024) ? 1 : 0;
  if (p.x_in_smem) smem = with_x;
  cudaError_t e = cudaFuncSetAttribute(bcm_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  bcm_nll_kernel<<<static_cast<unsigned>(E), EX_THREADS, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_rows_reduce(dTotal, dPerExpert, E, 1 + n_hypers, s);
}

}  // namespace sgp
