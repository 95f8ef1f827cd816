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
#include "sgp_internal.h"

namespace sgp {
namespace {

constexpr int NLL_THREADS = 256;
constexpr int MAX_HYPERS = 72;      // 1 scale + 64 ARD betas + a few more

struct NllParams {
  const double* X;          // packed expert-major: expert e owns rows off[e] .. off[e+1]-1 (row-major, d columns)
  const double* y;
  const long long* off;     // [E+1]
  int d;
  int n_max;                // largest expert
  int n_terms;              // non-Eye terms
  double scale[kMaxTerms];
  const double* beta;       // [n_terms][d]  per-term coordinate scales (ARD betas; RBF: 1/(sqrt2 sigma))
  double eye_sum;
  int n_hypers;
  const int* h_kind;        // [n_hypers] 0 = SCALE, 1 = ARD_BETA, 2 = RBF_SIGMA
  const int* h_term;        // [n_hypers] term index (ARD_BETA / RBF_SIGMA)
  const int* h_dim;         // [n_hypers] feature index (ARD_BETA)
  const double* h_coef;     // [n_hypers][kMaxTerms+1]  SCALE: d(scale_t)/d(theta_i) per term, last = d(eye_sum)/d(theta_i)
  const double* h_value;    // [n_hypers] current value of the hyper-parameter (beta_k or sigma)
  double* out;              // [E][1 + n_hypers]   per-expert (nll, -2*grad sums)
  int* flags;               // bit 0: a pivot was not positive
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < NLL_THREADS / 32; ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(NLL_THREADS, 1) bcm_nll_kernel(const NllParams p) {
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
  double* gacc = red + 8;                     // [MAX_HYPERS]
  const int tid = threadIdx.x;
  const double* Xe = p.X + static_cast<size_t>(r0) * p.d;

  for (int i = tid; i < n; i += NLL_THREADS) yv[i] = p.y[r0 + i];
  // ---- K (lower triangle incl. diagonal; mirrored on the fly where needed) -----------------------------------
  for (int idx = tid; idx < n * n; idx += NLL_THREADS) {
    const int a = idx / n, b = idx % n;
    if (b > a) continue;
    double v = 0.0;
    for (int t = 0; t < p.n_terms; ++t) {
      const double* bt = p.beta + t * p.d;
      double q = 0.0;
      for (int k = 0; k < p.d; ++k) {
        const double df = (Xe[a * p.d + k] - Xe[b * p.d + k]) * bt[k];
        q = fma(df, df, q);
      }
      v += p.scale[t] * exp(-q);
    }
    if (a == b) v += p.eye_sum;
    K[a * ld + b] = v;
  }
  __syncthreads();
  // ---- Cholesky, in place, lower ---------------------------------------------------------------------------------
  double logdet = 0.0;
  bool bad = false;
  for (int j = 0; j < n; ++j) {
    const double djj = K[j * ld + j];
    if (!(djj > 0.0)) bad = true;
    const double ljj = sqrt(djj > 0.0 ? djj : 1.0);
    logdet += 2.0 * log(ljj);
    __syncthreads();
    if (tid == 0) K[j * ld + j] = ljj;
    for (int i = j + 1 + tid; i < n; i += NLL_THREADS) K[i * ld + j] /= ljj;
    __syncthreads();
    const int rem = n - j - 1;                                   // trailing update, lower triangle
    for (int idx = tid; idx < rem * rem; idx += NLL_THREADS) {
      const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
      if (k <= i) K[i * ld + k] -= K[i * ld + j] * K[k * ld + j];
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicOr(p.flags, 1);
  // ---- L^-1 in place (row by row: row i of L^-1 needs rows < i of L^-1 and row i of L) ---------------------------
  for (int i = 0; i < n; ++i) {
    const double lii = K[i * ld + i];
    for (int j = tid; j < i; j += NLL_THREADS) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += K[i * ld + k] * K[k * ld + j];   // K[k][j] already holds L^-1 for k < i
      rowbuf[j] = -s / lii;
    }
    __syncthreads();
    for (int j = tid; j < i; j += NLL_THREADS) K[i * ld + j] = rowbuf[j];
    if (tid == 0) K[i * ld + i] = 1.0 / lii;
    __syncthreads();
  }
  // ---- K^-1 = L^-T L^-1 in place (row i needs rows >= i of L^-1; rows are finalised top-down) -------------------
  for (int i = 0; i < n; ++i) {
    for (int j = tid; j <= i; j += NLL_THREADS) {
      double s = 0.0;
      for (int k = i; k < n; ++k) s += K[k * ld + i] * K[k * ld + j];
      rowbuf[j] = s;
    }
    __syncthreads();
    for (int j = tid; j <= i; j += NLL_THREADS) K[i * ld + j] = rowbuf[j];
    __syncthreads();
  }
  // ---- alpha = K^-1 y ;  nll ----------------------------------------------------------------------------------------
  for (int a = tid; a < n; a += NLL_THREADS) {
    double s = 0.0;
    for (int b = 0; b < n; ++b) s += ((b <= a) ? K[a * ld + b] : K[b * ld + a]) * yv[b];
    alpha[a] = s;
  }
  __syncthreads();
  double part = 0.0;
  for (int a = tid; a < n; a += NLL_THREADS) part += yv[a] * alpha[a];
  const double yay = block_sum(part, red);
  double* out = p.out + static_cast<size_t>(e) * (1 + p.n_hypers);
  if (tid == 0) out[0] = 0.5 * yay + 0.5 * logdet;
  // ---- gradient: g_i = sum_ab dK_i[a,b] W_ab,  W = alpha alpha^T - K^-1  (out = -1/2 g) ---------------------------
  for (int i = tid; i < p.n_hypers; i += NLL_THREADS) gacc[i] = 0.0;
  __syncthreads();
  for (int h0 = 0; h0 < p.n_hypers; h0 += 8) {                    // 8 hyper-parameters per sweep over the pairs
    const int hn = (p.n_hypers - h0 < 8) ? (p.n_hypers - h0) : 8;
    double g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int idx = tid; idx < n * n; idx += NLL_THREADS) {
      const int a = idx / n, b = idx % n;
      const double W = alpha[a] * alpha[b] - ((b <= a) ? K[a * ld + b] : K[b * ld + a]);
      double kt[kMaxTerms], sq[kMaxTerms];
      for (int t = 0; t < p.n_terms; ++t) {
        const double* bt = p.beta + t * p.d;
        double q = 0.0, s2 = 0.0;
        for (int k = 0; k < p.d; ++k) {
          const double dx = Xe[a * p.d + k] - Xe[b * p.d + k];
          const double df = dx * bt[k];
          q = fma(df, df, q);
          s2 = fma(dx, dx, s2);
        }
        kt[t] = exp(-q);
        sq[t] = s2;
      }
      for (int hh = 0; hh < hn; ++hh) {
        const int i = h0 + hh;
        const int kind = p.h_kind[i];
        double dk;
        if (kind == 0) {                                               // trainable scalar above a sub-tree
          const double* cf = p.h_coef + static_cast<size_t>(i) * (kMaxTerms + 1);
          dk = (a == b) ? cf[kMaxTerms] : 0.0;
          for (int t = 0; t < p.n_terms; ++t) dk += cf[t] * kt[t];
        } else if (kind == 1) {                                        // ARD beta_k: -2 beta_k dx_k^2 * C k   (ARDRBFKernel.scala:61-79)
          const int t = p.h_term[i], k = p.h_dim[i];
          const double dx = Xe[a * p.d + k] - Xe[b * p.d + k];
          dk = p.scale[t] * (-2.0 * p.h_value[i] * dx * dx) * kt[t];
        } else {                                                       // RBF sigma: sqdist * k / sigma^3   (RBFKernel.scala:56-64)
          const int t = p.h_term[i];
          const double sg = p.h_value[i];
          dk = p.scale[t] * sq[t] * kt[t] / (sg * sg * sg);
        }
        g[hh] = fma(dk, W, g[hh]);
      }
    }
    for (int hh = 0; hh < hn; ++hh) {
      const double s = block_sum(g[hh], red);
      if (tid == 0) out[1 + h0 + hh] = -0.5 * s;
    }
  }
}

// sum the per-expert rows in a fixed order: deterministic
__global__ void nll_reduce_kernel(double* __restrict__ total, const double* __restrict__ per_expert, long long E, int width) {
  const int c = threadIdx.x;
  if (c >= width) return;
  double s = 0.0;
  for (long long e = 0; e < E; ++e) s += per_expert[e * width + c];
  total[c] = s;
}

}  // namespace

size_t bcm_nll_smem_bytes(int n_max) {
  return sizeof(double) * (static_cast<size_t>(n_max) * (n_max + 1) + 3 * static_cast<size_t>(n_max) + 8 + MAX_HYPERS);
}
int bcm_nll_max_hypers() { return MAX_HYPERS; }

cudaError_t launch_bcm_nll(const double* dX, const double* dy, const long long* dOff, long long E, int d, int n_max,
                           const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind, const int* dTerm,
                           const int* dDim, const double* dCoef, const double* dValue, double* dPerExpert,
                           double* dTotal, int* dFlags, cudaStream_t s) {
  NllParams p{};
  p.X = dX; p.y = dy; p.off = dOff; p.d = d; p.n_max = n_max;
  p.n_terms = kf.n_terms;
  for (int t = 0; t < kMaxTerms; ++t) p.scale[t] = kf.scale[t];
  p.beta = dBeta; p.eye_sum = kf.eye_sum;
  p.n_hypers = n_hypers; p.h_kind = dKind; p.h_term = dTerm; p.h_dim = dDim; p.h_coef = dCoef; p.h_value = dValue;
  p.out = dPerExpert; p.flags = dFlags;
  const size_t smem = bcm_nll_smem_bytes(n_max);
  cudaError_t e = cudaFuncSetAttribute(bcm_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  bcm_nll_kernel<<<static_cast<unsigned>(E), NLL_THREADS, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  nll_reduce_kernel<<<1, 128, 0, s>>>(dTotal, dPerExpert, E, 1 + n_hypers);
  return cudaGetLastError();
}

}  // namespace sgp
