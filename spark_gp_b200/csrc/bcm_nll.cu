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
  int x_in_smem;            // the expert's rows are staged in shared memory
  int any_ard;              // at least one ARD_BETA hyper-parameter (needs the per-dimension sums)
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

// Thread mapping used by the triangular sweeps: warp w owns rows a = w, w + 8, ...; its lanes own the columns
// b = lane, lane + 32, ... <= a.  No integer division, coalesced / conflict-free row accesses.
constexpr int NLL_WARPS = NLL_THREADS / 32;
constexpr int DCH = 16;             // ARD dimensions handled per sweep over the pairs (register accumulators)

__global__ void __launch_bounds__(NLL_THREADS, 2) bcm_nll_kernel(const NllParams p) {
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
  double* sums = red + 8;                     // [2 kMaxTerms + 1 + DCH]  block-reduced sufficient sums of one sweep
  double* Xs = sums + 2 * kMaxTerms + 1 + DCH;           // [n_max][xld] staged copy of the expert's rows (if it fits)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* Xg = p.X + static_cast<size_t>(r0) * p.d;
  const double* Xe = Xg;
  int xld = p.d;
  if (p.x_in_smem) {
    xld = p.d | 1;                            // odd stride: rows a, a+1, ... hit different banks
    for (int idx = tid; idx < n * p.d; idx += NLL_THREADS) Xs[(idx / p.d) * xld + idx % p.d] = Xg[idx];
    Xe = Xs;
  }
  for (int i = tid; i < n; i += NLL_THREADS) yv[i] = p.y[r0 + i];
  __syncthreads();
  // ---- K (lower triangle incl. diagonal) -------------------------------------------------------------------------------
  for (int a = warp; a < n; a += NLL_WARPS) {
    for (int b = lane; b <= a; b += 32) {
      double v = 0.0;
      for (int t = 0; t < p.n_terms; ++t) {
        const double* bt = p.beta + t * p.d;
        double q = 0.0;
        for (int k = 0; k < p.d; ++k) {
          const double df = (Xe[a * xld + k] - Xe[b * xld + k]) * bt[k];
          q = fma(df, df, q);
        }
        v += p.scale[t] * exp(-q);
      }
      if (a == b) v += p.eye_sum;
      K[a * ld + b] = v;
    }
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
    const double inv = 1.0 / ljj;
    for (int i = j + 1 + tid; i < n; i += NLL_THREADS) rowbuf[i] = K[i * ld + j] * inv;
    __syncthreads();
    for (int i = j + 1 + warp; i < n; i += NLL_WARPS) {                // trailing update, lower triangle
      const double lij = rowbuf[i];
      for (int k = j + 1 + lane; k <= i; k += 32) K[i * ld + k] = fma(-lij, rowbuf[k], K[i * ld + k]);
      if (lane == 0) K[i * ld + j] = lij;
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicOr(p.flags, 1);
  // ---- L^-1 in place (row by row: row i of L^-1 needs rows < i of L^-1 and row i of L) ---------------------------
  for (int i = 0; i < n; ++i) {
    const double lii = K[i * ld + i];
    for (int j = tid; j < i; j += NLL_THREADS) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s = fma(K[i * ld + k], K[k * ld + j], s);   // K[k][j] already holds L^-1 for k < i
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
      for (int k = i; k < n; ++k) s = fma(K[k * ld + i], K[k * ld + j], s);
      rowbuf[j] = s;
    }
    __syncthreads();
    for (int j = tid; j <= i; j += NLL_THREADS) K[i * ld + j] = rowbuf[j];
    __syncthreads();
  }
  // ---- alpha = K^-1 y ;  nll ----------------------------------------------------------------------------------------
  for (int a = tid; a < n; a += NLL_THREADS) {
    double s = 0.0;
    for (int b = 0; b < n; ++b) s = fma((b <= a) ? K[a * ld + b] : K[b * ld + a], yv[b], s);
    alpha[a] = s;
  }
  __syncthreads();
  double part = 0.0;
  for (int a = tid; a < n; a += NLL_THREADS) part += yv[a] * alpha[a];
  const double yay = block_sum(part, red);
  double* out = p.out + static_cast<size_t>(e) * (1 + p.n_hypers);
  if (tid == 0) out[0] = 0.5 * yay + 0.5 * logdet;
  // ---- gradient: g_i = sum_ab dK_i[a,b] W_ab,  W = alpha alpha^T - K^-1  (out = -1/2 g) ---------------------------
  // Every derivative the DSL can produce is a combination of a few sums over the pairs, per non-Eye term t:
  //   S_t = sum k_t W           Q_t = sum |x_a - x_b|^2 k_t W          D_tk = sum (x_ak - x_bk)^2 k_t W        trW
  //   SCALE     (ScalarTimesKernel.scala:50-54,93-97): sum_t coef_t S_t + coef_eye trW
  //   ARD_BETA  (ARDRBFKernel.scala:61-79)           : -2 beta_k scale_t D_tk
  //   RBF_SIGMA (RBFKernel.scala:56-64)              : scale_t Q_t / sigma^3
  // W and dK are symmetric: one sweep over the lower triangle, off-diagonal pairs weighted 2 (their dx is nonzero only
  // there).  ARD dimensions go through register accumulators DCH at a time; the first sweep also yields S, Q, trW.
  const int chunks = (p.d + DCH - 1) / DCH;
  const int n_sweeps = p.any_ard ? p.n_terms * chunks : 1;
  double* sS = sums;                          // [kMaxTerms]
  double* sQ = sums + kMaxTerms;              // [kMaxTerms]
  double* sTr = sums + 2 * kMaxTerms;         // [1]
  double* sD = sTr + 1;                       // [DCH]  per-dimension sums of this sweep's (term, dimension chunk)
  for (int sw = 0; sw < n_sweeps; ++sw) {
    const int ts = p.any_ard ? sw / chunks : -1;
    const int k0 = p.any_ard ? (sw % chunks) * DCH : 0;
    const int kn = p.any_ard ? ((p.d - k0 < DCH) ? (p.d - k0) : DCH) : 0;
    double S[kMaxTerms], Q[kMaxTerms], D[DCH], trW = 0.0;
#pragma unroll
    for (int t = 0; t < kMaxTerms; ++t) { S[t] = 0.0; Q[t] = 0.0; }
#pragma unroll
    for (int k = 0; k < DCH; ++k) D[k] = 0.0;
    for (int a = warp; a < n; a += NLL_WARPS) {
      const double al = alpha[a];
      for (int b = lane; b <= a; b += 32) {
        const double W = ((a == b) ? 1.0 : 2.0) * (al * alpha[b] - K[a * ld + b]);
        if (a == b) trW += W;
        double kws = 0.0;
#pragma unroll
        for (int t = 0; t < kMaxTerms; ++t) {
          if (t < p.n_terms && (sw == 0 || t == ts)) {
            const double* bt = p.beta + t * p.d;
            double q = 0.0, s2 = 0.0;
            for (int k = 0; k < p.d; ++k) {
              const double dx = Xe[a * xld + k] - Xe[b * xld + k];
              const double df = dx * bt[k];
              q = fma(df, df, q);
              s2 = fma(dx, dx, s2);
            }
            const double kw = exp(-q) * W;
            if (sw == 0) { S[t] += kw; Q[t] = fma(s2, kw, Q[t]); }
            if (t == ts) kws = kw;
          }
        }
#pragma unroll
        for (int k = 0; k < DCH; ++k) {
          if (k < kn) {
            const double dx = Xe[a * xld + k0 + k] - Xe[b * xld + k0 + k];
            D[k] = fma(dx * dx, kws, D[k]);
          }
        }
      }
    }
    // block-reduce into shared memory (fixed order: deterministic)
    if (sw == 0) {
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t) {
        if (t < p.n_terms) {
          const double s_ = block_sum(S[t], red), q_ = block_sum(Q[t], red);
          if (tid == 0) { sS[t] = s_; sQ[t] = q_; }
        }
      }
      const double v = block_sum(trW, red);
      if (tid == 0) sTr[0] = v;
    }
#pragma unroll
    for (int k = 0; k < DCH; ++k) {
      if (k < kn) {
        const double v = block_sum(D[k], red);
        if (tid == 0) sD[k] = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < p.n_hypers; i += NLL_THREADS) {
      const int kind = p.h_kind[i];
      double g = 0.0;
      bool mine = (sw == 0);
      if (kind == 0) {                                                 // trainable scalar above a sub-tree
        const double* cf = p.h_coef + static_cast<size_t>(i) * (kMaxTerms + 1);
        g = cf[kMaxTerms] * sTr[0];
        for (int t = 0; t < p.n_terms; ++t) g = fma(cf[t], sS[t], g);
      } else if (kind == 1) {                                          // ARD beta_k
        const int t = p.h_term[i], k = p.h_dim[i];
        mine = (t == ts && k >= k0 && k < k0 + kn);
        if (mine) g = p.scale[t] * (-2.0 * p.h_value[i]) * sD[k - k0];
      } else {                                                         // RBF sigma
        const int t = p.h_term[i];
        const double sg = p.h_value[i];
        g = p.scale[t] * sQ[t] / (sg * sg * sg);
      }
      if (mine) out[1 + i] = -0.5 * g;
    }
    __syncthreads();
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

static size_t nll_base_doubles(int n_max) {
  return static_cast<size_t>(n_max) * (n_max + 1) + 3 * static_cast<size_t>(n_max) + 8 + 2 * kMaxTerms + 1 + DCH;
}
size_t bcm_nll_smem_bytes(int n_max) { return sizeof(double) * nll_base_doubles(n_max); }
int bcm_nll_max_hypers() { return MAX_HYPERS; }

cudaError_t launch_bcm_nll(const double* dX, const double* dy, const long long* dOff, long long E, int d, int n_max,
                           const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind, const int* dTerm,
                           const int* dDim, const double* dCoef, const double* dValue, int any_ard,
                           double* dPerExpert, double* dTotal, int* dFlags, cudaStream_t s) {
  NllParams p{};
  p.X = dX; p.y = dy; p.off = dOff; p.d = d; p.n_max = n_max;
  p.n_terms = kf.n_terms;
  for (int t = 0; t < kMaxTerms; ++t) p.scale[t] = kf.scale[t];
  p.beta = dBeta; p.eye_sum = kf.eye_sum;
  p.n_hypers = n_hypers; p.h_kind = dKind; p.h_term = dTerm; p.h_dim = dDim; p.h_coef = dCoef; p.h_value = dValue;
  p.out = dPerExpert; p.flags = dFlags;
  p.any_ard = any_ard;
  // stage the expert's rows in shared memory when two CTAs per SM still fit (113 KB each)
  size_t smem = bcm_nll_smem_bytes(n_max);
  const size_t with_x = smem + sizeof(double) * static_cast<size_t>(n_max) * (d | 1);
  p.x_in_smem = (with_x <= 113 * 1024) ? 1 : 0;
  if (p.x_in_smem) smem = with_x;
  cudaError_t e = cudaFuncSetAttribute(bcm_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  bcm_nll_kernel<<<static_cast<unsigned>(E), NLL_THREADS, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  nll_reduce_kernel<<<1, 128, 0, s>>>(dTotal, dPerExpert, E, 1 + n_hypers);
  return cudaGetLastError();
}

}  // namespace sgp
