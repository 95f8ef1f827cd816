// Batched per-expert Laplace approximation for binary GP classification (fp64), one CTA per expert.
//
// Replaces, for all experts of a rank at once, classification/GaussianProcessClassifier.scala:74-129
// `likelihoodAndGradient`: the Newton iteration for the mode of q(f | X, y) (Rasmussen & Williams Alg. 3.1) WITH the
// reference's step halving and its warm start (f persists across objective evaluations, GPCls:53-55,105), then the
// approximate log marginal likelihood and its gradient (R&W Alg. 5.1) -- including the reference's quirks: the
// post-loop quantities (pi, W, L, a, grad log p) are the ones of the LAST Newton iteration, i.e. computed from the f
// before that iteration's update, while d3logP (GPCls:118, sign as in the reference) uses exp(-f) of the final f.
// Summed over experts like commons/GaussianProcessCommons.scala:73-78.
#include "sgp_internal.h"

namespace sgp {
namespace {

constexpr int LT = 256;            // threads per CTA

struct LapParams {
  const double* X; const double* y; double* f; const long long* off;
  int d, n_max, n_terms;
  double scale[kMaxTerms];
  const double* beta;        // [n_terms][d]
  double eye_sum;
  int n_hypers;
  const int* h_kind; const int* h_term; const int* h_dim;
  const double* h_coef;      // [n_hypers][kMaxTerms+1]
  const double* h_value;
  double tol;
  double* out;               // [E][1 + n_hypers]  (-logZ, -gradLogZ)
  int* flags;                // bit 0: Cholesky pivot not positive
};

__device__ __forceinline__ double bsum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < LT / 32; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// in-place lower Cholesky of the n x n matrix M (leading dimension ld); returns sum(log(diag L)) and a bad-pivot flag
__device__ double chol_inplace(double* M, int n, int ld, bool& bad) {
  double sumlog = 0.0;
  for (int j = 0; j < n; ++j) {
    const double djj = M[j * ld + j];
    if (!(djj > 0.0)) bad = true;
    const double ljj = sqrt(djj > 0.0 ? djj : 1.0);
    sumlog += log(ljj);
    __syncthreads();
    if (threadIdx.x == 0) M[j * ld + j] = ljj;
    for (int i = j + 1 + threadIdx.x; i < n; i += LT) M[i * ld + j] /= ljj;
    __syncthreads();
    const int rem = n - j - 1;
    for (int idx = threadIdx.x; idx < rem * rem; idx += LT) {
      const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
      if (k <= i) M[i * ld + k] -= M[i * ld + j] * M[k * ld + j];
    }
    __syncthreads();
  }
  return sumlog;
}
// x <- L^-1 x  (forward substitution, column oriented)
__device__ void fwd_solve(const double* L, int n, int ld, double* x) {
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    if (threadIdx.x == 0) x[j] /= L[j * ld + j];
    __syncthreads();
    const double xj = x[j];
    for (int i = j + 1 + threadIdx.x; i < n; i += LT) x[i] -= L[i * ld + j] * xj;
  }
  __syncthreads();
}
// x <- L^-T x  (backward substitution)
__device__ void bwd_solve(const double* L, int n, int ld, double* x) {
  for (int j = n - 1; j >= 0; --j) {
    __syncthreads();
    if (threadIdx.x == 0) x[j] /= L[j * ld + j];
    __syncthreads();
    const double xj = x[j];
    for (int i = threadIdx.x; i < j; i += LT) x[i] -= L[j * ld + i] * xj;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LT, 1) laplace_kernel(const LapParams p) {
  extern __shared__ double sm[];
  const long long e = blockIdx.x;
  const long long r0 = p.off[e];
  const int n = static_cast<int>(p.off[e + 1] - r0);
  const int ld = p.n_max + 1;
  double* K = sm;                                        // full symmetric
  double* M = K + static_cast<size_t>(p.n_max) * ld;     // B -> L -> L^-1 -> B^-1 (lower)
  double* v = M + static_cast<size_t>(p.n_max) * ld;     // vectors, n_max each
  double *f = v, *yv = v + p.n_max, *pi = v + 2 * p.n_max, *sw = v + 3 * p.n_max, *glp = v + 4 * p.n_max,
         *bv = v + 5 * p.n_max, *av = v + 6 * p.n_max, *nf = v + 7 * p.n_max, *t1 = v + 8 * p.n_max,
         *t2 = v + 9 * p.n_max, *s2 = v + 10 * p.n_max, *wv = v + 11 * p.n_max;
  double* red = v + 12 * p.n_max;                        // [8]
  const int tid = threadIdx.x;
  const double* Xe = p.X + static_cast<size_t>(r0) * p.d;

  for (int i = tid; i < n; i += LT) { f[i] = p.f[r0 + i]; yv[i] = p.y[r0 + i]; }
  for (int idx = tid; idx < n * n; idx += LT) {          // K, both triangles
    const int a = idx / n, b = idx % n;
    if (b > a) continue;
    double val = 0.0;
    for (int t = 0; t < p.n_terms; ++t) {
      const double* bt = p.beta + t * p.d;
      double q = 0.0;
      for (int k = 0; k < p.d; ++k) {
        const double df = (Xe[a * p.d + k] - Xe[b * p.d + k]) * bt[k];
        q = fma(df, df, q);
      }
      val += p.scale[t] * exp(-q);
    }
    if (a == b) val += p.eye_sum;
    K[a * ld + b] = val;
    K[b * ld + a] = val;
  }
  __syncthreads();

  // ---- Newton iteration for the mode (GPCls:79-111) ------------------------------------------------------------------
  double old_obj = -INFINITY, new_obj = -1.7976931348623157e308;   // Double.NegativeInfinity, Double.MinValue
  double step = 1.0, sumlogL = 0.0;
  bool bad = false;
  while (fabs(old_obj - new_obj) > p.tol && step > p.tol) {
    for (int i = tid; i < n; i += LT) {
      const double s = sigmoid(f[i]);
      pi[i] = s;
      wv[i] = s * (1.0 - s);
      sw[i] = sqrt(wv[i]);
      glp[i] = yv[i] - s;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += LT) {        // B = I + sqrtW K sqrtW   (GPCls:95-97), lower
      const int a = idx / n, b = idx % n;
      if (b <= a) M[a * ld + b] = sw[a] * sw[b] * K[a * ld + b] + (a == b ? 1.0 : 0.0);
    }
    __syncthreads();
    sumlogL = chol_inplace(M, n, ld, bad);               // GPCls:98
    for (int i = tid; i < n; i += LT) bv[i] = wv[i] * f[i] + glp[i];           // b = W f + grad log p   (:100)
    __syncthreads();
    for (int i = tid; i < n; i += LT) {                  // rhs = sqrtW (K b)
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += K[i * ld + k] * bv[k];
      t1[i] = sw[i] * s;
    }
    fwd_solve(M, n, ld, t1);                             // L \ rhs
    bwd_solve(M, n, ld, t1);                             // L^T \ .
    for (int i = tid; i < n; i += LT) av[i] = bv[i] - sw[i] * t1[i];           // a   (:101)
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < n; i += LT) {                  // f_candidate = (1-step) f + step K a   (:102)
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += K[i * ld + k] * av[k];
      const double c = (1.0 - step) * f[i] + step * s;
      nf[i] = c;
      part += -0.5 * av[i] * c + log(sigmoid((yv[i] * 2.0 - 1.0) * c));        // (:103)
    }
    const double cand = bsum(part, red);
    if (cand > old_obj) {                                // (:104-107)
      for (int i = tid; i < n; i += LT) f[i] = nf[i];
      old_obj = new_obj;
      new_obj = cand;
    } else {
      step *= 0.5;
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicOr(p.flags, 1);
  for (int i = tid; i < n; i += LT) p.f[r0 + i] = f[i];   // the mode persists (warm start of the next evaluation)

  // ---- log Z and gradient (GPCls:114-128) --------------------------------------------------------------------------
  double* out = p.out + static_cast<size_t>(e) * (1 + p.n_hypers);
  if (tid == 0) out[0] = -(new_obj - sumlogL);
  // L^-1 in place
  for (int i = 0; i < n; ++i) {
    const double lii = M[i * ld + i];
    for (int j = tid; j < i; j += LT) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += M[i * ld + k] * M[k * ld + j];
      t1[j] = -s / lii;
    }
    __syncthreads();
    for (int j = tid; j < i; j += LT) M[i * ld + j] = t1[j];
    if (tid == 0) M[i * ld + i] = 1.0 / lii;
    __syncthreads();
  }
  // diag(C^T C), C = L^-1 (sqrtW K):  cc_i = | L^-1 (sw o K[:,i]) |^2   (GPCls:117,119)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < n; i += LT / 32) {
      double acc = 0.0;
      for (int r = lane; r < n; r += 32) {
        double s = 0.0;
        for (int k = 0; k <= r; ++k) s += M[r * ld + k] * sw[k] * K[k * ld + i];
        acc += s * s;
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) t2[i] = acc;
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += LT) {
    const double d3 = -(2.0 * pi[i] - 1.0) * pi[i] * pi[i] * exp(-f[i]);        // (:118) sign as in the reference
    s2[i] = -0.5 * (K[i * ld + i] - t2[i]) * d3;                                // (:119)
  }
  // B^-1 = L^-T L^-1 in place (lower);  R = sqrtW B^-1 sqrtW   (GPCls:116)
  for (int i = 0; i < n; ++i) {
    for (int j = tid; j <= i; j += LT) {
      double s = 0.0;
      for (int k = i; k < n; ++k) s += M[k * ld + i] * M[k * ld + j];
      t1[j] = s;
    }
    __syncthreads();
    for (int j = tid; j <= i; j += LT) M[i * ld + j] = t1[j];
    __syncthreads();
  }
  auto Rab = [&](int a, int b) { return sw[a] * sw[b] * ((b <= a) ? M[a * ld + b] : M[b * ld + a]); };

  for (int h = 0; h < p.n_hypers; ++h) {                 // (GPCls:121-126), one derivative matrix at a time
    const int kind = p.h_kind[h];
    double q1 = 0.0, q2 = 0.0;
    for (int a = tid; a < n; a += LT) {                  // one row per thread: bb[a] = (dK glp)[a], partial sums
      double bb = 0.0;
      for (int b = 0; b < n; ++b) {
        double dk;
        if (kind == 0) {
          const double* cf = p.h_coef + static_cast<size_t>(h) * (kMaxTerms + 1);
          dk = (a == b) ? cf[kMaxTerms] : 0.0;
          for (int t = 0; t < p.n_terms; ++t) {
            if (cf[t] == 0.0) continue;
            const double* bt = p.beta + t * p.d;
            double q = 0.0;
            for (int k = 0; k < p.d; ++k) {
              const double df = (Xe[a * p.d + k] - Xe[b * p.d + k]) * bt[k];
              q = fma(df, df, q);
            }
            dk += cf[t] * exp(-q);
          }
        } else {
          const int t = p.h_term[h];
          const double* bt = p.beta + t * p.d;
          double q = 0.0, sq = 0.0;
          for (int k = 0; k < p.d; ++k) {
            const double dx = Xe[a * p.d + k] - Xe[b * p.d + k];
            const double df = dx * bt[k];
            q = fma(df, df, q);
            sq = fma(dx, dx, sq);
          }
          const double kt = exp(-q);
          if (kind == 1) {
            const double dx = Xe[a * p.d + p.h_dim[h]] - Xe[b * p.d + p.h_dim[h]];
            dk = p.scale[t] * (-2.0 * p.h_value[h] * dx * dx) * kt;
          } else {
            const double sg = p.h_value[h];
            dk = p.scale[t] * sq * kt / (sg * sg * sg);
          }
        }
        bb = fma(dk, glp[b], bb);
        q1 = fma(av[a] * dk, av[b], q1);
        q2 = fma(Rab(a, b), dk, q2);
      }
      t1[a] = bb;
    }
    __syncthreads();
    for (int a = tid; a < n; a += LT) {                  // t2 = R bb
      double s = 0.0;
      for (int b = 0; b < n; ++b) s += Rab(a, b) * t1[b];
      t2[a] = s;
    }
    __syncthreads();
    double part = 0.5 * q1 - 0.5 * q2;                   // s1 pieces
    for (int a = tid; a < n; a += LT) {                  // s3 = bb - K (R bb);  s2 . s3
      double s = 0.0;
      for (int b = 0; b < n; ++b) s += K[a * ld + b] * t2[b];
      part += s2[a] * (t1[a] - s);
    }
    const double g = bsum(part, red);
    if (tid == 0) out[1 + h] = -g;                       // -gradLogZ
    __syncthreads();
  }
}

__global__ void lap_reduce_kernel(double* __restrict__ total, const double* __restrict__ per_expert, long long E, int width) {
  const int c = threadIdx.x;
  if (c >= width) return;
  double s = 0.0;
  for (long long e = 0; e < E; ++e) s += per_expert[e * width + c];
  total[c] = s;
}

}  // namespace

size_t laplace_smem_bytes(int n_max) {
  return sizeof(double) * (2 * static_cast<size_t>(n_max) * (n_max + 1) + 12 * static_cast<size_t>(n_max) + 8);
}

cudaError_t launch_laplace(const double* dX, const double* dy, double* df, const long long* dOff, long long E, int d,
                           int n_max, const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind,
                           const int* dTerm, const int* dDim, const double* dCoef, const double* dValue, double tol,
                           double* dPerExpert, double* dTotal, int* dFlags, cudaStream_t s) {
  LapParams p{};
  p.X = dX; p.y = dy; p.f = df; p.off = dOff; p.d = d; p.n_max = n_max; p.n_terms = kf.n_terms;
  for (int t = 0; t < kMaxTerms; ++t) p.scale[t] = kf.scale[t];
  p.beta = dBeta; p.eye_sum = kf.eye_sum;
  p.n_hypers = n_hypers; p.h_kind = dKind; p.h_term = dTerm; p.h_dim = dDim; p.h_coef = dCoef; p.h_value = dValue;
  p.tol = tol; p.out = dPerExpert; p.flags = dFlags;
  const size_t smem = laplace_smem_bytes(n_max);
  cudaError_t e = cudaFuncSetAttribute(laplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  laplace_kernel<<<static_cast<unsigned>(E), LT, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  lap_reduce_kernel<<<1, 128, 0, s>>>(dTotal, dPerExpert, E, 1 + n_hypers);
  return cudaGetLastError();
}

}  // namespace sgp
