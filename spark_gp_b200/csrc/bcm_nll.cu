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
#include <cstdlib>
#include <string>

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

// ---------------------------------------------------------------------------------------------------------------------
// Register-resident variant (one non-Eye term, experts of <= 128 points: every default configuration).
// The shared-memory kernel above spends its time in ~600 block barriers and latency-bound shared-memory updates
// (Cholesky, triangular inverse, L^-T L^-1: 10.3 ms per evaluation of 10^4 experts of 100 points, ~5 % of the fp64 rate).
// Here the matrix lives in REGISTERS: 512 threads form a 32 x 16 grid, thread (ty, tx) owns the entries
// (ty + 32 r, tx + 16 c) (cyclic: the work stays balanced while pivots move through the matrix; 16 warps hide the fp64
// and barrier latencies -- a 16 x 16 grid of 256 threads ran at 36 % issue utilisation, 5.1 ms), and K^-1 and
// log|det K| come from n SWEEPS (Goodnight's sweep operator = Gauss-Jordan on the symmetric matrix, no pivoting needed
// for an SPD matrix): sweep k broadcasts column k through shared memory (ONE barrier) and every thread applies the
// rank-1 update to its RB x RB block from registers; the pivots are the Schur complements (> 0 iff K is positive
// definite, their logs sum to log|det K|), and after n sweeps the registers hold -K^-1.
// The gradient needs no second exp(): the unscaled kernel values k_ab are parked in shared memory by the build, and with
// M = k o W (W = alpha alpha^T - K^-1) the per-dimension sums are quadratic forms,
//   D_k = sum_ab (x_ak - x_bk)^2 M_ab = 2 (sum_a x_ak^2 m_a - x_k^T M x_k),  m = M 1,
// evaluated from registers (the expert's rows are translated by its first row when staged, so the expansion does not
// cancel).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int REG_GY = 32, REG_THREADS = REG_GY * 16, REG_WARPS = REG_THREADS / 32;   // 32 x 16 thread grid
template <int RBR, int RBC>      // rows / columns per thread: thread (ty, tx) owns entries (ty + 32 r, tx + 16 c)
__global__ void __launch_bounds__(REG_THREADS, 1) bcm_nll_reg_kernel(const NllParams p) {
  static_assert(RBC == 2 * RBR || RBC == 2 * RBR - 1, "the column extent covers the row extent");
  constexpr int NP = 16 * RBC;                // padded order of the matrix (rows beyond NP are never touched)
  constexpr int NR = 32 * RBR;                // >= NP
  extern __shared__ double sm[];
  const long long e = blockIdx.x;
  const long long r0 = p.off[e];
  const int n = static_cast<int>(p.off[e + 1] - r0);
  const int d = p.hv.d, xld = d | 1;
  double* ys = sm;                            // [NR]
  double* alpha = ys + NR;                    // [NR]
  double* colbuf = alpha + NR;                // [2][NR]  column of the current sweep (double buffered by sweep parity)
  double* piv = colbuf + 2 * NR;              // [NR]     pivots
  double* red = piv + NR;                     // [16]
  double* sums = red + 16;                    // [EX_SUMS]
  double* Xs = sums + EX_SUMS;                // [NR][xld]  rows of the expert, minus its first row; zero padded
  double* Ks = Xs + static_cast<size_t>(NR) * xld;   // [NR][NP]  exponents, then unscaled kernel values (0 on padding)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  {
    const double* Xg = p.X + static_cast<size_t>(r0) * d;
    for (int idx = tid; idx < NR * d; idx += REG_THREADS) {
      const int i = idx / d, k = idx % d;
      Xs[i * xld + k] = (i < n) ? Xg[idx] - Xg[k] : 0.0;
    }
    for (int i = tid; i < NR; i += REG_THREADS) ys[i] = (i < n) ? p.y[r0 + i] : 0.0;
  }
  __syncthreads();
  // ---- build: A = scale exp(-sum_k ((x_ak - x_bk) beta_k)^2) + eye_sum I; identity on the padding -----------------------
  double A[RBR][RBC];
#pragma unroll
  for (int r = 0; r < RBR; ++r)
#pragma unroll
    for (int c = 0; c < RBC; ++c) A[r][c] = 0.0;
  for (int k = 0; k < d; ++k) {
    const double bk = p.hv.beta[k];
    double xa[RBR], xb[RBC];
#pragma unroll
    for (int r = 0; r < RBR; ++r) xa[r] = Xs[(ty + 32 * r) * xld + k] * bk;
#pragma unroll
    for (int c = 0; c < RBC; ++c) xb[c] = Xs[(tx + 16 * c) * xld + k] * bk;
#pragma unroll
    for (int r = 0; r < RBR; ++r)
#pragma unroll
      for (int c = 0; c < RBC; ++c) {
        const double df = xa[r] - xb[c];
        A[r][c] = fma(df, df, A[r][c]);
      }
  }
  // exp() through shared memory in a rolled loop: 28-56 inlined fp64 exp() are ~30 KB of straight-line code per thread
#pragma unroll
  for (int r = 0; r < RBR; ++r)
#pragma unroll
    for (int c = 0; c < RBC; ++c) Ks[(ty + 32 * r) * NP + tx + 16 * c] = A[r][c];
  __syncwarp();                               // a thread reads back only its own entries
#pragma unroll 1
  for (int rc = 0; rc < RBR * RBC; ++rc) {
    const int r = rc / RBC, c = rc % RBC;
    const int a = ty + 32 * r, b = tx + 16 * c;
    double* kp = Ks + a * NP + b;
    *kp = (a < n && b < n) ? exp(-*kp) : 0.0;
  }
  const double scale0 = p.hv.scale[0];
#pragma unroll
  for (int r = 0; r < RBR; ++r)
#pragma unroll
    for (int c = 0; c < RBC; ++c) {
      const int a = ty + 32 * r, b = tx + 16 * c;
      const bool real = a < n && b < n;
      A[r][c] = real ? fma(scale0, Ks[a * NP + b], (a == b) ? p.hv.eye_sum : 0.0) : ((a == b) ? 1.0 : 0.0);
    }
  // ---- n sweeps: A -> -K^-1.  Pivot k = 16 ck + kx sits in row block rk = ck / 2 (ty == 16 (ck % 2) + kx) ------------------
  bool bad = false;
#pragma unroll
  for (int ck = 0; ck < RBC; ++ck) {
    constexpr int dummy = 0; (void)dummy;
    const int rk = ck >> 1;                   // static after unrolling
    for (int kx = 0; kx < 16; ++kx) {
      const int k = kx + 16 * ck;
      if (k >= n) break;
      const int ky = 16 * (ck & 1) + kx;
      double* cb = colbuf + (k & 1) * NR;
      if (tx == kx) {
#pragma unroll
        for (int r = 0; r < RBR; ++r) cb[ty + 32 * r] = A[r][ck];
      }
      __syncthreads();
      const double pv = cb[k];
      if (!(pv > 0.0)) bad = true;
      const double pinv = __drcp_rn((pv > 0.0) ? pv : 1.0);
      if (tid == 0) piv[k] = (pv > 0.0) ? pv : 1.0;
      double ci[RBR], cj[RBC];
#pragma unroll
      for (int r = 0; r < RBR; ++r) ci[r] = cb[ty + 32 * r];
#pragma unroll
      for (int c = 0; c < RBC; ++c) cj[c] = cb[tx + 16 * c] * pinv;
      const bool rowk = (ty == ky), colk = (tx == kx);
      // the generic rank-1 update everywhere, then the pivot row / column are overwritten by their owners (selecting per
      // element cost 70 of the 140 instructions of a sweep)
#pragma unroll
      for (int r = 0; r < RBR; ++r)
#pragma unroll
        for (int c = 0; c < RBC; ++c) A[r][c] = fma(-ci[r], cj[c], A[r][c]);
      if (colk) {                             // column k: A_ik / p
#pragma unroll
        for (int r = 0; r < RBR; ++r) A[r][ck] = ci[r] * pinv;
      }
      if (rowk) {                             // row k: A_kj / p, and -1/p on the diagonal
#pragma unroll
        for (int c = 0; c < RBC; ++c) A[rk][c] = cj[c];
        if (colk) A[rk][ck] = -pinv;
      }
    }
  }
  if (bad && tid == 0) atomicOr(p.flags, 1);
  __syncthreads();
  // ---- log|det K| = sum log pivot;  alpha = K^-1 y ------------------------------------------------------------------------
  double part = 0.0;
  for (int k = tid; k < n; k += REG_THREADS) part += log(piv[k]);
  const double logdet = ex_block_sum<REG_WARPS>(part, red);
  {
    double yb[RBC];
#pragma unroll
    for (int c = 0; c < RBC; ++c) yb[c] = ys[tx + 16 * c];
#pragma unroll
    for (int r = 0; r < RBR; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < RBC; ++c) s = fma(-A[r][c], yb[c], s);
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // over the 16 threads of this row group
      if (tx == 0) alpha[ty + 32 * r] = (ty + 32 * r < n) ? s : 0.0;
    }
  }
  __syncthreads();
  part = 0.0;
  for (int a = tid; a < n; a += REG_THREADS) part += ys[a] * alpha[a];
  const double yay = ex_block_sum<REG_WARPS>(part, red);
  double* out = p.out + static_cast<size_t>(e) * (1 + p.hv.n_hypers);
  if (tid == 0) out[0] = 0.5 * yay + 0.5 * logdet;
  // ---- gradient: M = k o (alpha alpha^T - K^-1) in place of A --------------------------------------------------------------
  double S[kMaxTerms], Q[kMaxTerms], D[DCH], trW = 0.0, m[RBR];
#pragma unroll
  for (int t = 0; t < kMaxTerms; ++t) { S[t] = 0.0; Q[t] = 0.0; }
  {
    double aa[RBR], ab[RBC];
#pragma unroll
    for (int r = 0; r < RBR; ++r) aa[r] = alpha[ty + 32 * r];
#pragma unroll
    for (int c = 0; c < RBC; ++c) ab[c] = alpha[tx + 16 * c];
#pragma unroll
    for (int r = 0; r < RBR; ++r) {
      double ms = 0.0;
#pragma unroll
      for (int c = 0; c < RBC; ++c) {
        const int a = ty + 32 * r, b = tx + 16 * c;
        const double W = fma(aa[r], ab[c], A[r][c]);             // A = -K^-1
        if (a == b && a < n) trW += W;
        const double Mv = Ks[a * NP + b] * W;                     // 0 on the padding
        A[r][c] = Mv;
        ms += Mv;
      }
      S[0] += ms;
      for (int o = 8; o > 0; o >>= 1) ms += __shfl_xor_sync(0xffffffffu, ms, o);
      m[r] = ms;                                                  // full row sum, on all 16 threads of the row group
    }
  }
  // quadratic forms per dimension: this thread's share of sum_a x_ak^2 m_a - sum_ab x_ak M_ab x_bk
  auto dim_sum = [&](int k) -> double {
    double xa[RBR], xb[RBC];
#pragma unroll
    for (int r = 0; r < RBR; ++r) xa[r] = Xs[(ty + 32 * r) * xld + k];
#pragma unroll
    for (int c = 0; c < RBC; ++c) xb[c] = Xs[(tx + 16 * c) * xld + k];
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < RBR; ++r) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < RBC; ++c) t = fma(A[r][c], xb[c], t);
      acc = fma(-xa[r], t, acc);
      if (tx == 0) acc = fma(xa[r] * xa[r], m[r], acc);
    }
    return 2.0 * acc;
  };
  const int chunks = (d + DCH - 1) / DCH;
  const int n_sweeps = p.hv.any_ard ? chunks : 1;
  if (!(p.hv.any_ard && chunks == 1)) {       // Q = sum_ab |x_a - x_b|^2 M_ab = sum_k D_k (RBF sigma); one ARD chunk: below
    double qsum = 0.0;
    for (int k = 0; k < d; ++k) qsum += dim_sum(k);
    Q[0] = qsum;
  }
  for (int sw = 0; sw < n_sweeps; ++sw) {
    const int ts = p.hv.any_ard ? 0 : -1;
    const int k0 = p.hv.any_ard ? sw * DCH : 0;
    const int kn = p.hv.any_ard ? ((d - k0 < DCH) ? (d - k0) : DCH) : 0;
#pragma unroll
    for (int k = 0; k < DCH; ++k) D[k] = (k < kn) ? dim_sum(k0 + k) : 0.0;
    if (p.hv.any_ard && chunks == 1) {
      double qsum = 0.0;
#pragma unroll
      for (int k = 0; k < DCH; ++k) qsum += D[k];
      Q[0] = qsum;
    }
    ex_gradient_emit<REG_WARPS>(p.hv, sw, ts, k0, kn, S, Q, D, trW, -0.5, out + 1, sums, red);
  }
}

size_t bcm_nll_reg_smem_bytes(int rbr, int rbc, int d) {
  const size_t nr = 32 * static_cast<size_t>(rbr), np = 16 * static_cast<size_t>(rbc);
  return sizeof(double) * (5 * nr + 16 + EX_SUMS + nr * (d | 1) + nr * np);
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
  // one non-Eye term, experts of <= 128 points: the register-resident kernel (SGP_BCM_IMPL=smem selects the other one)
  static const bool force_smem = [] { const char* e = getenv("SGP_BCM_IMPL"); return e && std::string(e) == "smem"; }();
  if (kf.n_terms == 1 && n_max <= 128 && !force_smem) {
    const int rb = (n_max + 15) / 16;            // 16-column blocks
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kern, int rbr, int rbc) {
      const size_t smem = bcm_nll_reg_smem_bytes(rbr, rbc, d);
      if (smem > 227 * 1024) return false;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return true;
      kern<<<static_cast<unsigned>(E), REG_THREADS, smem, s>>>(p);
      e = cudaGetLastError();
      return true;
    };
    bool done = false;
    if (rb <= 4) done = go(bcm_nll_reg_kernel<2, 4>, 2, 4);
    else if (rb <= 6) done = go(bcm_nll_reg_kernel<3, 6>, 3, 6);
    else if (rb <= 7) done = go(bcm_nll_reg_kernel<4, 7>, 4, 7);
    else done = go(bcm_nll_reg_kernel<4, 8>, 4, 8);
    if (done) {
      if (e != cudaSuccess) return e;
      return launch_rows_reduce(dTotal, dPerExpert, E, 1 + n_hypers, s);
    }
  }
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
