// Internal declarations shared by the translation units of libsgp.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cublas_v2.h>
#include <nccl.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sgp.h"

namespace sgp {

// SGP_PREC_AUTO runs the int8 Gram on accumulate calls of at least this many points (smaller ones: fp64 DMMA kernel)
constexpr long long kAutoI8MinPoints = 32768;
constexpr int kMaxTerms = 4;      // non-Eye terms of a flattened kernel
constexpr int kTile = 128;        // edge of one G tile (active-set indices per CTA tile)
constexpr int kSMsB200 = 148;

// Flattened, device-ready kernel description (built by sgp_stats_begin).
struct KernelFlat {
  int n_terms = 0;                 // non-Eye terms
  double scale[kMaxTerms] = {0};   // C_t
  double eye_sum = 0.0;            // sum of Eye coefficients = whiteNoiseVar
  double self_kernel = 0.0;        // sum of all scales (every leaf has k(x,x) = 1)
};

// Launch parameters of the fused K_mn + Gram kernel (gram_f64.cu).
struct GramParams {
  const void* X;        // n x d row-major, fp32 or fp64
  const double* y;      // n
  long long n;
  int x_is_f32;
  int d, dpad;          // dpad = d rounded up to a multiple of 4
  int m, m_pad;         // m_pad = m rounded up to a multiple of kTile
  int n_terms;
  double scale[kMaxTerms];
  const double* Zs;     // [n_terms][m_pad][dpad]  beta-scaled active set (zero padded)
  const double* beta;   // [n_terms][dpad]         per-term feature scales (zero padded)
  double* Gpart;        // [n_slices][m_pad*m_pad] per-slice partial tiles (lower block triangle)
  double* bpart;        // [n_slices][m_pad]
  int n_slices;
  int n_tiles_1d;       // m_pad / kTile
};

// Grow-only device scratch block owned by a context (no cudaMalloc / cudaFree in steady state).
struct DevScratch {
  void* p = nullptr;
  size_t cap = 0;
};

struct Ctx {
  int device = 0;
  int precision = SGP_PREC_AUTO;
  int last_path = -1;    // arithmetic path of the last statistics launch (SGP_PREC_F64 / _STRICT / _I8)
  std::string err;
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_stream = nullptr;  // H2D staging
  cudaStream_t tail_stream = nullptr;  // second chain of the m x m tail (inv(K_mm) next to the A chain)
  cusolverDnHandle_t solver = nullptr;
  cusolverDnHandle_t solver2 = nullptr;   // bound to tail_stream
  cublasHandle_t blas = nullptr;
  cudaEvent_t tail_fork = nullptr, tail_join = nullptr;
  DevScratch tail_ws, tail_ws2, predict_ws, cross_ws, bcm_ws, sweep_ws, greedy_ws;
  bool has_magic_run = false;
  bool tail_fast = false;              // last sgp_magic took the Cholesky path for both matrices
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int num_sms = kSMsB200;

  // active set / kernel (valid after begin)
  bool begun = false, finished = false, has_magic = false;
  int m = 0, d = 0, dpad = 0, m_pad = 0;
  int alloc_terms = 0;   // term count the active-set buffers were sized for
  KernelFlat kf;
  double* dZ = nullptr;      // m x d raw active set (fp64)
  double* dZs = nullptr;     // [n_terms][m_pad][dpad]
  double* dBeta = nullptr;   // [n_terms][dpad]
  double* dGb = nullptr;     // packed [G (m*m) ; b (m)] -- one all-reduce
  double* dGpart = nullptr;  size_t gpart_bytes = 0;
  double* dBpart = nullptr;  size_t bpart_bytes = 0;
  int n_slices = 1;
  // tcgen05 int8 path (gram_i8.cu)
  bool i8_ok = false;            // kernel/shape qualifies for tensor-core distances (one non-Eye term, d <= 32)
  bool i8_direct_ok = false;     // ... for the direct-distance mode (<= 4 non-Eye terms, n_terms * (dpad4 + 4) <= 72)
  int i8_dpad4 = 0;
  float* dI8Zd = nullptr;        // direct mode: active-set tiles [tile][term][128][dpad4 + 4] fp32
  double* dI8DScale = nullptr;   // [n_terms][dpad4] sqrt(log2 e) * beta_tk ; then [dpad4] centre
  bool i8_direct_used = false;
  bool i8_used = false;          // an int8 launch contributed to the current statistics
  std::vector<double> i8d_sc;      // direct mode: host copy of [term][k] scales + centre of the current begin() window
  bool i8d_prepared = false;       // ... uploaded and the active-set tiles built (on first use of the mode in the window)
  size_t i8d_zd_bytes = 0, i8d_sc_bytes = 0;
  double i8d_z_norm_mean = 0.0;    // direct mode: mean scaled squared norm of the active set (widest term) for AUTO's budget
  float i8_direct_r2max = 2048.f;  // direct mode: largest scaled squared norm accepted (env SGP_I8_DIRECT_R2MAX); measured
                                   // dG 6.6e-7 at ~1300 and 2.7e-6 at ~7700 on clustered data (tests: error_growth)
  int call_path = 0;             // AUTO's decision for the current accumulate call (taken on its first chunk): 0 fp64,
                                 // 1 int8 Gram with tensor-core distances, 2 int8 Gram with direct fp32 distances
  double* dI8Scale = nullptr;    // [dp16] sqrt(log2 e) * beta_k
  double* dI8Centre = nullptr;   // [dp16] per-feature centre (active-set mean)
  int* dI8Flags = nullptr;       // bit 0: coordinates out of fp16 operand range
  double* dI8NormSum = nullptr;  // [0]: sum of |x^|^2 over every int8 chunk since begin; [1]: first chunk of this call
  long long i8_points = 0;       // points that went through the int8 kernel since begin
  double i8_z_norm_mean = 0.0;   // mean over the active set of |z^|^2
  double i8_norm_budget = 6.0;   // AUTO: int8 path only if mean|x^|^2 + mean|z^|^2 <= budget (posterior mean vs the all-fp64
                                 // mode at 4.3 / 5.8 / 7.2 / 7.9: 1.8e-6 / 2.4e-6 / 5.7e-6 / 1.0e-5, profiles/r02p_i8_budget_edge.txt)
  uint8_t* dI8Zt = nullptr;      // active-set operand images
  uint8_t* dI8Xt = nullptr;  size_t i8_xt_bytes = 0;   // point operand images (scratch)
  float* dI8Ys = nullptr;    size_t i8_ys_bytes = 0;
  // BCM objective state (experts resident on the device)
  double* dEx = nullptr; double* dEy = nullptr; long long* dEoff = nullptr;
  double* dEf = nullptr;         // per-point latent mode of the Laplace approximation (warm start across evaluations)
  long long n_experts = 0, ex_n = 0; int ex_d = 0; int ex_nmax = 0;
  double* dNllPer = nullptr; size_t nll_per_cap = 0;
  bool bcm_general = false;      // the last sgp_bcm_nll took the global-memory LU path
  void* dNllScratch = nullptr;   // hyper descriptors + totals + flags
  // tail / predict state
  double* dMagicVec = nullptr;   // m
  double* dMagicMat = nullptr;   // m x m
  size_t magic_cap = 0;
  // host->device staging (double buffered)
  void* stageX[2] = {nullptr, nullptr};
  double* stageY[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;       // per X buffer
  long long stage_points = 0;
  cudaEvent_t stage_free[2] = {nullptr, nullptr};   // recorded when the kernel consuming buffer i is done
  cudaEvent_t stage_ready[2] = {nullptr, nullptr};  // recorded when the H2D into buffer i is done
  // instrumentation
  int64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gram_events;   // event POOL (created once, reused)
  size_t gram_events_used = 0;                                    // pairs recorded since the last begin
  cudaEvent_t user_events[8] = {nullptr};
  float* dbgT = nullptr;       // sgp_debug_i8_tile
  uint32_t* dbgW = nullptr;
  long long* dbgClk = nullptr; // [2 CTAs][5 roles][32 units][8 events] clock64 timeline
  int* i8_pm_host = nullptr;   // host-mapped post-mortem record of the int8 kernel (8 ints), and its device alias
  void* i8_pm_dev = nullptr;
  int i8_impl = 0;             // 0: self-contained kernel (gram_i8.cu), 1: shared-panel ring kernel (gram_i8_ring.cu)
  DevScratch i8_share;         // L2 ring + flags through which diagonal CTAs publish their digit planes
};

// error helpers ---------------------------------------------------------------------------------
int fail(Ctx* c, int code, const std::string& msg);
#define SGP_CUDA(c, expr)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return ::sgp::fail((c), SGP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
  } while (0)

// kernels (each returns a cudaError_t from the launch) -----------------------------------------
cudaError_t launch_gram_f64(const GramParams& p, bool strict_elements, cudaStream_t s);
cudaError_t launch_gram_reduce(double* G /*m x m*/, double* b, const double* Gpart, const double* bpart,
                               int n_slices, int m, int m_pad, cudaStream_t s);
cudaError_t launch_gram_reduce_upper(double* G, double* b, const double* Gpart, const double* bpart, int n_slices, int m,
                                     int m_pad, int row_lo, int row_hi, cudaStream_t s);
// same, restricted to the G columns / b entries [col_lo, col_hi) (one int8 launch covers whole tile columns)
cudaError_t launch_gram_reduce_cols(double* G, double* b, const double* Gpart, const double* bpart, int n_slices, int m,
                                    int m_pad, int col_lo, int col_hi, cudaStream_t s);
cudaError_t launch_scale_rows(double* out /*[rows][dpad]*/, const double* in /*rows x d*/, const double* beta,
                              int rows_in, int rows_out, int d, int dpad, cudaStream_t s);
cudaError_t launch_kmm_build(double* Kmm /*m x m*/, const double* Zs, const double* scale_dev_or_null,
                             const KernelFlat& kf, int m, int m_pad, int dpad, cudaStream_t s);
cudaError_t launch_cross_kernel(double* K /*n x m*/, const double* X /*n x d*/, const double* Zs,
                                const double* beta, const KernelFlat& kf, long long n, int d, int dpad,
                                int m, int m_pad, cudaStream_t s);
cudaError_t launch_axpby_diag(double* A, const double* K, const double* G, double wn, int m, cudaStream_t s);
cudaError_t launch_set_identity(double* I, int m, cudaStream_t s);
cudaError_t launch_magic_matrix(double* out, const double* invA, const double* invK, double wn, int m,
                                cudaStream_t s);
cudaError_t launch_group_experts(double* Xe, double* ye, const void* dX, int x_is_f32, const double* dy, long long n, int d,
                                 long long E, long long p0, long long cn, cudaStream_t s);
cudaError_t launch_expert_offsets(long long* off, long long n, long long E, cudaStream_t s);
cudaError_t launch_status_to_double(double* dst, const int* flags, int mask, const double* norm_sum, double norm_limit,
                                    cudaStream_t s);
cudaError_t launch_predict_finish(double* mean, double* var, const double* K /*n x m*/,
                                  const double* W /*n x m = K*M*/, const double* mv, double self_k,
                                  long long n, int m, cudaStream_t s);

// tcgen05 path
size_t i8_points_scratch_bytes(long long n, int nchunks);
size_t i8_active_scratch_bytes(int m_pad, int nchunks);
int i8_nchunks(int d);
cudaError_t launch_i8_prep_active(uint8_t* Zt, const double* dZ, int m, int m_pad, int d, const double* dScale,
                                  const double* dCentre, int* dFlags, cudaStream_t s);
cudaError_t launch_i8_prep_points(uint8_t* Xt, float* ys, const void* dX, int x_is_f32, const double* dy, long long n,
                                  int d, const double* dScale, const double* dCentre, int* dFlags, double* dNormSum,
                                  double* dNormSumCall, cudaStream_t s);
// direct-distance mode of the ring kernel: fp32 coordinate tiles instead of fp16 operand images
cudaError_t launch_i8_prep_active_direct(float* Zd, const double* dZ, int m, int m_pad, int d, int dpad4, int n_terms,
                                         const double* dScale, const double* dCentre, int* flags, float r2max,
                                         cudaStream_t s);
cudaError_t launch_i8_prep_points_direct(float* Xd, float* ys, const void* dX, int x_is_f32, const double* dy, long long n,
                                         int d, int dpad4, int n_terms, const double* dScale, const double* dCentre, int* flags,
                                         float r2max, double* dNormSum, double* dNormSumCall, cudaStream_t s);
struct I8Direct {
  int on = 0;            // 1: exponents from fp32 direct-form distances on the CUDA cores (any norms, up to 4 terms)
  int n_terms = 0, dpad4 = 0;
  float w[4] = {0, 0, 0, 0};   // C_t / sum C
  double csum = 0.0;     // sum of the term scales = the fixed-point scale
};

// One cooperative launch of the int8 Gram kernel: whole tile columns [col_lo, col_hi) x n_slices point slices.
struct I8Launch {
  int col_lo, col_hi, tiles, n_slices;
};
int i8_plan(int m_pad, int num_sms, long long n_units, I8Launch* out, int max_out);
size_t i8_share_bytes(int m_pad, int n_slices);
size_t i8_share_flag_bytes(int m_pad, int n_slices);
cudaError_t launch_gram_i8_ring(const uint8_t* Xt, const float* ys, const uint8_t* Zt, long long n, int d, int m_pad,
                                const I8Launch& plan, const I8Direct& direct, double* Gpart, double* bpart, double C,
                                uint8_t* share, float* dbg_T, uint32_t* dbg_w, long long* dbg_clk, void* post_mortem,
                                cudaStream_t s);
// round-1 kernel (every CTA builds both panels of its tile; SGP_I8_IMPL=v1): kept as the A/B reference of the ring kernel
cudaError_t launch_gram_i8(const uint8_t* Xt, const float* ys, const uint8_t* Zt, long long n, int d, int m_pad,
                           int n_slices, double* Gpart, double* bpart, double C, float* dbg_T, uint32_t* dbg_w,
                           long long* dbg_clk, cudaStream_t s);

// K_nm sweep (kmn_sweep.cu): fp32 cross kernel of a block of points, row-major n x m
cudaError_t launch_kmn_sweep(const uint8_t* Xt, const uint8_t* Zt, long long n, int d, int m, int m_pad, int num_sms, double C,
                             float* K, cudaStream_t s);

// BCM objective (bcm_nll.cu)
size_t bcm_nll_smem_bytes(int n_max);
int bcm_nll_max_hypers();
cudaError_t launch_bcm_nll(const double* dX, const double* dy, const long long* dOff, long long E, int d, int n_max,
                           const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind, const int* dTerm,
                           const int* dDim, const double* dCoef, const double* dValue, int any_ard,
                           double* dPerExpert, double* dTotal, int* dFlags, cudaStream_t s);

// general path (any expert size, LU like the reference): global-memory kernel matrices + cuBLAS batched LU / inverse
size_t bcm_general_workspace_bytes(long long E, int n_max);
cudaError_t launch_bcm_nll_general(cublasHandle_t blas, int* blas_status, void* ws, const double* dX, const double* dy,
                                   const long long* dOff, long long E, int d, int n_max, const KernelFlat& kf,
                                   const double* dBeta, int n_hypers, const int* dKind, const int* dTerm, const int* dDim,
                                   const double* dCoef, const double* dValue, int any_ard, double* dPerExpert,
                                   double* dTotal, int* dFlags, cudaStream_t s);

size_t laplace_smem_bytes(int n_max);
cudaError_t launch_laplace(const double* dX, const double* dy, double* df, const long long* dOff, long long E, int d,
                           int n_max, const KernelFlat& kf, const double* dBeta, int n_hypers, const int* dKind,
                           const int* dTerm, const int* dDim, const double* dCoef, const double* dValue, int any_ard,
                           double tol,
                           double* dPerExpert, double* dTotal, int* dFlags, cudaStream_t s);

int ctx_scratch(Ctx* c, DevScratch& s, size_t bytes);
// csrc/greedy.cu: forward selection with rank-1 updates (ActiveSetProvider.scala:58-139); beta_flat = [n_terms][d]
int run_greedy(Ctx* c, const KernelFlat& kf, const std::vector<double>& beta_flat, const double* X, const double* y,
               long long n, int d, long long n_experts, long long first_index, int m_target, long long* indices_out);
int run_tail(Ctx* c, double* magic_vector, double* magic_matrix);
int run_predict(Ctx* c, const double* X, long long n, double* mean_out, double* var_out);

}  // namespace sgp
