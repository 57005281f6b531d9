// common.cuh — shared declarations for libbopl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bopl_b200.h"

namespace bopl {

constexpr int TRSM_BM = 128;   // rows of L per block row (= diagonal block size)
constexpr int TRSM_BC = 128;   // candidate granule of the host-path chunking (both tile widths divide it)
constexpr int TRSM_BK = 16;    // k-depth of one pipeline stage
constexpr int MAX_D = 24;      // input dimensions the shared-memory staging is sized for
constexpr int MAX_M = 16;      // attributes
constexpr int MAX_TOPK = 256;
constexpr int SMALL_N = 32;     // batches up to this size take the matrix-vector path (explicit inverse factor) when Linv is loaded

// Device-side view of one GP (attribute j under hyper-sample h).  All pointers device memory owned by the handle.
struct GpDev {
  const double* Lneg;    // [n_pad*n_pad] front-padded factor, off-diagonal 128-blocks NEGATED (acc += (-L)·V); diagonal blocks unused
  const double* Dinv;    // [n_pad/128][128][128] inverses of the diagonal blocks (lower triangular)
  const double* StageBlk;  // [n_pad/128][(d+1)*128] per block: scaled training rows transposed [q][128], then alpha[128]
  const double* DinvFrag;  // [n_pad/128][136*64] the same inverses packed as DMMA B fragments (posterior_trsm_t.cu)
  const double* Xs;      // [n_pad*d] training inputs divided by the lengthscales, front-padded with zeros
  const double* alpha;   // [n_pad] woodbury vector, front-padded with zeros
  const double* Winv;    // [n*n] or nullptr
  const double* Linv;    // [n*n] explicit inverse of the Cholesky factor (lower triangular, row-major) or nullptr
  const double* ell;     // [d]
  const signed char* OzL;   // int8 digit images of the off-diagonal 64x64 blocks (posterior_ozaki.cu) or nullptr
  const double* OzStage;    // [n_pad/64][...] per 64-row block: Dinv B fragments, scaled training rows^T, alpha, row scales of the digits
  double variance, noise, ymean;
  int kind;
  int _pad;
};

}  // namespace bopl

struct bopl_handle {
  int device = 0;
  int sms = 0;
  int n = 0, d = 0, m = 0, H = 0;
  int n_pad = 0, pad = 0;
  bool has_model = false, has_winv = false, has_linv = false;
  int small_path = 1;                  // BOPL_SMALL_PATH=0 forces the blocked TRSM kernel for every batch size
  bool small_grad_follows = false;     // set by the value+gradient entry points around their posterior + posterior-gradient pair (small-batch path)
  double* small_scratch = nullptr; size_t small_scratch_bytes = 0;
  std::string err;
  long long launches = 0;
  int trsm_stages = 3;                 // BOPL_TRSM_STAGES=4: deeper pipeline when shared memory allows (measured no faster)
  double trsm_l2_keep_frac = 0.75;     // share of the L2 the posterior kernel may pin (factor + first panel block rows)
  int trsm_keep_blocks = -1;           // BOPL_TRSM_KEEP_BLOCKS: panel block rows per CTA kept in L2 (-1: from the budget)
  size_t attr_smem_t[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // dynamic-smem attribute already set on THIS device, per posterior_trsm_t instantiation
  int trsm_l2_hints = 1;               // BOPL_TRSM_L2_HINTS: 0 none, 1 eviction-priority hints on the L / V tile loads, 2 also on the V stores
  // model storage
  std::vector<void*> owned;            // every cudaMalloc'ed model buffer
  bopl::GpDev* gp_dev = nullptr;       // [m*H] device array, index j*H + h
  std::vector<bopl::GpDev> gp_host;    // host mirror
  std::vector<int> fit_info;           // per (attribute, hyper-sample) instance of the last bopl_fit_model: 0, or 1 + the failed pivot
  double* X_dev = nullptr;             // [n*d] unscaled training inputs
  // scratch
  double* trsm_scratch = nullptr; size_t trsm_scratch_bytes = 0;   // per-CTA V panels
  double* mean_s = nullptr; double* var_s = nullptr; size_t mean_bytes = 0, var_bytes = 0;        // [m*N]
  double* dmean_s = nullptr; double* dvar_s = nullptr; size_t dmean_bytes = 0, dvar_bytes = 0;    // [m*N*d]
  double* grad_scratch = nullptr; size_t grad_scratch_bytes = 0;               // K*, G, beta panels + padded Winv of the gradient path
  bool gemm_attr_set = false;
  double* acq_s = nullptr; size_t acq_cap = 0;                                 // [N] (host path / top-k only callers)
  double* F_s = nullptr; size_t F_bytes = 0; bool F_valid = false;             // [H*m*n] posterior mean at the training points
  double* topk_val_s = nullptr; long long* topk_idx_s = nullptr; size_t topk_cap = 0;
  // host-path staging
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_order = nullptr; cudaStream_t order_stream = nullptr; bool order_valid = false;   // cross-stream ordering of the scratch (api.cu)
  double* hx_dev = nullptr; size_t hx_cap = 0;       // [2][chunk*d] candidate double buffer
  double* hacq_dev = nullptr; size_t hacq_cap = 0;   // [N]
  double* gh_dev = nullptr; size_t gh_cap = 0;       // staging of the host-buffer value+gradient call
  void* fit_ws = nullptr; size_t fit_ws_bytes = 0;   // workspace of bopl_log_marginal (factor, inverse, K^-1 of every instance)
  double* hsmall_dev = nullptr;        // Z, theta, wts, best staging
  double* coll_s = nullptr; size_t coll_bytes = 0;   // send / receive records of bopl_topk_allgather (collective.cu)
  void* path = nullptr;                // Thompson sample path in progress (path_kernels.cu)
  // posterior kernel for batches above the small-batch limit (bopl_set_posterior_kernel / BOPL_POSTERIOR):
  // use_ozaki = 1 (default): the update on the int8 tensor pipe (posterior_ozaki.cu), oz_digits per operand; 0: fp64 DMMA kernel
  int use_ozaki = 1;
  bool has_oz = false;
  int oz_cluster = 1;                  // BOPL_OZ_CLUSTER: 1 (default) row-paired CTA clusters where they are the faster launch form, 2 always, 0 never (posterior_ozaki.cu)
  int oz_digits = 7;                   // digits of the model's images (fixed at upload): 7 = 55 fractional bits (default), 6 = 47
  bool oz_auto = false;                // BOPL_POSTERIOR_INT8_AUTO: 6 digits when the upload-time probe of the model passes, else 7
  void* oz_scratch = nullptr; size_t oz_scratch_bytes = 0; size_t oz_attr_smem[3] = {0, 0, 0};
  size_t hsmall_cap = 0;
};

namespace bopl {

extern thread_local std::string g_err;   // error text when no handle exists

int fail(bopl_handle* h, int code, const std::string& msg);

#define BOPL_CUDA(h, call)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return bopl::fail((h), BOPL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));  \
  } while (0)

#define BOPL_LAUNCH_CHECK(h, name)                                                                \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return bopl::fail((h), BOPL_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(_e)); \
    (h)->launches++;                                                                              \
  } while (0)

#define BOPL_CHECK_HANDLE(h)                                             \
  do {                                                                   \
    if (!(h)) return bopl::fail(nullptr, BOPL_ERR_INVALID, "null handle"); \
    BOPL_CUDA((h), cudaSetDevice((h)->device));                           \
  } while (0)

#define BOPL_NEED_MODEL(h)                                                                        \
  do {                                                                                            \
    if (!(h)->has_model) return bopl::fail((h), BOPL_ERR_STATE, "bopl_set_model has not been called"); \
  } while (0)

// The handle's scratch buffers (mean / variance, gradients, solved panels, F, top-k levels) are shared by every entry point.  Calls on
// ONE stream are ordered by the stream; when consecutive calls on a handle use DIFFERENT streams (a caller's torch stream, then a
// host-buffer entry point on the library's own s_compute) the later stream first waits for an event recorded behind the earlier call.
struct StreamOrder {
  bopl_handle* h;
  cudaStream_t st;
  StreamOrder(bopl_handle* h_, cudaStream_t st_) : h(h_), st(st_) {
    if (h->order_valid && h->order_stream != st) cudaStreamWaitEvent(st, h->ev_order, 0);
  }
  ~StreamOrder() {
    if (cudaEventRecord(h->ev_order, st) == cudaSuccess) {
      h->order_stream = st;
      h->order_valid = true;
    }
  }
};

int ensure_bytes(bopl_handle* h, void** p, size_t* cap, size_t bytes);
void free_model(bopl_handle* h);
void path_free(bopl_handle* h);         // drops the sample path (it refers to the model's factor)

// device-side model fit (fit_kernels.cu); FitDev is private to that translation unit
size_t fit_dev_struct_bytes();
void fit_dev_fill(void* slot, double* A, double* Dinv, double* DinvFrag, double* U, double* Winv, double* Linv, double* Xs,
                  double* alpha, double* StageBlk, double* ell, const double* y, double* logdiag, double* quad, int* info,
                  double variance, double diag_add, int kind);
int fit_factorize(bopl_handle* h, const void* fits_dev, int G, const double* X_dev, int n, int d, bool want_inverse,
                  bool want_winv, bool want_linv, cudaStream_t st);
int fit_lml_grad(bopl_handle* h, const void* fits_dev, int G, int n, int d, double* partial, double* out, cudaStream_t st);

// launchers implemented in the kernel translation units
int launch_posterior_trsm(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                          cudaStream_t st);
int launch_posterior_trsm_t(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                            cudaStream_t st);
bool ozaki_available(const bopl_handle* h);
bool ozaki_covers(int d, int n_pad);
void ozaki_prepare(const double* Lp, int n_pad, int S, std::vector<signed char>& img, std::vector<double>& scale);
int ozaki_prepare_device(bopl_handle* h, cudaStream_t st);
void ozaki_pack_stage(const double* Dinv64, const double* Xs, const double* alpha, const double* scale, int n_pad, int d,
                      std::vector<double>& out);
int launch_posterior_ozaki(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags, cudaStream_t st);
void pack_dinv_fragments(const double* Dinv_block, double* out);
void pack_stage_blocks(const double* Xs, const double* alpha, int n_pad, int d, double* out);
bool use_small_path(const bopl_handle* h, int N);
int launch_posterior_small(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                           cudaStream_t st);
int launch_posterior_mean(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, cudaStream_t st);
int launch_posterior_mean_strided(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, long long out_stride,
                                  cudaStream_t st);
int launch_cross_cov(bopl_handle* h, const double* Xc, int N, int hsel, int j, double* K, cudaStream_t st);
int launch_posterior_grad(bopl_handle* h, const double* Xc, int N, int hsel, double* dmean, double* dvar,
                          cudaStream_t st);
int launch_incumbent(bopl_handle* h, const double* F, int util_kind, const double* theta, int Lt, int p,
                     double* best, cudaStream_t st);
int launch_uei_mc(bopl_handle* h, const double* mean, const double* var, int N, const double* Z, int nw, int util_kind,
                  const double* theta, int Lt, int p, const double* wts, const double* best, double scale,
                  int accumulate, double* acq, cudaStream_t st, int plain = 0);
int launch_uei_mc_grad(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                       int N, const double* Z, int nw, int util_kind, const double* theta, int Lt, int p,
                       const double* wts, const double* best, double scale, int accumulate, double* acq, double* dacq,
                       cudaStream_t st, int plain = 0);
int launch_uei_affine(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                      int N, const double* theta, int Lt, const double* wts, const double* best, double scale,
                      int accumulate, double* acq, double* dacq, cudaStream_t st);
int launch_uei_constrained(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                           int N, const double* theta, int Lt, const double* wts, const double* best, double scale,
                           int accumulate, double* acq, double* dacq, cudaStream_t st);
int launch_topk(bopl_handle* h, const double* acq, int N, int k, long long index_offset, double* val, long long* idx,
                cudaStream_t st);

// ---------------------------------------------------------------------------------------------- device math
// Branch-free fp64 sqrt and exp for the cross-covariance.  CUDA's libm versions carry slow-path branches (denormals,
// overflow), which keep the compiler from interleaving the independent evaluations of neighbouring kernel entries;
// K* generation was latency-bound because of it (profiles/r01_trsm_v1_ncu.md).  Both are accurate to ~1 ulp over the
// ranges the kernels produce (checked against the oracle at 1e-13 in tests/test_gpu_parity.py::test_cross_cov_kernels).

// sqrt(a) for a >= 0 (0 and denormals -> 0): MUFU.RSQ64H seed (2^-22.9), ONE Newton step on 1/sqrt (2^-45), then the residual
// step on sqrt itself, which squares the error again (tools/kern_math_check.cu: <= 0.5 ulp + rounding over 1e-300 .. 1e300).
__device__ __forceinline__ double sqrt_nonneg(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  const double e = fma(-h, y * y, 0.5);
  y = fma(y, e, y);
  double r = a * y;
  r = fma(fma(-r, r, a), 0.5 * y, r);
  return (a >= 1e-300) ? r : 0.0;
}

// 2^(j/64), j = 0..63, correctly rounded.
__device__ const double EXP2_64TH[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// exp(x) for x <= 0 (clamped at -700: the result is then ~1e-304 instead of a denormal or 0).
// x = (64 e + j) ln2/64 + r, |r| <= ln2/128 (Cody-Waite, two-part ln2/64): exp(x) = 2^e T_j (1 + expm1(r)) with T_j = 2^(j/64) from a
// 512-byte table and expm1(r) = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120) (remainder 3.5e-17) — 6 fp64 operations after the reduction
// instead of the 13-term polynomial the |r| <= ln2/2 reduction needs: K* generation competes with the tensor work for the FP64
// pipe (profiles/r01_fp64_vs_int8.json), the table read and the index arithmetic do not.  T_j (1 + em) is formed as fma(T_j, em, T_j):
// only the table entry's and the final rounding remain (tools/kern_math_check.cu: <= 1.0 ulp).
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, -700.0);
  const double magic = 6755399441055744.0;   // 1.5 * 2^52: round-to-nearest integer in the low word
  double kd = fma(x, 0x1.71547652b82fep+6, magic);
  const int n = __double2loint(kd);
  kd -= magic;
  double r = fma(kd, -0x1.62e42fee00000p-7, x);
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  const double t = __ldg(&EXP2_64TH[n & 63]);
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  const double em = fma(q, r * r, r);
  const double p = fma(t, em, t);
  return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
}

// exp(x) for any finite x, same table-driven evaluation (clamped at +-700 instead of producing 0 / inf): the exponential utility of
// the MC kernel (experiments/test_vlmop3_exp.py:41-52) evaluates it m n_w L times per candidate — 6.4e9 per launch at the benchmark
// shape, 14 ms with libm's exp against ~4 ms with this one (10 fp64 operations, <= 1 ulp).
__device__ __forceinline__ double exp_any(double x) {
  x = fmin(fmax(x, -700.0), 700.0);
  const double magic = 6755399441055744.0;
  double kd = fma(x, 0x1.71547652b82fep+6, magic);
  const int n = __double2loint(kd);
  kd -= magic;
  double r = fma(kd, -0x1.62e42fee00000p-7, x);
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  const double t = __ldg(&EXP2_64TH[n & 63]);
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  const double em = fma(q, r * r, r);
  const double p = fma(t, em, t);
  return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
}

// Stationary kernels as functions of the squared scaled distance r2 (already clipped at 0).
// K_of_r: GPy/kern/src/stationary.py:529-530 (Mat52), :440-441 (Mat32), :594-595 (ExpQuad); rbf.py:42-43.
__device__ __forceinline__ double kern_value(int kind, double variance, double r2) {
  if (kind == BOPL_KERN_MAT52) {
    const double s5 = 2.23606797749978969641;
    double r = sqrt_nonneg(r2);
    return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp_nonpos(-s5 * r);
  } else if (kind == BOPL_KERN_MAT32) {
    const double s3 = 1.73205080756887729353;
    double r = sqrt_nonneg(r2);
    return variance * (1.0 + s3 * r) * exp_nonpos(-s3 * r);
  } else {
    double r = sqrt_nonneg(r2);
    return variance * exp_nonpos(-0.5 * (r * r));
  }
}

// Eight kernel values at once, written operation-by-operation across the eight lanes so that the eight independent
// sqrt / exp dependency chains interleave in the instruction stream (the same arithmetic as kern_value, bit for bit).
__device__ __forceinline__ void kern_value8(int kind, double variance, const double* r2, double* out) {
  double r[8], x[8], pre[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2[u]));
    r[u] = y;
  }
  {
    double e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = fma(-(0.5 * r2[u]), r[u] * r[u], 0.5);
#pragma unroll
    for (int u = 0; u < 8; ++u) r[u] = fma(r[u], e[u], r[u]);      // 1/sqrt
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = r2[u] * r[u];                // sqrt estimate
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = fma(fma(-e[u], e[u], r2[u]), 0.5 * r[u], e[u]);
#pragma unroll
    for (int u = 0; u < 8; ++u) r[u] = (r2[u] >= 1e-300) ? e[u] : 0.0;
  }
  if (kind == BOPL_KERN_MAT52) {
    const double s5 = 2.23606797749978969641;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      pre[u] = variance * (1.0 + s5 * r[u] + (5.0 / 3.0) * (r[u] * r[u]));
      x[u] = -s5 * r[u];
    }
  } else if (kind == BOPL_KERN_MAT32) {
    const double s3 = 1.73205080756887729353;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      pre[u] = variance * (1.0 + s3 * r[u]);
      x[u] = -s3 * r[u];
    }
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      pre[u] = variance;
      x[u] = -0.5 * (r[u] * r[u]);
    }
  }
  int n[8];
  double t[8], q[8];
  const double magic = 6755399441055744.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    x[u] = fmax(x[u], -700.0);
    double kd = fma(x[u], 0x1.71547652b82fep+6, magic);
    n[u] = __double2loint(kd);
    kd -= magic;
    const double rr = fma(kd, -0x1.62e42fee00000p-7, x[u]);
    x[u] = fma(kd, -0x1.a39ef35793c76p-39, rr);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) t[u] = __ldg(&EXP2_64TH[n[u] & 63]);
#pragma unroll
  for (int u = 0; u < 8; ++u) q[u] = fma(x[u], 1.0 / 120.0, 1.0 / 24.0);
#pragma unroll
  for (int u = 0; u < 8; ++u) q[u] = fma(q[u], x[u], 1.0 / 6.0);
#pragma unroll
  for (int u = 0; u < 8; ++u) q[u] = fma(q[u], x[u], 0.5);
#pragma unroll
  for (int u = 0; u < 8; ++u) q[u] = fma(q[u], x[u] * x[u], x[u]);
#pragma unroll
  for (int u = 0; u < 8; ++u) q[u] = fma(t[u], q[u], t[u]);
#pragma unroll
  for (int u = 0; u < 8; ++u)
    out[u] = pre[u] * __hiloint2double(__double2hiint(q[u]) + ((n[u] >> 6) << 20), __double2loint(q[u]));
}

// value and g = (dK/dr)/r with 1/r := 0 at r == 0  (stationary.py:227-234 _inv_dist; dK_dr :532-533, :443-444, :597-598; rbf.py:45-46)
__device__ __forceinline__ void kern_value_grad(int kind, double variance, double r2, double& k, double& g) {
  double r = sqrt_nonneg(r2);
  double invr = (r != 0.0) ? 1.0 / r : 0.0;
  if (kind == BOPL_KERN_MAT52) {
    const double s5 = 2.23606797749978969641;
    double e = exp_nonpos(-s5 * r);
    k = variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * e;
    g = variance * ((10.0 / 3.0) * r - 5.0 * r - (5.0 * s5 / 3.0) * (r * r)) * e * invr;
  } else if (kind == BOPL_KERN_MAT32) {
    const double s3 = 1.73205080756887729353;
    double e = exp_nonpos(-s3 * r);
    k = variance * (1.0 + s3 * r) * e;
    g = -3.0 * variance * r * e * invr;
  } else {
    k = variance * exp_nonpos(-0.5 * (r * r));
    g = -r * k * invr;
  }
}

}  // namespace bopl
