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
constexpr int TRSM_BC = 128;   // candidates per tile
constexpr int TRSM_BK = 16;    // k-depth of one pipeline stage
constexpr int MAX_D = 24;      // input dimensions the shared-memory staging is sized for
constexpr int MAX_M = 16;      // attributes
constexpr int MAX_TOPK = 256;

// Device-side view of one GP (attribute j under hyper-sample h).  All pointers device memory owned by the handle.
struct GpDev {
  const double* Lneg;    // [n_pad*n_pad] front-padded factor, off-diagonal 128-blocks NEGATED (acc += (-L)·V); diagonal blocks unused
  const double* Dinv;    // [n_pad/128][128][128] inverses of the diagonal blocks (lower triangular)
  const double* Xs;      // [n_pad*d] training inputs divided by the lengthscales, front-padded with zeros
  const double* alpha;   // [n_pad] woodbury vector, front-padded with zeros
  const double* Winv;    // [n*n] or nullptr
  const double* ell;     // [d]
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
  bool has_model = false, has_winv = false;
  std::string err;
  long long launches = 0;
  // model storage
  std::vector<void*> owned;            // every cudaMalloc'ed model buffer
  bopl::GpDev* gp_dev = nullptr;       // [m*H] device array, index j*H + h
  std::vector<bopl::GpDev> gp_host;    // host mirror
  double* X_dev = nullptr;             // [n*d] unscaled training inputs
  // scratch
  double* trsm_scratch = nullptr; size_t trsm_scratch_bytes = 0;   // per-CTA V panels
  double* mean_s = nullptr; double* var_s = nullptr; size_t mean_bytes = 0, var_bytes = 0;        // [m*N]
  double* dmean_s = nullptr; double* dvar_s = nullptr; size_t dmean_bytes = 0, dvar_bytes = 0;    // [m*N*d]
  double* grad_scratch = nullptr; size_t grad_scratch_bytes = 0;               // K*, G, beta panels of the gradient path
  double* acq_s = nullptr; size_t acq_cap = 0;                                 // [N] (host path / top-k only callers)
  double* F_s = nullptr; size_t F_bytes = 0; bool F_valid = false;             // [H*m*n] posterior mean at the training points
  double* topk_val_s = nullptr; long long* topk_idx_s = nullptr; size_t topk_cap = 0;
  // host-path staging
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  double* hx_dev = nullptr; size_t hx_cap = 0;       // [2][chunk*d] candidate double buffer
  double* hacq_dev = nullptr; size_t hacq_cap = 0;   // [N]
  double* hsmall_dev = nullptr;        // Z, theta, wts, best staging
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

int ensure_bytes(bopl_handle* h, void** p, size_t* cap, size_t bytes);

// launchers implemented in the kernel translation units
int launch_posterior_trsm(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
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
                  int accumulate, double* acq, cudaStream_t st);
int launch_uei_mc_grad(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                       int N, const double* Z, int nw, int util_kind, const double* theta, int Lt, int p,
                       const double* wts, const double* best, double scale, int accumulate, double* acq, double* dacq,
                       cudaStream_t st);
int launch_uei_affine(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                      int N, const double* theta, int Lt, const double* wts, const double* best, double scale,
                      int accumulate, double* acq, double* dacq, cudaStream_t st);
int launch_topk(bopl_handle* h, const double* acq, int N, int k, long long index_offset, double* val, long long* idx,
                cudaStream_t st);

// ---------------------------------------------------------------------------------------------- device math
// Stationary kernels as functions of the squared scaled distance r2 (already clipped at 0).
// K_of_r: GPy/kern/src/stationary.py:529-530 (Mat52), :440-441 (Mat32), :594-595 (ExpQuad); rbf.py:42-43.
__device__ __forceinline__ double kern_value(int kind, double variance, double r2) {
  if (kind == BOPL_KERN_MAT52) {
    const double s5 = 2.23606797749978969641;
    double r = sqrt(r2);
    return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r);
  } else if (kind == BOPL_KERN_MAT32) {
    const double s3 = 1.73205080756887729353;
    double r = sqrt(r2);
    return variance * (1.0 + s3 * r) * exp(-s3 * r);
  } else {
    double r = sqrt(r2);
    return variance * exp(-0.5 * (r * r));
  }
}

// value and g = (dK/dr)/r with 1/r := 0 at r == 0  (stationary.py:227-234 _inv_dist; dK_dr :532-533, :443-444, :597-598; rbf.py:45-46)
__device__ __forceinline__ void kern_value_grad(int kind, double variance, double r2, double& k, double& g) {
  double r = sqrt(r2);
  double invr = (r != 0.0) ? 1.0 / r : 0.0;
  if (kind == BOPL_KERN_MAT52) {
    const double s5 = 2.23606797749978969641;
    double e = exp(-s5 * r);
    k = variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * e;
    g = variance * ((10.0 / 3.0) * r - 5.0 * r - (5.0 * s5 / 3.0) * (r * r)) * e * invr;
  } else if (kind == BOPL_KERN_MAT32) {
    const double s3 = 1.73205080756887729353;
    double e = exp(-s3 * r);
    k = variance * (1.0 + s3 * r) * e;
    g = -3.0 * variance * r * e * invr;
  } else {
    k = variance * exp(-0.5 * (r * r));
    g = -r * k * invr;
  }
}

}  // namespace bopl
