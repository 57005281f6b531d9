// fit_kernels.cu — exact-GP model fit on the device (the model-update side of the path).
//
// What the reference does on the host with LAPACK once per model update and per (attribute, hyper-sample), and once per
// likelihood evaluation inside MAP / HMC hyper-parameter inference:
//   K = K(X, X) + (noise + 1e-8) I                 GPy/inference/latent_function_inference/exact_gaussian_inference.py:44-46
//   L = chol(K), alpha = K^-1 (y - ymean)          :46-49  (GPy/util/linalg.py:53-78 jitchol -> dpotrf, :115-125 dpotrs)
//   log|K| = 2 sum log L_ii                         :51
//   Winv = K^-1                                     posterior.py:171-191 (dpotri)
//   dL/dK = 0.5 (alpha alpha^T - K^-1)              exact_gaussian_inference.py:57
//   dL/d variance, dL/d lengthscale_q, dL/d noise   GPy/kern/src/stationary.py:191-215,236-245; likelihoods/gaussian.py:64-71
// Here every (attribute, hyper-sample) instance g is one slice of a batched launch, everything stays in HBM, and the
// results are produced directly in the layouts the posterior kernels consume (front-padded factor with negated
// off-diagonal 128-blocks, inverted diagonal blocks + their DMMA fragments, staged training rows).
//
// Algorithm: right-looking blocked Cholesky with 128 x 128 blocks.  Per block column b:
//   fit_potrf_diag_kernel   one CTA per instance: diagonal block in shared memory (LDL^T-style column sweep, one CTA
//                           barrier per column), its triangular inverse (the posterior kernel needs it anyway), log-diagonal
//   fit_panel_kernel        (-L)(i,b) = -A(i,b) Dinv_b^T             \  both are the same 128x128x(16k) "NT" tile product on
//   fit_syrk_kernel         A(i,j) -= (-L)(i,b) (-L)(j,b)^T, j <= i  /  DMMA.8x8x4 (gemm_nt), operands staged by cp.async
// then  fit_solve_kernel (alpha by blocked forward / backward substitution with the inverted diagonal blocks),
//       fit_linvT_kernel (U = L^-T block row by block row: U(b,i) = [sum_k U(b,k) (-L)(i,k)^T] Dinv_i^T),
//       fit_winv_kernel  (K^-1 = U U^T), fit_linv_kernel (L^-1 = U^T, the factor of the small-batch path),
//       lml_grad_kernel  (the n^2 contraction of dL/dK with dK/d theta, deterministic two-level reduction).
#include <cmath>

#include "pipeline.cuh"

namespace bopl {

namespace {

constexpr int FB = TRSM_BM;                 // 128
constexpr int FNT = 256;
constexpr int FTILE = FB * TRSM_BK;         // 2048 doubles
constexpr int FSTAGES = 3;
constexpr int FSTAGE_DBL = 2 * FTILE;
constexpr size_t GEMM_SMEM = (size_t)FSTAGES * FSTAGE_DBL * sizeof(double);   // 96 KB
constexpr int DS = FB + 1;                  // row stride of the diagonal block in shared memory
constexpr size_t DIAG_SMEM = ((size_t)FB * DS + 3 * FB) * sizeof(double);

// One GP instance (attribute j, hyper-sample h) of a batched fit.  All device pointers.
struct FitDev {
  double* A;         // [n_pad*n_pad] in: K + diag (lower blocks); out: diagonal blocks L_bb (lower), off-diagonal blocks -L
  double* Dinv;      // [nblk][128][128]
  double* DinvFrag;  // [nblk][136*64]
  double* U;         // [n_pad*n_pad] L^-T (upper triangular, row-major) or nullptr
  double* Winv;      // [n*n] or nullptr
  double* Linv;      // [n*n] or nullptr
  double* Xs;        // [n_pad*d] scaled training inputs, front-padded
  double* alpha;     // [n_pad]
  double* StageBlk;  // [nblk][(d+1)*128] or nullptr
  double* ell;       // [d]
  const double* y;   // [n] centred targets
  double* logdiag;   // [nblk]
  double* quad;      // [1] y^T alpha
  int* info;         // 0, or 1 + padded index of the first non-positive pivot
  double variance, diag_add;
  int kind, _pad;
};

__constant__ unsigned char c_frag_nt[136], c_frag_kt[136];

// acc += A[128 x 16 nk] * B[128 x 16 nk]^T.  A, B row-major in global memory with k contiguous.  Warp w owns rows 16w..16w+15
// of A against all 128 rows of B; thread (g, t): acc[mi][nt][e] = C[16w + 8mi + g][8nt + 2t + e].
__device__ __forceinline__ void gemm_nt(double (&acc)[2][16][2], const double* __restrict__ A, size_t lda,
                                        const double* __restrict__ B, size_t ldb, int nk, double* P, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int Cw = warp * 16, swz = (g & 1) << 2;
#pragma unroll
  for (int s = 0; s < FSTAGES - 1; ++s) {
    if (s < nk) {
      load_tile<FB, TRSM_BK, FNT>(P + s * FSTAGE_DBL, A + (size_t)s * TRSM_BK, lda, tid);
      load_tile<FB, TRSM_BK, FNT>(P + s * FSTAGE_DBL + FTILE, B + (size_t)s * TRSM_BK, ldb, tid);
    }
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<FSTAGES - 2>();
    __syncthreads();
    const int nxt = it + FSTAGES - 1;
    if (nxt < nk) {
      double* st = P + (nxt % FSTAGES) * FSTAGE_DBL;
      load_tile<FB, TRSM_BK, FNT>(st, A + (size_t)nxt * TRSM_BK, lda, tid);
      load_tile<FB, TRSM_BK, FNT>(st + FTILE, B + (size_t)nxt * TRSM_BK, ldb, tid);
    }
    cp_async_commit();
    const double* As = P + (it % FSTAGES) * FSTAGE_DBL;
    const double* Bs = As + FTILE;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ch = ((t + 4 * kk) ^ swz) << 1;
      double2 a[2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) a[mi] = *reinterpret_cast<const double2*>(As + (Cw + 8 * mi + g) * TRSM_BK + ch);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        double2 b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) b[u] = *reinterpret_cast<const double2*>(Bs + (8 * (8 * hf + u) + g) * TRSM_BK + ch);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][8 * hf + u][0], acc[mi][8 * hf + u][1], a[mi].x, b[u].x);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][8 * hf + u][0], acc[mi][8 * hf + u][1], a[mi].y, b[u].y);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ void zero_acc(double (&acc)[2][16][2]) {
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
}

// Narrow form for latency-bound chains (L^-T block rows): acc += A[32 x 16 nk] * B[128 x 16 nk]^T.  Warp w owns rows
// 16 (w & 1).. of A against rows 32 (w >> 1).. of B; thread (g, t): acc[mi][nt][e] = C[16 (w&1) + 8mi + g][32 (w>>1) + 8nt + 2t + e].
constexpr int F32_STAGES = 4;
constexpr int F32_STAGE_DBL = 32 * TRSM_BK + FTILE;
constexpr size_t GEMM32_SMEM = (size_t)F32_STAGES * F32_STAGE_DBL * sizeof(double);   // 80 KB
__device__ __forceinline__ void gemm_nt32(double (&acc)[2][4][2], const double* __restrict__ A, size_t lda,
                                          const double* __restrict__ B, size_t ldb, int nk, double* P, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int Rw = (warp & 1) * 16, Cw = (warp >> 1) * 32, swz = (g & 1) << 2;
#pragma unroll
  for (int s = 0; s < F32_STAGES - 1; ++s) {
    if (s < nk) {
      load_tile<32, TRSM_BK, FNT>(P + s * F32_STAGE_DBL, A + (size_t)s * TRSM_BK, lda, tid);
      load_tile<FB, TRSM_BK, FNT>(P + s * F32_STAGE_DBL + 32 * TRSM_BK, B + (size_t)s * TRSM_BK, ldb, tid);
    }
    cp_async_commit();
  }
  for (int it = 0; it < nk; ++it) {
    cp_async_wait<F32_STAGES - 2>();
    __syncthreads();
    const int nxt = it + F32_STAGES - 1;
    if (nxt < nk) {
      double* st = P + (nxt % F32_STAGES) * F32_STAGE_DBL;
      load_tile<32, TRSM_BK, FNT>(st, A + (size_t)nxt * TRSM_BK, lda, tid);
      load_tile<FB, TRSM_BK, FNT>(st + 32 * TRSM_BK, B + (size_t)nxt * TRSM_BK, ldb, tid);
    }
    cp_async_commit();
    const double* As = P + (it % F32_STAGES) * F32_STAGE_DBL;
    const double* Bs = As + 32 * TRSM_BK;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ch = ((t + 4 * kk) ^ swz) << 1;
      double2 a[2], b[4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) a[mi] = *reinterpret_cast<const double2*>(As + (Rw + 8 * mi + g) * TRSM_BK + ch);
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = *reinterpret_cast<const double2*>(Bs + (Cw + 8 * u + g) * TRSM_BK + ch);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][u][0], acc[mi][u][1], a[mi].x, b[u].x);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][u][0], acc[mi][u][1], a[mi].y, b[u].y);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ void store_tile32(const double (&acc)[2][4][2], double* C, size_t ldc, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int Rw = (warp & 1) * 16, Cw = (warp >> 1) * 32;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<double2*>(C + (size_t)(Rw + 8 * mi + g) * ldc + Cw + 8 * nt + 2 * t) = make_double2(acc[mi][nt][0], acc[mi][nt][1]);
}

// C[row][col] (128 x 128 tile, leading dimension ldc, 16-byte aligned rows) = sign * acc (+ C when ACCUM)
template <bool ACCUM>
__device__ __forceinline__ void store_tile(const double (&acc)[2][16][2], double* C, size_t ldc, double sign, int tid) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      double2* dst = reinterpret_cast<double2*>(C + (size_t)(warp * 16 + 8 * mi + g) * ldc + 8 * nt + 2 * t);
      double2 v = make_double2(sign * acc[mi][nt][0], sign * acc[mi][nt][1]);
      if (ACCUM) {
        const double2 o = *dst;
        v.x += o.x;
        v.y += o.y;
      }
      *dst = v;
    }
}

// ---------------------------------------------------------------------------------------------- preparation
// Xs = X / ell (front-padded with zeros), alpha = 0, staged blocks are packed after the solve.
__global__ void fit_prep_kernel(const FitDev* fits, const double* X, int n, int d, int n_pad, int pad) {
  const FitDev f = fits[blockIdx.y];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_pad * d; idx += gridDim.x * blockDim.x) {
    const int i = idx / d, q = idx - i * d;
    f.Xs[idx] = (i >= pad) ? X[(size_t)(i - pad) * d + q] / f.ell[q] : 0.0;   // stationary.py:161-162
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *f.info = 0;
}

// Lower 128-blocks of A = K(X, X) + diag_add I on the real rows, identity on the pad (stationary.py:104-113 with the
// zeroed diagonal of _unscaled_dist; exact_gaussian_inference.py:46).  grid (nblk*(nblk+1)/2, G), 256 threads.
__global__ void __launch_bounds__(256) fit_cov_kernel(const FitDev* fits, int d, int n_pad, int pad) {
  const FitDev f = fits[blockIdx.y];
  int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  __shared__ double xi[FB * MAX_D], xj[FB * MAX_D];
  for (int idx = threadIdx.x; idx < FB * d; idx += 256) {
    xi[idx] = f.Xs[(size_t)bi * FB * d + idx];
    xj[idx] = f.Xs[(size_t)bj * FB * d + idx];
  }
  __syncthreads();
  const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;      // 16 x 16 threads, 8 x 8 entries each (strided by 16)
  for (int a = 0; a < 8; ++a) {
    const int r = tr + 16 * a, R = bi * FB + r;
    for (int b = 0; b < 8; ++b) {
      const int c = tc + 16 * b, C = bj * FB + c;
      double v;
      if (R < pad || C < pad) {
        v = (R == C) ? 1.0 : 0.0;
      } else if (R == C) {
        v = f.variance + f.diag_add;
      } else {
        double r2 = 0.0;
        for (int q = 0; q < d; ++q) {
          const double dq = xi[r * d + q] - xj[c * d + q];
          r2 = fma(dq, dq, r2);
        }
        v = kern_value(f.kind, f.variance, r2);
      }
      f.A[(size_t)R * n_pad + C] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- diagonal block
// Cholesky factor of the 128 x 128 diagonal block b, its inverse (plain + DMMA-fragment order) and sum(log L_ii).
// One CTA of 16 x 16 threads per instance; thread (ty, tx) keeps the 8 x 8 elements (ty + 16 ia, tx + 16 kb) of the block
// in REGISTERS (cyclic distribution: every thread stays busy until the last columns).  Both sweeps publish one vector per
// step through a double-buffered shared-memory line, so a step costs one CTA barrier:
//   factorisation  column sweep on the unscaled columns, s_ik -= s_ij s_kj / s_jj (L = S D^-1/2 afterwards);
//   inversion      forward substitution on all columns of the identity at once: row j of X is final after step j-1,
//                  x_j /= l_jj is published, every later row subtracts l_ij x_j.
template <int JB>
__device__ __forceinline__ void chol_sweep16(double (&a)[8][8], double* colbuf, int tx, int ty, int jbase, int& bad) {
#pragma unroll 1
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * JB + jj;
    double* buf = colbuf + (j & 1) * FB;
    if (tx == jj) {
#pragma unroll
      for (int ia = JB; ia < 8; ++ia) buf[ty + 16 * ia] = a[ia][JB];
    }
    __syncthreads();
    double dj = buf[j];
    if (!(dj > 0.0)) {
      if (bad == 0) bad = jbase + j + 1;
      dj = 1.0;
    }
    const double rinv = 1.0 / dj;
    double ri[8];
#pragma unroll
    for (int ia = JB; ia < 8; ++ia) ri[ia] = buf[ty + 16 * ia] * rinv;
#pragma unroll
    for (int kb = JB; kb < 8; ++kb) {
      const double ck = buf[tx + 16 * kb];
      if (kb > JB || tx > jj) {
#pragma unroll
        for (int ia = kb; ia < 8; ++ia) a[ia][kb] = fma(-ri[ia], ck, a[ia][kb]);
      }
    }
  }
}

template <int JB>
__device__ __forceinline__ void inv_sweep16(double (&x)[8][8], const double* S, double* rowbuf, int tx, int ty) {
#pragma unroll 1
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * JB + jj;
    double* buf = rowbuf + (j & 1) * FB;
    if (ty == jj) {
      const double ljj = S[j * DS + j];
#pragma unroll
      for (int cb = 0; cb <= JB; ++cb) {
        x[JB][cb] = x[JB][cb] / ljj;
        buf[tx + 16 * cb] = x[JB][cb];
      }
    }
    __syncthreads();
    double xr[8];
#pragma unroll
    for (int cb = 0; cb <= JB; ++cb) xr[cb] = buf[tx + 16 * cb];
#pragma unroll
    for (int ia = JB; ia < 8; ++ia) {
      if (ia > JB || ty > jj) {
        const double lij = S[(ty + 16 * ia) * DS + j];
#pragma unroll
        for (int cb = 0; cb <= JB; ++cb) x[ia][cb] = fma(-lij, xr[cb], x[ia][cb]);
      }
    }
  }
}

__global__ void __launch_bounds__(FNT, 1) fit_potrf_diag_kernel(const FitDev* fits, int b, int n_pad) {
  extern __shared__ __align__(128) double sm[];
  double* S = sm;                // [128][129] L, then L^-1
  double* line = S + FB * DS;    // [2][128] published column / row
  double* sd = line + 2 * FB;    // [128] sqrt of the pivots
  const FitDev f = fits[blockIdx.x];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double* Ab = f.A + (size_t)b * FB * n_pad + (size_t)b * FB;
  double a[8][8];
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) a[ia][kb] = (kb <= ia) ? Ab[(size_t)(ty + 16 * ia) * n_pad + tx + 16 * kb] : 0.0;
  int bad = 0;
  chol_sweep16<0>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<1>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<2>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<3>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<4>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<5>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<6>(a, line, tx, ty, b * FB, bad);
  chol_sweep16<7>(a, line, tx, ty, b * FB, bad);
  if (tx == ty) {
#pragma unroll
    for (int ia = 0; ia < 8; ++ia) {
      double dj = a[ia][ia];
      if (!(dj > 0.0)) dj = 1.0;
      sd[ty + 16 * ia] = sqrt(dj);
    }
  }
  __syncthreads();
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      const int i = ty + 16 * ia, k = tx + 16 * kb;
      double v = 0.0;
      if (kb <= ia) v = (k < i) ? a[ia][kb] / sd[k] : (k == i ? sd[k] : 0.0);
      S[i * DS + k] = v;
      if (kb <= ia) Ab[(size_t)i * n_pad + k] = v;
    }
  if (tid < 32) {
    double s = 0.0;
    for (int k = tid; k < FB; k += 32) s += log(sd[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0) f.logdiag[b] = s;
  }
  if (bad && tid == 0 && *f.info == 0) *f.info = bad;   // every thread saw the same pivots
  __syncthreads();
  // inverse
  double x[8][8];
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) x[ia][cb] = (ia == cb && tx == ty) ? 1.0 : 0.0;
  inv_sweep16<0>(x, S, line, tx, ty);
  inv_sweep16<1>(x, S, line, tx, ty);
  inv_sweep16<2>(x, S, line, tx, ty);
  inv_sweep16<3>(x, S, line, tx, ty);
  inv_sweep16<4>(x, S, line, tx, ty);
  inv_sweep16<5>(x, S, line, tx, ty);
  inv_sweep16<6>(x, S, line, tx, ty);
  inv_sweep16<7>(x, S, line, tx, ty);
  __syncthreads();               // every read of L is done: S becomes L^-1
#pragma unroll
  for (int ia = 0; ia < 8; ++ia)
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
      const int i = ty + 16 * ia, c = tx + 16 * cb;
      S[i * DS + c] = (c <= i) ? x[ia][cb] : 0.0;
    }
  __syncthreads();
  double* Db = f.Dinv + (size_t)b * FB * FB;
  for (int idx = tid; idx < FB * FB; idx += FNT) Db[idx] = S[(idx >> 7) * DS + (idx & 127)];
  double* Fb = f.DinvFrag + (size_t)b * 136 * 64;
  for (int idx = tid; idx < 136 * 64; idx += FNT) {
    const int fi = idx >> 6, w = idx & 63, lane = w >> 1, e = w & 1;
    const int g = lane >> 2, t = lane & 3;
    Fb[idx] = S[(8 * c_frag_nt[fi] + g) * DS + 8 * c_frag_kt[fi] + 2 * t + e];
  }
}

// (-L)(i,b) = -A(i,b) Dinv_b^T for i > b.  grid (nblk-b-1, G).
__global__ void __launch_bounds__(FNT, 1) fit_panel_kernel(const FitDev* fits, int b, int n_pad) {
  extern __shared__ __align__(128) double sm[];
  const FitDev f = fits[blockIdx.y];
  const int i = b + 1 + blockIdx.x;
  double* T = f.A + (size_t)i * FB * n_pad + (size_t)b * FB;
  double acc[2][16][2];
  zero_acc(acc);
  gemm_nt(acc, T, n_pad, f.Dinv + (size_t)b * FB * FB, FB, FB / TRSM_BK, sm, threadIdx.x);
  store_tile<false>(acc, T, n_pad, -1.0, threadIdx.x);
}

// A(i,j) -= (-L)(i,b) (-L)(j,b)^T for b < j <= i.  grid (T(T+1)/2 with T = nblk-b-1, G).
__global__ void __launch_bounds__(FNT, 1) fit_syrk_kernel(const FitDev* fits, int b, int n_pad) {
  extern __shared__ __align__(128) double sm[];
  const FitDev f = fits[blockIdx.y];
  int ti = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= (int)blockIdx.x) ++ti;
  while (ti * (ti + 1) / 2 > (int)blockIdx.x) --ti;
  const int tj = blockIdx.x - ti * (ti + 1) / 2;
  const int i = b + 1 + ti, j = b + 1 + tj;
  double acc[2][16][2];
  zero_acc(acc);
  gemm_nt(acc, f.A + (size_t)i * FB * n_pad + (size_t)b * FB, n_pad, f.A + (size_t)j * FB * n_pad + (size_t)b * FB, n_pad,
          FB / TRSM_BK, sm, threadIdx.x);
  store_tile<true>(acc, f.A + (size_t)i * FB * n_pad + (size_t)j * FB, n_pad, -1.0, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------- alpha
// alpha = L^-T L^-1 y by blocked (left-looking) substitution with the inverted diagonal blocks; one CTA of 1024 threads per
// instance, every warp keeps four independent row streams in flight.  Also packs the staged blocks of the posterior kernel
// and y^T alpha.
constexpr int SNT = 1024;
constexpr int SNW = SNT / 32;

// out[r] (r = warp + 32 i, i < 4) = sum_{c < ncols} M[r][c] * v[c] for the 128 rows of a block; ncols a multiple of 128.
__device__ __forceinline__ void rows_dot4(const double* __restrict__ M, size_t ld, int ncols, const double* v, double (&out)[4],
                                          int warp, int lane) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int c0 = 0; c0 < ncols; c0 += 128) {
    const double4 vv = *reinterpret_cast<const double4*>(v + c0 + 4 * lane);
    double4 m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = *reinterpret_cast<const double4*>(M + (size_t)(warp + 32 * i) * ld + c0 + 4 * lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[i] = fma(m[i].x, vv.x, s[i]);
      s[i] = fma(m[i].y, vv.y, s[i]);
      s[i] = fma(m[i].z, vv.z, s[i]);
      s[i] = fma(m[i].w, vv.w, s[i]);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = s[i];
}

// part[warp][c] = sum over this warp's rows r (r = warp + 32 i < nrows) of M[r][c] * v[r], c = 4 lane .. 4 lane + 3 (128 columns).
__device__ __forceinline__ void cols_dot(const double* __restrict__ M, size_t ld, int nrows, const double* v, double* part, int warp,
                                         int lane) {
  double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
  int r = warp;
  for (; r + 96 < nrows; r += 128) {
    double4 m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = *reinterpret_cast<const double4*>(M + (size_t)(r + 32 * i) * ld + 4 * lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double vr = v[r + 32 * i];
      s.x = fma(m[i].x, vr, s.x);
      s.y = fma(m[i].y, vr, s.y);
      s.z = fma(m[i].z, vr, s.z);
      s.w = fma(m[i].w, vr, s.w);
    }
  }
  for (; r < nrows; r += 32) {
    const double4 m = *reinterpret_cast<const double4*>(M + (size_t)r * ld + 4 * lane);
    const double vr = v[r];
    s.x = fma(m.x, vr, s.x);
    s.y = fma(m.y, vr, s.y);
    s.z = fma(m.z, vr, s.z);
    s.w = fma(m.w, vr, s.w);
  }
  *reinterpret_cast<double4*>(part + warp * FB + 4 * lane) = s;
}

__global__ void __launch_bounds__(SNT, 1) fit_solve_kernel(const FitDev* fits, int n, int d, int n_pad, int pad) {
  extern __shared__ __align__(128) double sm[];
  double* rhs = sm;                 // [n_pad]  y, then z = L^-1 y, then alpha
  double* tmp = rhs + n_pad;        // [128]
  double* part = tmp + FB;          // [32][128]
  const FitDev f = fits[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nblk = n_pad / FB;
  for (int i = tid; i < n_pad; i += SNT) rhs[i] = (i >= pad) ? f.y[i - pad] : 0.0;
  __syncthreads();
  // forward: z_b = Dinv_b (y_b + sum_{k<b} (-L)(b,k) z_k)
  for (int b = 0; b < nblk; ++b) {
    double o[4];
    rows_dot4(f.A + (size_t)b * FB * n_pad, n_pad, b * FB, rhs, o, warp, lane);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tmp[warp + 32 * i] = rhs[b * FB + warp + 32 * i] + o[i];
    }
    __syncthreads();
    rows_dot4(f.Dinv + (size_t)b * FB * FB, FB, FB, tmp, o, warp, lane);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rhs[b * FB + warp + 32 * i] = o[i];
    }
    __syncthreads();
  }
  // backward: alpha_b = Dinv_b^T (z_b + sum_{k>b} (-L)(k,b)^T alpha_k)
  for (int b = nblk - 1; b >= 0; --b) {
    cols_dot(f.A + (size_t)(b + 1) * FB * n_pad + (size_t)b * FB, n_pad, (nblk - 1 - b) * FB, rhs + (b + 1) * FB, part, warp, lane);
    __syncthreads();
    if (tid < FB) {
      double s = rhs[b * FB + tid];
      for (int w = 0; w < SNW; ++w) s += part[w * FB + tid];
      tmp[tid] = s;
    }
    __syncthreads();
    cols_dot(f.Dinv + (size_t)b * FB * FB, FB, FB, tmp, part, warp, lane);
    __syncthreads();
    if (tid < FB) {
      double s = 0.0;
      for (int w = 0; w < SNW; ++w) s += part[w * FB + tid];
      rhs[b * FB + tid] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < n_pad; i += SNT) f.alpha[i] = rhs[i];
  if (f.StageBlk) {
    for (int idx = tid; idx < n_pad * (d + 1); idx += SNT) {
      const int blk = idx / (FB * (d + 1)), rem = idx - blk * FB * (d + 1);
      const int q = rem / FB, r = rem - q * FB;
      f.StageBlk[idx] = (q < d) ? f.Xs[(size_t)(blk * FB + r) * d + q] : rhs[blk * FB + r];
    }
  }
  // y^T alpha, fixed order
  double s = 0.0;
  for (int i = pad + tid; i < n_pad; i += SNT) s = fma(f.y[i - pad], rhs[i], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) tmp[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < SNW; ++w) tot += tmp[w];
    *f.quad = tot;
  }
}

// ---------------------------------------------------------------------------------------------- inverse factor and K^-1
// U = L^-T in strips of 32 rows (the rows of U are independent of one another): strip s of block row b holds
// U(b,b) = Dinv_b^T on the diagonal, then U(b,i) = [sum_{k=b}^{i-1} U(b,k) (-L)(i,k)^T] Dinv_i^T for i > b, left to right.
// grid (4 nblk, G): the chain of a strip is serial in i, so narrow strips shorten the critical path 4x and fill the SMs.
__global__ void __launch_bounds__(FNT, 1) fit_linvT_kernel(const FitDev* fits, int n_pad) {
  extern __shared__ __align__(128) double sm[];
  const FitDev f = fits[blockIdx.y];
  const int b = blockIdx.x >> 2, strip = blockIdx.x & 3, tid = threadIdx.x, nblk = n_pad / FB;
  double* Ub = f.U + ((size_t)b * FB + 32 * strip) * n_pad;     // first row of the strip
  {
    const double* Db = f.Dinv + (size_t)b * FB * FB;
    // U[R0 + r][b FB + c] = Dinv_b[c][32 strip + r]: transposed copy through shared memory, 32 x 32 at a time
    double* tile = sm;   // [32][33]
    for (int sub = 0; sub < 4; ++sub) {
      __syncthreads();
      for (int idx = tid; idx < 1024; idx += FNT) tile[(idx >> 5) * 33 + (idx & 31)] = Db[(size_t)(32 * sub + (idx >> 5)) * FB + 32 * strip + (idx & 31)];
      __syncthreads();
      for (int idx = tid; idx < 1024; idx += FNT) Ub[(size_t)(idx >> 5) * n_pad + b * FB + 32 * sub + (idx & 31)] = tile[(idx & 31) * 33 + (idx >> 5)];
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int i = b + 1; i < nblk; ++i) {
    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
    gemm_nt32(acc, Ub + (size_t)b * FB, n_pad, f.A + (size_t)i * FB * n_pad + (size_t)b * FB, n_pad, (i - b) * (FB / TRSM_BK), sm, tid);
    double* T = Ub + (size_t)i * FB;
    store_tile32(acc, T, n_pad, tid);
    __threadfence_block();
    __syncthreads();
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
    gemm_nt32(acc, T, n_pad, f.Dinv + (size_t)i * FB * FB, FB, FB / TRSM_BK, sm, tid);
    store_tile32(acc, T, n_pad, tid);
    __threadfence_block();
    __syncthreads();
  }
}

// K^-1 = U U^T: tile (a, b), a >= b: sum_{k >= a} U(a,k) U(b,k)^T, written (with its mirror image) into the unpadded n x n Winv.
__global__ void __launch_bounds__(FNT, 1) fit_winv_kernel(const FitDev* fits, int n, int n_pad, int pad) {
  extern __shared__ __align__(128) double sm[];
  const FitDev f = fits[blockIdx.y];
  int ta = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((ta + 1) * (ta + 2) / 2 <= (int)blockIdx.x) ++ta;
  while (ta * (ta + 1) / 2 > (int)blockIdx.x) --ta;
  const int tb = blockIdx.x - ta * (ta + 1) / 2, nblk = n_pad / FB;
  double acc[2][16][2];
  zero_acc(acc);
  gemm_nt(acc, f.U + (size_t)ta * FB * n_pad + (size_t)ta * FB, n_pad, f.U + (size_t)tb * FB * n_pad + (size_t)ta * FB, n_pad,
          (nblk - ta) * (FB / TRSM_BK), sm, threadIdx.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int R = ta * FB + warp * 16 + 8 * mi + g - pad, C = tb * FB + 8 * nt + 2 * t + e - pad;
        if (R >= 0 && C >= 0) {
          f.Winv[(size_t)R * n + C] = acc[mi][nt][e];
          if (ta != tb) f.Winv[(size_t)C * n + R] = acc[mi][nt][e];
        }
      }
}

// Linv[i][k] = U[pad+k][pad+i] for k <= i, 0 above the diagonal.  grid (ceil(n/32), ceil(n/32), G), 32 x 8 threads.
__global__ void fit_linv_kernel(const FitDev* fits, int n, int n_pad, int pad) {
  __shared__ double tile[32][33];
  const FitDev f = fits[blockIdx.z];
  const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, i = i0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < n && i < n && k <= i) ? f.U[(size_t)(pad + k) * n_pad + pad + i] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, k = k0 + threadIdx.x;
    if (i < n && k < n) f.Linv[(size_t)i * n + k] = tile[threadIdx.x][r];
  }
}

// Negated off-diagonal blocks are what the posterior kernels read; nothing to do.  The strictly-upper blocks of A are never
// read (they hold the unreferenced upper triangle of K).

// ---------------------------------------------------------------------------------------------- likelihood gradient
// partial[blk][0] = sum w K/variance, [1..d] = sum w (-(dK/dr)/r) dq^2 / ell_q, [d+1] = trace(w), w = 0.5 (alpha alpha^T - Winv)
// over a 32-row strip of the matrix.  grid (ceil(n/32), G), 256 threads.
__global__ void __launch_bounds__(256) lml_grad_kernel(const FitDev* fits, int n, int d, int n_pad, int pad, double* partial) {
  const FitDev f = fits[blockIdx.y];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * 32;
  double gv = 0.0, gn = 0.0, ge[MAX_D];
#pragma unroll
  for (int q = 0; q < MAX_D; ++q) ge[q] = 0.0;
  __shared__ double xi[32 * MAX_D], ai[32];
  for (int idx = tid; idx < 32 * d; idx += 256) {
    const int r = idx / d;
    xi[idx] = (i0 + r < n) ? f.Xs[(size_t)(pad + i0 + r) * d + (idx - r * d)] : 0.0;
  }
  if (tid < 32) ai[tid] = (i0 + tid < n) ? f.alpha[pad + i0 + tid] : 0.0;
  __syncthreads();
  for (int k = tid; k < n; k += 256) {
    double xk[MAX_D];
#pragma unroll
    for (int q = 0; q < MAX_D; ++q) xk[q] = (q < d) ? f.Xs[(size_t)(pad + k) * d + q] : 0.0;
    const double ak = f.alpha[pad + k];
    for (int r = 0; r < 32 && i0 + r < n; ++r) {
      const int i = i0 + r;
      const double w = 0.5 * (ai[r] * ak - f.Winv[(size_t)i * n + k]);
      if (i == k) {
        gv = fma(w, f.variance, gv);      // K_ii = variance (stationary.py:168-171); the reduction divides by it
        gn += w;
      } else {
        double r2 = 0.0, dq2[MAX_D];
#pragma unroll
        for (int q = 0; q < MAX_D; ++q) {
          const double dq = (q < d) ? xi[r * d + q] - xk[q] : 0.0;
          dq2[q] = dq * dq;
          r2 += dq2[q];
        }
        double kv, gg;
        kern_value_grad(f.kind, f.variance, r2, kv, gg);
        gv = fma(w, kv, gv);
        const double wg = -w * gg;
#pragma unroll
        for (int q = 0; q < MAX_D; ++q) ge[q] = fma(wg, dq2[q], ge[q]);
      }
    }
  }
  // block reduction in a fixed order
  __shared__ double red[8][MAX_D + 2];
  const int warp = tid >> 5, lane = tid & 31;
  auto wsum = [&](double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  gv = wsum(gv);
  gn = wsum(gn);
#pragma unroll
  for (int q = 0; q < MAX_D; ++q) ge[q] = wsum(ge[q]);
  if (lane == 0) {
    red[warp][0] = gv;
    red[warp][MAX_D + 1] = gn;
#pragma unroll
    for (int q = 0; q < MAX_D; ++q) red[warp][1 + q] = ge[q];
  }
  __syncthreads();
  if (tid < d + 2) {
    const int slot = (tid == d + 1) ? MAX_D + 1 : tid;
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][slot];
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (d + 2) + tid] = s;
  }
}

// out[g][0] = dL/d variance, [1..d] = dL/d ell_q, [d+1] = dL/d noise.  grid (G), 32 threads.
__global__ void lml_grad_reduce_kernel(const FitDev* fits, const double* partial, int nstrips, int d, double* out) {
  const FitDev f = fits[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid < d + 2) {
    double s = 0.0;
    for (int b = 0; b < nstrips; ++b) s += partial[((size_t)blockIdx.x * nstrips + b) * (d + 2) + tid];
    if (tid == 0) s /= f.variance;
    else if (tid <= d) s /= f.ell[tid - 1];
    out[(size_t)blockIdx.x * (d + 2) + tid] = s;
  }
}

bool g_tables_ready[64] = {};

int ensure_tables(bopl_handle* h) {
  if (h->device < 64 && g_tables_ready[h->device]) return BOPL_OK;
  unsigned char nt[136], kt[136];
  int fi = 0;
  for (int G = 3; G >= 0; --G)
    for (int k = 0; k <= 4 * G + 3; ++k)
      for (int u = 0; u < 4; ++u) {
        if (k > 4 * G + u) continue;
        nt[fi] = (unsigned char)(4 * G + u);
        kt[fi] = (unsigned char)k;
        ++fi;
      }
  BOPL_CUDA(h, cudaMemcpyToSymbol(c_frag_nt, nt, sizeof(nt)));
  BOPL_CUDA(h, cudaMemcpyToSymbol(c_frag_kt, kt, sizeof(kt)));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_linvT_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM32_SMEM));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_winv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
  BOPL_CUDA(h, cudaFuncSetAttribute(fit_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (h->device < 64) g_tables_ready[h->device] = true;
  return BOPL_OK;
}

}  // namespace

// Factorise G instances described by `fits` (device array; `host` is its host mirror).  Everything is queued on `st`.
int fit_factorize(bopl_handle* h, const void* fits_dev, int G, const double* X_dev, int n, int d, bool want_inverse,
                  bool want_winv, bool want_linv, cudaStream_t st) {
  const FitDev* fits = (const FitDev*)fits_dev;
  int rc = ensure_tables(h);
  if (rc) return rc;
  const int n_pad = ((n + FB - 1) / FB) * FB, pad = n_pad - n, nblk = n_pad / FB;
  if ((size_t)(n_pad + FB + SNW * FB) * sizeof(double) > 200 * 1024) return fail(h, BOPL_ERR_UNSUPPORTED, "device fit: n > 21000 training points");
  fit_prep_kernel<<<dim3(64, G), 256, 0, st>>>(fits, X_dev, n, d, n_pad, pad);
  BOPL_LAUNCH_CHECK(h, "fit_prep_kernel");
  fit_cov_kernel<<<dim3(nblk * (nblk + 1) / 2, G), 256, 0, st>>>(fits, d, n_pad, pad);
  BOPL_LAUNCH_CHECK(h, "fit_cov_kernel");
  for (int b = 0; b < nblk; ++b) {
    fit_potrf_diag_kernel<<<G, FNT, DIAG_SMEM, st>>>(fits, b, n_pad);
    BOPL_LAUNCH_CHECK(h, "fit_potrf_diag_kernel");
    const int T = nblk - b - 1;
    if (T > 0) {
      fit_panel_kernel<<<dim3(T, G), FNT, GEMM_SMEM, st>>>(fits, b, n_pad);
      BOPL_LAUNCH_CHECK(h, "fit_panel_kernel");
      fit_syrk_kernel<<<dim3(T * (T + 1) / 2, G), FNT, GEMM_SMEM, st>>>(fits, b, n_pad);
      BOPL_LAUNCH_CHECK(h, "fit_syrk_kernel");
    }
  }
  fit_solve_kernel<<<G, SNT, (size_t)(n_pad + FB + SNW * FB) * sizeof(double), st>>>(fits, n, d, n_pad, pad);
  BOPL_LAUNCH_CHECK(h, "fit_solve_kernel");
  if (want_inverse) {
    fit_linvT_kernel<<<dim3(4 * nblk, G), FNT, GEMM32_SMEM, st>>>(fits, n_pad);
    BOPL_LAUNCH_CHECK(h, "fit_linvT_kernel");
    if (want_winv) {
      fit_winv_kernel<<<dim3(nblk * (nblk + 1) / 2, G), FNT, GEMM_SMEM, st>>>(fits, n, n_pad, pad);
      BOPL_LAUNCH_CHECK(h, "fit_winv_kernel");
    }
    if (want_linv) {
      fit_linv_kernel<<<dim3((n + 31) / 32, (n + 31) / 32, G), dim3(32, 8), 0, st>>>(fits, n, n_pad, pad);
      BOPL_LAUNCH_CHECK(h, "fit_linv_kernel");
    }
  }
  return BOPL_OK;
}

int fit_lml_grad(bopl_handle* h, const void* fits_dev, int G, int n, int d, double* partial, double* out, cudaStream_t st) {
  const FitDev* fits = (const FitDev*)fits_dev;
  const int n_pad = ((n + FB - 1) / FB) * FB, pad = n_pad - n, nstrips = (n + 31) / 32;
  lml_grad_kernel<<<dim3(nstrips, G), 256, 0, st>>>(fits, n, d, n_pad, pad, partial);
  BOPL_LAUNCH_CHECK(h, "lml_grad_kernel");
  lml_grad_reduce_kernel<<<G, 32, 0, st>>>(fits, partial, nstrips, d, out);
  BOPL_LAUNCH_CHECK(h, "lml_grad_reduce_kernel");
  return BOPL_OK;
}

size_t fit_dev_struct_bytes() { return sizeof(FitDev); }

// Fill one FitDev on the host (the struct is private to this translation unit).
void fit_dev_fill(void* slot, double* A, double* Dinv, double* DinvFrag, double* U, double* Winv, double* Linv, double* Xs,
                  double* alpha, double* StageBlk, double* ell, const double* y, double* logdiag, double* quad, int* info,
                  double variance, double diag_add, int kind) {
  FitDev f{};
  f.A = A;
  f.Dinv = Dinv;
  f.DinvFrag = DinvFrag;
  f.U = U;
  f.Winv = Winv;
  f.Linv = Linv;
  f.Xs = Xs;
  f.alpha = alpha;
  f.StageBlk = StageBlk;
  f.ell = ell;
  f.y = y;
  f.logdiag = logdiag;
  f.quad = quad;
  f.info = info;
  f.variance = variance;
  f.diag_add = diag_add;
  f.kind = kind;
  memcpy(slot, &f, sizeof(f));
}

}  // namespace bopl
