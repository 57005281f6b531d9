// gp_kernels.cu — the non-TRSM GP kernels of the hot path:
//   kstar_kernel            K(Xc, X) (and g = dK/dr / r) materialised, for bopl_cross_cov and the gradient path
//   posterior_mean_kernel   mu = K(Xc, X) alpha + ymean without the triangular solve
//   beta_gemm_kernel        beta = -2 K(Xc, X) Winv                       (GPy/core/gp.py:474)
//   grad_contract_kernel    d mu / dx and d var / dx                      (GPy/kern/src/stationary.py:312-322)
// Reference: GPy/kern/src/stationary.py:104-171,227-254,312-342; GPy/core/gp.py:380-400,438-490;
// inference/latent_function_inference/posterior.py:299-305.
#include "pipeline.cuh"

namespace bopl {

namespace {

// --------------------------------------------------------------------------------------------- K*, G
// Block: 16 candidates x all training rows.  Thread i-strided over training rows; writes are coalesced along n.
constexpr int KS_TC = 16;
constexpr int KS_NT = 256;

template <bool WITH_G>
__global__ void __launch_bounds__(KS_NT) kstar_kernel(GpDev gp, const double* __restrict__ Xc, int N, int d, int n, int pad,
                                                       double* __restrict__ K, double* __restrict__ G, int ldk) {
  __shared__ double xc_s[KS_TC * MAX_D];
  const int c0 = blockIdx.x * KS_TC;
  for (int idx = threadIdx.x; idx < KS_TC * d; idx += KS_NT) {
    int c = idx / d, q = idx - c * d;
    int cc = min(c0 + c, N - 1);
    xc_s[c * d + q] = Xc[(size_t)cc * d + q] / gp.ell[q];
  }
  __syncthreads();
  const double* Xs = gp.Xs + (size_t)pad * d;
  for (int i = blockIdx.y * KS_NT + threadIdx.x; i < n; i += gridDim.y * KS_NT) {
    double xt[MAX_D];
#pragma unroll
    for (int q = 0; q < MAX_D; ++q)
      if (q < d) xt[q] = Xs[(size_t)i * d + q];
    for (int c = 0; c < KS_TC && c0 + c < N; ++c) {
      double r2 = 0.0;
#pragma unroll
      for (int q = 0; q < MAX_D; ++q)
        if (q < d) {
          double df = xt[q] - xc_s[c * d + q];
          r2 = fma(df, df, r2);
        }
      if (WITH_G) {
        double k, g;
        kern_value_grad(gp.kind, gp.variance, r2, k, g);
        K[(size_t)(c0 + c) * ldk + i] = k;
        G[(size_t)(c0 + c) * n + i] = g;
      } else {
        K[(size_t)(c0 + c) * ldk + i] = kern_value(gp.kind, gp.variance, r2);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------- mean only
// Block: 32 candidates; 256 threads = 32 candidates x 8 row groups.  Training rows are staged through shared
// memory 128 at a time; the 8 row-group partial sums are combined in a fixed order (bit-stable run to run).
constexpr int PM_TC = 32;
constexpr int PM_RG = 8;
constexpr int PM_ROWS = 128;

struct MeanParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  int H, hsel, m, N, d, n, pad;
  long long out_stride;   // elements between attributes in `mean`
};

__global__ void __launch_bounds__(PM_TC* PM_RG) posterior_mean_kernel(MeanParams p) {
  __shared__ double xc_s[MAX_D * PM_TC];
  __shared__ double xt_s[PM_ROWS * MAX_D];
  __shared__ double al_s[PM_ROWS];
  __shared__ double red[PM_RG * PM_TC];
  const int j = blockIdx.y;
  const GpDev gp = p.gps[j * p.H + p.hsel];
  const int d = p.d, n = p.n;
  const int c = threadIdx.x & (PM_TC - 1), rg = threadIdx.x / PM_TC;
  const int c0 = blockIdx.x * PM_TC;
  for (int idx = threadIdx.x; idx < PM_TC * d; idx += PM_TC * PM_RG) {
    int cc = idx / d, q = idx - cc * d;
    int cg = min(c0 + cc, p.N - 1);
    xc_s[q * PM_TC + cc] = p.Xc[(size_t)cg * d + q] / gp.ell[q];
  }
  const double* Xs = gp.Xs + (size_t)p.pad * d;
  const double* al = gp.alpha + p.pad;
  double acc = 0.0;
  for (int r0 = 0; r0 < n; r0 += PM_ROWS) {
    const int rows = min(PM_ROWS, n - r0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * d; idx += PM_TC * PM_RG) xt_s[idx] = Xs[(size_t)r0 * d + idx];
    if (threadIdx.x < rows) al_s[threadIdx.x] = al[r0 + threadIdx.x];
    __syncthreads();
    for (int i = rg; i < rows; i += PM_RG) {
      double r2 = 0.0;
      for (int q = 0; q < d; ++q) {
        double df = xt_s[i * d + q] - xc_s[q * PM_TC + c];
        r2 = fma(df, df, r2);
      }
      acc = fma(kern_value(gp.kind, gp.variance, r2), al_s[i], acc);
    }
  }
  red[rg * PM_TC + c] = acc;
  __syncthreads();
  if (rg == 0 && c0 + c < p.N) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < PM_RG; ++g) s += red[g * PM_TC + c];
    p.mean[(size_t)j * p.out_stride + c0 + c] = s + gp.ymean;
  }
}

// --------------------------------------------------------------------------------------------- beta = -2 K* Winv
// B[Nc x n] = -2 * K[Nc x n] * Winv[n x n] on the FP64 tensor pipe (DMMA.8x8x4), for every batch size: an output
// element's accumulation order does not depend on the other rows of the batch, so a candidate's gradient is bit-identical
// whether it is evaluated alone (the L-BFGS call shape) or inside a batch (the lock-step multi-start driver relies on it).
// beta = -2 K* Winv (GPy/core/gp.py:474) as a DMMA tile product in the form of posterior_trsm_t.cu's main loop: CTA tile 128 candidates x
// 128 columns, k-step 16, eight warps each owning 16 candidates x all 128 columns (2 x 16 DMMA.8x8x4 accumulator tiles), operands through
// a 3-stage cp.async ring of 128 x 16 tiles in the 16-byte XOR-swizzled layout (conflict-free LDS.128 fragments; the two halves of a
// double2 feed two DMMAs with the k-sets {0,2,..} / {1,3,..} on both operands).  Operands are PADDED copies: K* [chunk x ldk] with zero
// columns up to a multiple of 16 (kstar_kernel writes it that way), Winv [rows up to a multiple of 128 x ldk] (winv_pad_kernel) — Winv is
// symmetric, so its row j is column j read K-major.  Per output element the k order is fixed: a candidate's beta is bit-identical in
// any batch above the small-batch limit.  Round 1's 64 x 64 tile kernel with synchronous loads ran at 0.33 of the DGEMM peak.
constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_NT = 256, GM_STAGES = 3;
constexpr int GM_TILE = GM_BM * GM_BK;                    // doubles per operand tile
constexpr size_t GM_SMEM = (size_t)GM_STAGES * 2 * GM_TILE * sizeof(double);

__global__ void winv_pad_kernel(const double* __restrict__ W, int n, double* __restrict__ WP, int rows, int ldk) {
  const size_t total = (size_t)rows * ldk;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / ldk), c = (int)(e - (size_t)r * ldk);
    WP[e] = (r < n && c < n) ? W[(size_t)r * n + c] : 0.0;
  }
}

__global__ void __launch_bounds__(GM_NT, 1) beta_gemm_kernel(const double* __restrict__ WP, const double* __restrict__ K, int nc, int n,
                                                             int ldk, double* __restrict__ B) {
  extern __shared__ __align__(128) double gsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int Cw = warp * 16, swz = (g & 1) << 2;
  const int r0 = blockIdx.y * GM_BM, col0 = blockIdx.x * GM_BN;
  const double* Arow = K + (size_t)r0 * ldk;              // candidates r0 .. r0 + 127 (the K* buffer holds whole 128-row tiles)
  const double* Brow = WP + (size_t)col0 * ldk;           // columns col0 .. col0 + 127 as rows of the symmetric matrix
  const int nsteps = ldk / GM_BK;
  double acc[2][16][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
  auto issue = [&](int it) {
    double* As = gsm + (size_t)(it % GM_STAGES) * 2 * GM_TILE;
    load_tile<GM_BM, GM_BK, GM_NT>(As, Arow + (size_t)it * GM_BK, ldk, tid);
    load_tile<GM_BN, GM_BK, GM_NT>(As + GM_TILE, Brow + (size_t)it * GM_BK, ldk, tid);
  };
#pragma unroll
  for (int s = 0; s < GM_STAGES - 1; ++s) {
    if (s < nsteps) issue(s);
    cp_async_commit();
  }
  for (int it = 0; it < nsteps; ++it) {
    cp_async_wait<GM_STAGES - 2>();
    __syncthreads();                       // tile `it` has landed for every thread; every warp is past tile it - 1: its buffer may be refilled
    if (it + GM_STAGES - 1 < nsteps) issue(it + GM_STAGES - 1);
    cp_async_commit();
    const double* As = gsm + (size_t)(it % GM_STAGES) * 2 * GM_TILE;
    const double* Bs = As + GM_TILE;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ch = ((t + 4 * kk) ^ swz) << 1;
      double2 a[2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) a[mi] = *reinterpret_cast<const double2*>(As + (Cw + 8 * mi + g) * GM_BK + ch);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        double2 b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) b[u] = *reinterpret_cast<const double2*>(Bs + (8 * (8 * hf + u) + g) * GM_BK + ch);
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
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int r = r0 + Cw + 8 * mi + g;
    if (r >= nc) continue;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cidx = col0 + 8 * nt + 2 * t + e;
        if (cidx < n) B[(size_t)r * n + cidx] = -2.0 * acc[mi][nt][e];
      }
  }
}

// --------------------------------------------------------------------------------------------- gradient contraction
// One block per (candidate, attribute):  out[q] = ( sum_i coef_i * G_i * (x_q - X_iq) ) / ell_q^2
// with coef = alpha (mean gradient) and coef = beta (variance gradient).  Fixed reduction order.
constexpr int GC_NT = 128;
__global__ void __launch_bounds__(GC_NT) grad_contract_kernel(GpDev gp, const double* __restrict__ Xtrain, const double* __restrict__ Xc,
                                                               int d, int n, int pad, const double* __restrict__ G,
                                                               const double* __restrict__ Bt, double* __restrict__ dmean,
                                                               double* __restrict__ dvar) {
  __shared__ double red[2][GC_NT / 32][MAX_D];
  const int c = blockIdx.x;
  const double* g = G + (size_t)c * n;
  const double* bt = Bt ? Bt + (size_t)c * n : nullptr;
  const double* al = gp.alpha + pad;
  double xq[MAX_D], am[MAX_D], av[MAX_D];
#pragma unroll
  for (int q = 0; q < MAX_D; ++q) {
    xq[q] = (q < d) ? Xc[(size_t)c * d + q] : 0.0;
    am[q] = 0.0;
    av[q] = 0.0;
  }
  for (int i = threadIdx.x; i < n; i += GC_NT) {
    double gi = g[i];
    double cm = gi * al[i];            // tmp = invdist * dK_dr * dL_dK  (stationary.py:318)
    double cv = bt ? gi * bt[i] : 0.0;
#pragma unroll
    for (int q = 0; q < MAX_D; ++q)
      if (q < d) {
        double df = xq[q] - Xtrain[(size_t)i * d + q];
        am[q] = fma(cm, df, am[q]);
        av[q] = fma(cv, df, av[q]);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < MAX_D; ++q)
    if (q < d) {
      double a = am[q], v = av[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
      }
      if (lane == 0) {
        red[0][warp][q] = a;
        red[1][warp][q] = v;
      }
    }
  __syncthreads();
  if (threadIdx.x < d) {
    const int q = threadIdx.x;
    double a = 0.0, v = 0.0;
#pragma unroll
    for (int w = 0; w < GC_NT / 32; ++w) {
      a += red[0][w][q];
      v += red[1][w][q];
    }
    double l2 = gp.ell[q] * gp.ell[q];
    if (dmean) dmean[(size_t)c * d + q] = a / l2;
    if (dvar) dvar[(size_t)c * d + q] = v / l2;
  }
}

// --------------------------------------------------------------------------------------------- small-batch path
// N <= SMALL_N candidates (the L-BFGS inner loop: 1 point; the lock-step multi-start driver: one point per anchor).
// The blocked TRSM kernel needs a full item time (~2.4 ms at n = 2000) however few candidates there are; here the solve
// is a matrix-vector product with the explicit inverse factor Linv = L^-1 (LAPACK dtrtri on the host, once per model
// update): V = Linv K*, every training row handled by one warp, all attributes in one launch.  Same for
// beta = -2 Winv K*.  Everything is per-(candidate, row) deterministic, so values are bit-identical whatever else is in
// the batch.  ~6 short launches per value+gradient call instead of a 2.4 ms persistent kernel.
struct SmallParams {
  const GpDev* gps;
  const double* Xc;      // [N*d]
  const double* Xtrain;  // [n*d] unscaled
  double* K;             // [m][N][n]
  double* G;             // [m][N][n]
  double* V;             // [m][N][n]  Linv K*
  double* B;             // [m][N][n]  -2 Winv K*
  double* mean;          // [m*N]
  double* var;           // [m*N]
  double* dmean;         // [m*N*d]
  double* dvar;          // [m*N*d]
  int H, hsel, m, N, d, n, pad, flags;
};

template <bool WITH_G>
__global__ void __launch_bounds__(256) small_kstar_kernel(SmallParams p) {
  __shared__ double xc_s[SMALL_N * MAX_D];
  const int j = blockIdx.y;
  const GpDev gp = p.gps[j * p.H + p.hsel];
  const int d = p.d, n = p.n;
  for (int idx = threadIdx.x; idx < p.N * d; idx += 256) xc_s[idx] = p.Xc[idx] / gp.ell[idx % d];
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double* xs = gp.Xs + (size_t)(p.pad + i) * d;
  double xt[MAX_D];
#pragma unroll
  for (int q = 0; q < MAX_D; ++q)
    if (q < d) xt[q] = xs[q];
  for (int c = 0; c < p.N; ++c) {
    double r2 = 0.0;
#pragma unroll
    for (int q = 0; q < MAX_D; ++q)
      if (q < d) {
        double df = xt[q] - xc_s[c * d + q];
        r2 = fma(df, df, r2);
      }
    const size_t o = ((size_t)j * p.N + c) * n + i;
    if (WITH_G) {
      double k, g;
      kern_value_grad(gp.kind, gp.variance, r2, k, g);
      p.K[o] = k;
      p.G[o] = g;
    } else {
      p.K[o] = kern_value(gp.kind, gp.variance, r2);
    }
  }
}

// out[j][c][i] = scale * sum_k M_j[i][k] K[j][c][k];  TRI: M = Linv (k <= i), scale 1;  else M = Winv, scale -2.
// A skinny matrix product on DMMA.8x8x4, bound by streaming M once (32 MB per attribute at n = 2000): CTA = 32 rows of M x up to 32
// candidates, four warps of 8 rows each; 32-deep k-tiles of M (coalesced 8-byte cp.async, zero-filled past n) and of the candidates' K* rows
// go through a double-buffered shared-memory ring; fragments are read with a 36-double row pitch (two wavefronts per 8-byte warp load, the
// minimum).  Per (candidate, row) the k order is fixed and does not depend on the other candidates: bit-identical alone or in a batch.
// Round 1's form (a warp per row, K* through shared memory for every FMA: one 8-byte shared load per fp64 operation) took 184 + 286 us for
// the two products of a 20-anchor L-BFGS step at n = 2000 — 15x the time the 144 MB of factors need to stream.
constexpr int GV_ROWS = 32, GV_KT = 32, GV_PITCH = 36, GV_NT = 128;
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gsrc, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes));
}

template <bool TRI, int NT>      // NT tiles of 8 candidates
__global__ void __launch_bounds__(GV_NT) small_gemv_kernel(SmallParams p) {
  constexpr int NC = 8 * NT;
  __shared__ __align__(16) double Ms[2][GV_ROWS][GV_PITCH];
  __shared__ __align__(16) double Ks[2][NC][GV_PITCH];
  const int j = blockIdx.y;
  const GpDev& gp = p.gps[j * p.H + p.hsel];
  const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int i0 = blockIdx.x * GV_ROWS;
  const double* M = TRI ? gp.Linv : gp.Winv;
  const double* Kj = p.K + (size_t)j * p.N * n;
  const int kend = TRI ? min(i0 + GV_ROWS, n) : n;          // the block's last row bounds the k-range of a triangular factor
  const int ntiles = (kend + GV_KT - 1) / GV_KT;
  auto issue = [&](int tile) {
    const int k0 = tile * GV_KT, buf = tile & 1;
    for (int idx = tid; idx < GV_ROWS * GV_KT; idx += GV_NT) {
      const int r = idx / GV_KT, kk = idx - r * GV_KT;
      const bool ok = (i0 + r < n) && (k0 + kk < n);
      cp_async8_zfill(&Ms[buf][r][kk], M + (size_t)min(i0 + r, n - 1) * n + min(k0 + kk, n - 1), ok);
    }
    for (int idx = tid; idx < NC * GV_KT; idx += GV_NT) {
      const int c = idx / GV_KT, kk = idx - c * GV_KT;
      const bool ok = (c < p.N) && (k0 + kk < n);
      cp_async8_zfill(&Ks[buf][c][kk], Kj + (size_t)min(c, p.N - 1) * n + min(k0 + kk, n - 1), ok);
    }
  };
  double acc[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
  const int row = i0 + 8 * warp + g;                        // this lane's row of M in the A fragments
  issue(0);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    if (tile + 1 < ntiles) issue(tile + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int buf = tile & 1, k0 = tile * GV_KT;
#pragma unroll
    for (int kk = 0; kk < GV_KT; kk += 4) {
      double a = Ms[buf][8 * warp + g][kk + t];
      if (TRI && k0 + kk + t > row) a = 0.0;               // the strictly upper part of a triangular factor is never used, whatever it holds
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma(acc[nt][0], acc[nt][1], a, Ks[buf][8 * nt + g][kk + t]);
    }
    __syncthreads();                                        // the buffer is refilled two tiles later: every warp must be done with it
  }
  double* out = (TRI ? p.V : p.B) + (size_t)j * p.N * n;
  const double scale = TRI ? 1.0 : -2.0;
  if (row < n) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * nt + 2 * t + e;
        if (c < p.N) out[(size_t)c * n + row] = scale * acc[nt][e];
      }
  }
}

template <bool TRI>
void launch_small_gemv(const SmallParams& p, int n, int m, cudaStream_t st) {
  dim3 grid((n + GV_ROWS - 1) / GV_ROWS, m);
  if (p.N <= 8) small_gemv_kernel<TRI, 1><<<grid, GV_NT, 0, st>>>(p);
  else if (p.N <= 16) small_gemv_kernel<TRI, 2><<<grid, GV_NT, 0, st>>>(p);
  else small_gemv_kernel<TRI, 4><<<grid, GV_NT, 0, st>>>(p);
}

// mean = K* alpha + ymean, var = variance - sum V^2 (+noise, clip).  One block per (candidate, attribute), fixed order.
__global__ void __launch_bounds__(256) small_reduce_kernel(SmallParams p) {
  __shared__ double red[2][8];
  const int c = blockIdx.x, j = blockIdx.y;
  const GpDev& gp = p.gps[j * p.H + p.hsel];
  const int n = p.n;
  const double* K = p.K + ((size_t)j * p.N + c) * n;
  const double* V = p.V + ((size_t)j * p.N + c) * n;
  const double* al = gp.alpha + p.pad;
  double a = 0.0, s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    a = fma(K[i], al[i], a);
    if (p.var) {
      const double v = V[i];
      s = fma(v, v, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double aa = 0.0, sv = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      aa += red[0][w];
      sv += red[1][w];
    }
    if (p.mean) p.mean[(size_t)j * p.N + c] = aa + gp.ymean;
    if (p.var) {
      double v = gp.variance - sv;
      if (p.flags & BOPL_VAR_INCLUDE_NOISE) v = gp.noise + v;
      if (p.flags & BOPL_VAR_CLIP) v = fmax(v, 1e-10);
      p.var[(size_t)j * p.N + c] = v;
    }
  }
}

// gradient contraction, all attributes in one launch (same arithmetic as grad_contract_kernel)
__global__ void __launch_bounds__(GC_NT) small_contract_kernel(SmallParams p) {
  __shared__ double red[2][GC_NT / 32][MAX_D];
  const int c = blockIdx.x, j = blockIdx.y;
  const GpDev& gp = p.gps[j * p.H + p.hsel];
  const int n = p.n, d = p.d;
  const double* g = p.G + ((size_t)j * p.N + c) * n;
  const double* bt = p.dvar ? p.B + ((size_t)j * p.N + c) * n : nullptr;
  const double* al = gp.alpha + p.pad;
  double xq[MAX_D], am[MAX_D], av[MAX_D];
#pragma unroll
  for (int q = 0; q < MAX_D; ++q) {
    xq[q] = (q < d) ? p.Xc[(size_t)c * d + q] : 0.0;
    am[q] = 0.0;
    av[q] = 0.0;
  }
  for (int i = threadIdx.x; i < n; i += GC_NT) {
    double gi = g[i];
    double cm = gi * al[i];
    double cv = bt ? gi * bt[i] : 0.0;
#pragma unroll
    for (int q = 0; q < MAX_D; ++q)
      if (q < d) {
        double df = xq[q] - p.Xtrain[(size_t)i * d + q];
        am[q] = fma(cm, df, am[q]);
        av[q] = fma(cv, df, av[q]);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < MAX_D; ++q)
    if (q < d) {
      double a = am[q], v = av[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
      }
      if (lane == 0) {
        red[0][warp][q] = a;
        red[1][warp][q] = v;
      }
    }
  __syncthreads();
  if (threadIdx.x < d) {
    const int q = threadIdx.x;
    double a = 0.0, v = 0.0;
#pragma unroll
    for (int w = 0; w < GC_NT / 32; ++w) {
      a += red[0][w][q];
      v += red[1][w][q];
    }
    double l2 = gp.ell[q] * gp.ell[q];
    if (p.dmean) p.dmean[((size_t)j * p.N + c) * d + q] = a / l2;
    if (p.dvar) p.dvar[((size_t)j * p.N + c) * d + q] = v / l2;
  }
}

}  // namespace

int launch_cross_cov(bopl_handle* h, const double* Xc, int N, int hsel, int j, double* K, cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  const GpDev& gp = h->gp_host[j * h->H + hsel];
  dim3 grid((N + KS_TC - 1) / KS_TC, std::max(1, std::min(8, (h->n + KS_NT - 1) / KS_NT)));
  kstar_kernel<false><<<grid, KS_NT, 0, st>>>(gp, Xc, N, h->d, h->n, h->pad, K, nullptr, h->n);
  BOPL_LAUNCH_CHECK(h, "kstar_kernel");
  return BOPL_OK;
}

int launch_posterior_mean_strided(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, long long out_stride,
                                  cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  MeanParams p;
  p.gps = h->gp_dev;
  p.Xc = Xc;
  p.mean = mean;
  p.H = h->H;
  p.hsel = hsel;
  p.m = h->m;
  p.N = N;
  p.d = h->d;
  p.n = h->n;
  p.pad = h->pad;
  p.out_stride = out_stride;
  dim3 grid((N + PM_TC - 1) / PM_TC, h->m);
  posterior_mean_kernel<<<grid, PM_TC * PM_RG, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "posterior_mean_kernel");
  return BOPL_OK;
}

int launch_posterior_mean(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, cudaStream_t st) {
  return launch_posterior_mean_strided(h, Xc, N, hsel, mean, N, st);
}

// dmean, dvar [m*N*d].  Scratch: K, G, B  [chunk x n] each, reused per attribute.
// Scratch of the small-batch path: K | G | V | B, each [m][SMALL_N][n].
static int small_setup(bopl_handle* h, const double* Xc, int N, int hsel, SmallParams& p) {
  const size_t per = (size_t)h->m * SMALL_N * h->n;
  int rc = ensure_bytes(h, (void**)&h->small_scratch, &h->small_scratch_bytes, 4 * per * sizeof(double));
  if (rc) return rc;
  p = SmallParams{};
  p.gps = h->gp_dev;
  p.Xc = Xc;
  p.Xtrain = h->X_dev;
  p.K = h->small_scratch;
  p.G = p.K + per;
  p.V = p.G + per;
  p.B = p.V + per;
  p.H = h->H;
  p.hsel = hsel;
  p.m = h->m;
  p.N = N;
  p.d = h->d;
  p.n = h->n;
  p.pad = h->pad;
  return BOPL_OK;
}

bool use_small_path(const bopl_handle* h, int N) { return h->has_linv && h->small_path && N > 0 && N <= SMALL_N; }

int launch_posterior_small(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                           cudaStream_t st) {
  SmallParams p;
  int rc = small_setup(h, Xc, N, hsel, p);
  if (rc) return rc;
  p.mean = mean;
  p.var = var;
  p.flags = flags;
  // value + gradient entry points (small_grad_follows): K* and (dK/dr)/r are generated once, the gradient half of the call reuses them
  if (h->small_grad_follows) small_kstar_kernel<true><<<dim3((h->n + 255) / 256, h->m), 256, 0, st>>>(p);
  else small_kstar_kernel<false><<<dim3((h->n + 255) / 256, h->m), 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "small_kstar_kernel");
  if (var) {
    launch_small_gemv<true>(p, h->n, h->m, st);
    BOPL_LAUNCH_CHECK(h, "small_gemv_kernel<Linv>");
  }
  small_reduce_kernel<<<dim3(N, h->m), 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "small_reduce_kernel");
  return BOPL_OK;
}

// V0 = Linv K*(x) of ONE point for every attribute (the sample-path step of path_kernels.cu); K0, V0 -> [m][n] scratch rows.
int launch_small_solve1(bopl_handle* h, const double* x_dev, int hsel, const double** K0, const double** V0, cudaStream_t st) {
  SmallParams p;
  int rc = small_setup(h, x_dev, 1, hsel, p);
  if (rc) return rc;
  small_kstar_kernel<false><<<dim3((h->n + 255) / 256, h->m), 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "small_kstar_kernel");
  launch_small_gemv<true>(p, h->n, h->m, st);
  BOPL_LAUNCH_CHECK(h, "small_gemv_kernel<Linv>");
  *K0 = p.K;
  *V0 = p.V;
  return BOPL_OK;
}

static int launch_posterior_grad_small(bopl_handle* h, const double* Xc, int N, int hsel, double* dmean, double* dvar,
                                       cudaStream_t st) {
  SmallParams p;
  int rc = small_setup(h, Xc, N, hsel, p);
  if (rc) return rc;
  p.dmean = dmean;
  p.dvar = dvar;
  if (!h->small_grad_follows) {        // standalone gradient call: nothing has generated K*, G for these candidates
    small_kstar_kernel<true><<<dim3((h->n + 255) / 256, h->m), 256, 0, st>>>(p);
    BOPL_LAUNCH_CHECK(h, "small_kstar_kernel<G>");
  }
  if (dvar) {
    launch_small_gemv<false>(p, h->n, h->m, st);
    BOPL_LAUNCH_CHECK(h, "small_gemv_kernel<Winv>");
  }
  small_contract_kernel<<<dim3(N, h->m), GC_NT, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "small_contract_kernel");
  return BOPL_OK;
}

int launch_posterior_grad(bopl_handle* h, const double* Xc, int N, int hsel, double* dmean, double* dvar,
                          cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  if (use_small_path(h, N)) return launch_posterior_grad_small(h, Xc, N, hsel, dmean, dvar, st);
  const int n = h->n, d = h->d;
  const int ldk = (n + GM_BK - 1) / GM_BK * GM_BK, wrows = (n + GM_BN - 1) / GM_BN * GM_BN;
  // candidates per pass: whole 128-row tiles of the beta GEMM, and a tile count that makes the GEMM grid (row tiles x column tiles) a whole
  // number of waves of one CTA per SM (n = 2000 on 148 SMs: 37 x 16 = 592 CTAs = 4 waves; 16 x 16 = 256 was 1.73)
  int row_tiles = 16;
  {
    const int ctiles = wrows / GM_BN;
    int a = h->sms, b = ctiles;
    while (b) {
      const int t = a % b;
      a = b;
      b = t;
    }
    const int rt = h->sms / a;
    if (rt * GM_BM <= 8192) row_tiles = rt;
  }
  const int chunk = std::min(row_tiles * GM_BM, (N + GM_BM - 1) / GM_BM * GM_BM);      // never more rows than the batch has (whole tiles)
  const size_t kb_cnt = (size_t)chunk * ldk, gb_cnt = (size_t)chunk * n, wp_cnt = dvar ? (size_t)wrows * ldk : 0;
  size_t need = (kb_cnt + 2 * gb_cnt + wp_cnt) * sizeof(double);
  int rc = ensure_bytes(h, (void**)&h->grad_scratch, &h->grad_scratch_bytes, need);
  if (rc) return rc;
  double* Kb = h->grad_scratch;                                    // K* [chunk x ldk], zero beyond column n
  double* Gb = Kb + kb_cnt;
  double* Bb = Gb + gb_cnt;
  double* WP = Bb + gb_cnt;                                        // zero-padded copy of the running attribute's Winv
  BOPL_CUDA(h, cudaMemsetAsync(Kb, 0, kb_cnt * sizeof(double), st));
  if (dvar && !h->gemm_attr_set) {
    BOPL_CUDA(h, cudaFuncSetAttribute(beta_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
    h->gemm_attr_set = true;
  }
  for (int j = 0; j < h->m; ++j) {
    const GpDev& gp = h->gp_host[j * h->H + hsel];
    if (dvar) {
      winv_pad_kernel<<<h->sms * 4, 256, 0, st>>>(gp.Winv, n, WP, wrows, ldk);
      BOPL_LAUNCH_CHECK(h, "winv_pad_kernel");
    }
    for (int s = 0; s < N; s += chunk) {
      const int nc = std::min(chunk, N - s);
      const double* xc = Xc + (size_t)s * d;
      dim3 gridk((nc + KS_TC - 1) / KS_TC, std::max(1, std::min(8, (n + KS_NT - 1) / KS_NT)));
      kstar_kernel<true><<<gridk, KS_NT, 0, st>>>(gp, xc, nc, d, n, h->pad, Kb, Gb, ldk);
      BOPL_LAUNCH_CHECK(h, "kstar_kernel<G>");
      if (dvar) {
        dim3 gg(wrows / GM_BN, (nc + GM_BM - 1) / GM_BM);
        beta_gemm_kernel<<<gg, GM_NT, GM_SMEM, st>>>(WP, Kb, nc, n, ldk, Bb);
        BOPL_LAUNCH_CHECK(h, "beta_gemm_kernel");
      }
      grad_contract_kernel<<<nc, GC_NT, 0, st>>>(gp, h->X_dev, xc, d, n, h->pad, Gb, dvar ? Bb : nullptr,
                                                 dmean ? dmean + ((size_t)j * N + s) * d : nullptr,
                                                 dvar ? dvar + ((size_t)j * N + s) * d : nullptr);
      BOPL_LAUNCH_CHECK(h, "grad_contract_kernel");
    }
  }
  return BOPL_OK;
}

}  // namespace bopl
