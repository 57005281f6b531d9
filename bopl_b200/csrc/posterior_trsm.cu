// posterior_trsm.cu — kernels (1)+(2) of the hot path, fused: per attribute, for a tile of 128 candidates,
//   K* = K(X, Xc)                      (generated on the fly, never written to HBM)
//   mu = K*^T alpha + ymean             (GPy posterior.py:299-305, gp.py:398-399)
//   V  = L^{-1} K*                      (LAPACK dtrtrs in the reference: posterior.py:311, linalg.py:92-111)
//   var = variance - colsum(V^2) (+noise, clip)   (posterior.py:312, gp.py:416-417, gpmodel.py:214)
//
// The triangular solve is a LEFT-LOOKING BLOCKED forward substitution over 128-row blocks of L:
//   V_ib = Dinv_ib * ( K*_ib + sum_{k<ib} (-L_ib,k) V_k ),   Dinv_ib = L_ib,ib^{-1} (precomputed per model update)
// so that all flops are GEMM-shaped and run on the FP64 tensor pipe (DMMA.8x8x4 via mma.sync.m8n8k4.f64 — the
// only FP64 MMA shape sm_100a SASS has; tcgen05 has no f64 kind).  CTA tile: 128 rows x 128 candidates, 8 warps
// (2 x 4), warp tile 64 x 32 = 32 DMMA accumulators.  L tiles and the already-solved V panel (candidate-major,
// per-CTA scratch that stays in L2) stream through a 4-stage cp.async pipeline of 128x16 / 128x16 tiles with a
// 16-byte XOR swizzle that makes every fragment LDS.128 conflict-free.  One persistent CTA per SM walks
// (attribute, candidate-tile) items, attribute-major so that all CTAs stream the same L out of L2.
#include "common.cuh"

namespace bopl {

namespace {

constexpr int NT = 256;
constexpr int STAGES = 4;
constexpr int DSTAGES = 2;
constexpr int TILE_DBL = TRSM_BM * TRSM_BK;                 // 2048 doubles = 16 KB
constexpr int P_DBL = STAGES * 2 * TILE_DBL;                // 16384 doubles = 128 KB (pipeline | K* tile | rhs tile)
constexpr int Q_DBL = DSTAGES * TILE_DBL;                   // 32 KB (Dinv tiles)
static_assert(P_DBL == TRSM_BM * TRSM_BC, "region P doubles as the 128x128 K*/rhs tile");

struct TrsmParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  double* scratch;
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Copy a 128-row x 16-double tile (row stride `ld` doubles in global) into a swizzled smem tile.
__device__ __forceinline__ void load_tile(double* dst, const double* src, size_t ld, int tid) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int idx = tid + NT * u;
    int row = idx >> 3, c = idx & 7;
    cp_async16(dst + row * TRSM_BK + ((c ^ ((row & 1) << 2)) << 1), src + (size_t)row * ld + 2 * c);
  }
}

__global__ void __launch_bounds__(NT, 1) posterior_trsm_kernel(TrsmParams p) {
  extern __shared__ __align__(128) double smem[];
  const int d = p.d, n_pad = p.n_pad;
  double* P = smem;
  double* Q = P + P_DBL;
  double* xc_s = Q + Q_DBL;                 // [d][128]   candidates / lengthscale, transposed
  double* xt_s = xc_s + TRSM_BC * d;        // [128][d]   training rows of the current block / lengthscale
  double* alpha_s = xt_s + TRSM_BM * d;     // [128]
  double* red = alpha_s + TRSM_BM;          // [4][128]   mu halves, ss halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int Rw = wm * 64, Cw = wn * 32;
  const int swz = (g & 1) << 2;
  const int nblk = n_pad / TRSM_BM;
  const int kstart = p.pad / TRSM_BK;

  double* Vt = p.scratch + (size_t)blockIdx.x * TRSM_BC * n_pad;   // [128 cands][n_pad] solved panel

  for (int item = blockIdx.x; item < p.m * p.ntiles; item += gridDim.x) {
    const int j = item / p.ntiles, tile = item - j * p.ntiles;
    const int c0 = tile * TRSM_BC;
    const GpDev gp = p.gps[j * p.H + p.hsel];
    const int kind = gp.kind;
    const double variance = gp.variance;

    __syncthreads();
    for (int idx = tid; idx < TRSM_BC * d; idx += NT) {
      int c = idx / d, q = idx - c * d;
      int cc = min(c0 + c, p.N - 1);
      xc_s[q * TRSM_BC + c] = p.Xc[(size_t)cc * d + q] / gp.ell[q];
    }

    double mu_acc = 0.0;
    double ss[4][2];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) ss[jj][0] = ss[jj][1] = 0.0;
    double acc[8][4][2];

    for (int ib = 0; ib < nblk; ++ib) {
      const int R0 = ib * TRSM_BM;
      for (int idx = tid; idx < TRSM_BM * d; idx += NT) xt_s[idx] = gp.Xs[(size_t)R0 * d + idx];
      if (tid < TRSM_BM) alpha_s[tid] = gp.alpha[R0 + tid];
      __threadfence_block();
      __syncthreads();  // S1: staging visible; previous block's V stores ordered before the loads below; P is free

      // ---- phase a: K* tile (128 rows x 128 cands) -> P as [row][cand] (swizzled), mean partial sums
      {
        const int c = tid & 127, half = tid >> 7;
        for (int it4 = 0; it4 < 16; ++it4) {
          double r2[4] = {0.0, 0.0, 0.0, 0.0};
          const int rbase = half + 8 * it4;   // rows rbase + 2u
          for (int q = 0; q < d; ++q) {
            double xq = xc_s[q * TRSM_BC + c];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              double df = xt_s[(rbase + 2 * u) * d + q] - xq;
              r2[u] = fma(df, df, r2[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int row = rbase + 2 * u;
            double kv = (R0 + row >= p.pad) ? kern_value(kind, variance, r2[u]) : 0.0;
            mu_acc = fma(kv, alpha_s[row], mu_acc);
            P[row * TRSM_BC + ((((c >> 1) ^ ((row & 1) << 2))) << 1) + (c & 1)] = kv;
          }
        }
      }
      __syncthreads();  // S2
      // ---- phase b: accumulators <- K* in the DMMA C-fragment layout
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          double2 v = *reinterpret_cast<const double2*>(P + (Rw + 8 * i + g) * TRSM_BC + ((((Cw >> 1) + 4 * jj + t) ^ swz) << 1));
          acc[i][jj][0] = v.x;
          acc[i][jj][1] = v.y;
        }
      __syncthreads();  // S3: P free for the pipeline

      // ---- phase c: acc += (-L[ib, k]) * V[k]  over the already-solved rows
      const double* Dblk = gp.Dinv + (size_t)ib * TRSM_BM * TRSM_BM;
      load_tile(Q, Dblk, TRSM_BM, tid);
      load_tile(Q + TILE_DBL, Dblk + TRSM_BK, TRSM_BM, tid);
      const int kt0 = kstart, nsteps = (ib == 0) ? 0 : ib * (TRSM_BM / TRSM_BK) - kt0;
      const double* Arow = gp.Lneg + (size_t)R0 * n_pad;
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nsteps) {
          load_tile(P + s * 2 * TILE_DBL, Arow + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
          load_tile(P + s * 2 * TILE_DBL + TILE_DBL, Vt + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
        }
        cp_async_commit();
      }
      for (int it = 0; it < nsteps; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
          int nxt = it + STAGES - 1;
          if (nxt < nsteps) {
            double* st = P + (nxt % STAGES) * 2 * TILE_DBL;
            load_tile(st, Arow + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
            load_tile(st + TILE_DBL, Vt + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
          }
          cp_async_commit();
        }
        const double* As = P + (it % STAGES) * 2 * TILE_DBL;
        const double* Bs = As + TILE_DBL;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int ch = ((t + 4 * kk) ^ swz) << 1;
          double2 a[8], b[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double2*>(As + (Rw + 8 * i + g) * TRSM_BK + ch);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) b[jj] = *reinterpret_cast<const double2*>(Bs + (Cw + 8 * jj + g) * TRSM_BK + ch);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a[i].x, b[jj].x);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a[i].y, b[jj].y);
        }
      }
      cp_async_wait<0>();
      __syncthreads();  // S4: Dinv tiles 0,1 landed; every warp is done with the pipeline stages

      // ---- phase d: rhs -> P as the B operand [cand][k] (swizzled)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            int kr = Rw + 8 * i + g, cand = Cw + 8 * jj + 2 * t + e;
            P[cand * TRSM_BM + (((kr >> 1) ^ ((cand & 1) << 2)) << 1) + (kr & 1)] = acc[i][jj][e];
            acc[i][jj][e] = 0.0;
          }
      __syncthreads();  // S5

      // ---- phase e: V = Dinv * rhs (lower-triangular 128x128 times 128x128), 8 k-steps, 2-stage Dinv pipeline
      for (int ks = 0; ks < TRSM_BM / TRSM_BK; ++ks) {
        if (ks >= 2) {
          cp_async_wait<1>();
          __syncthreads();
        }
        const double* As = Q + (ks & 1) * TILE_DBL;
        if (TRSM_BK * ks <= Rw + 63) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int ch = ((t + 4 * kk) ^ swz) << 1;
            const int chb = ((8 * ks + t + 4 * kk) ^ swz) << 1;
            double2 b[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) b[jj] = *reinterpret_cast<const double2*>(P + (Cw + 8 * jj + g) * TRSM_BM + chb);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (TRSM_BK * ks <= Rw + 8 * i + 7) {
                double2 a = *reinterpret_cast<const double2*>(As + (Rw + 8 * i + g) * TRSM_BK + ch);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a.x, b[jj].x);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a.y, b[jj].y);
              }
            }
          }
        }
        __syncthreads();
        if (ks + 2 < TRSM_BM / TRSM_BK) load_tile(Q + (ks & 1) * TILE_DBL, Dblk + (size_t)(ks + 2) * TRSM_BK, TRSM_BM, tid);
        cp_async_commit();
      }
      cp_async_wait<0>();

      // ---- phase f: column square-sums; publish V for the following block rows
      const bool publish = (ib + 1 < nblk);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = acc[i][jj][e];
            ss[jj][e] = fma(v, v, ss[jj][e]);
            if (publish) Vt[(size_t)(Cw + 8 * jj + 2 * t + e) * n_pad + R0 + Rw + 8 * i + g] = v;
          }
    }

    // ---- finalize the item: reduce over the row-owning lanes / warps, write mean and variance
    red[(tid >> 7) * TRSM_BC + (tid & 127)] = mu_acc;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = ss[jj][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) red[(2 + wm) * TRSM_BC + Cw + 8 * jj + 2 * t + e] = v;
      }
    __syncthreads();
    if (tid < TRSM_BC && c0 + tid < p.N) {
      if (p.mean) p.mean[(size_t)j * p.N + c0 + tid] = (red[tid] + red[TRSM_BC + tid]) + gp.ymean;
      if (p.var) {
        double v = variance - (red[2 * TRSM_BC + tid] + red[3 * TRSM_BC + tid]);
        if (p.flags & BOPL_VAR_INCLUDE_NOISE) v = gp.noise + v;
        if (p.flags & BOPL_VAR_CLIP) v = fmax(v, 1e-10);
        p.var[(size_t)j * p.N + c0 + tid] = v;
      }
    }
  }
}

}  // namespace

int launch_posterior_trsm(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                          cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  TrsmParams p;
  p.gps = h->gp_dev;
  p.Xc = Xc;
  p.mean = mean;
  p.var = var;
  p.H = h->H;
  p.hsel = hsel;
  p.m = h->m;
  p.N = N;
  p.d = h->d;
  p.n_pad = h->n_pad;
  p.pad = h->pad;
  p.ntiles = (N + TRSM_BC - 1) / TRSM_BC;
  p.flags = flags;
  long long items = (long long)p.ntiles * h->m;
  int grid = (int)std::min<long long>(items, h->sms);
  size_t need = (size_t)grid * TRSM_BC * h->n_pad * sizeof(double);
  int rc = ensure_bytes(h, (void**)&h->trsm_scratch, &h->trsm_scratch_bytes, need);
  if (rc) return rc;
  p.scratch = h->trsm_scratch;
  size_t smem = (size_t)(P_DBL + Q_DBL + 2 * TRSM_BC * h->d + TRSM_BM + 4 * TRSM_BC) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  posterior_trsm_kernel<<<grid, NT, smem, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "posterior_trsm_kernel");
  return BOPL_OK;
}

}  // namespace bopl
