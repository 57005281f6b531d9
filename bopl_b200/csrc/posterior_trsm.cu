// posterior_trsm.cu — kernels (1)+(2) of the hot path, fused: per attribute, for a tile of 128 candidates,
//   K* = K(X, Xc)                      (generated on the fly, never written to HBM)
//   mu = K*^T alpha + ymean             (GPy posterior.py:299-305, gp.py:398-399)
//   V  = L^{-1} K*                      (LAPACK dtrtrs in the reference: posterior.py:311, linalg.py:92-111)
//   var = variance - colsum(V^2) (+noise, clip)   (posterior.py:312, gp.py:416-417, gpmodel.py:214)
//
// The triangular solve is a LEFT-LOOKING BLOCKED forward substitution over 128-row blocks of L:
//   V_ib = Dinv_ib * ( K*_ib + sum_{k<ib} (-L_ib,k) V_k ),   Dinv_ib = L_ib,ib^{-1} (precomputed per model update)
// so that all flops are GEMM-shaped and run on the FP64 tensor pipe (DMMA.8x8x4 via mma.sync.m8n8k4.f64 — the
// only FP64 MMA shape sm_100a SASS has; tcgen05 has no f64 kind).  CTA tile: 128 rows x BC candidates, warps arranged
// 2 x (BC/32), warp tile 64 x 32 = 32 DMMA accumulators.  L tiles and the already-solved V panel (candidate-major,
// per-CTA scratch in L2/HBM) stream through a cp.async pipeline of 128x16 / BCx16 tiles with a 16-byte XOR swizzle
// that makes every fragment LDS.128 conflict-free.  Persistent CTAs walk (attribute, candidate-tile) items,
// attribute-major so that all CTAs stream the same L out of L2.
//
// Two configurations (same code, template parameters):
//   <BC=64,  4 warps, 3 stages, Dinv k-step 8>  ~100 KB smem, <=255 regs x 128 threads  -> TWO CTAs per SM.  The two
//       resident CTAs are at different phases most of the time, so one CTA's non-tensor phases (K* generation with its
//       long exp/sqrt dependency chains, operand re-layout, barriers) overlap the other's DMMA main loop.  Round-1 ncu
//       of the one-CTA form showed DMMA active 74 % with ~14 % of the time in K* generation and ~10 % in the
//       other non-MMA phases (profiles/r01_trsm_v1_ncu.md) — this is the fix.
//   <BC=128, 8 warps, 4 stages, Dinv k-step 16> ~185 KB smem, one CTA per SM (the round-1 first version; kept for A/B).
#include "pipeline.cuh"

namespace bopl {

namespace {

template <int BC_, int STAGES_, int DBK_>
struct TrsmCfg {
  static constexpr int BC = BC_;                              // candidates per tile
  static constexpr int WN = BC_ / 32;                         // warps along the candidate axis
  static constexpr int NT = 2 * WN * 32;                      // threads
  static constexpr int STAGES = STAGES_;                      // main-loop pipeline depth
  static constexpr int DBK = DBK_;                            // k-step of the Dinv tiles (8 or 16)
  static constexpr int A_DBL = TRSM_BM * TRSM_BK;             // L tile   128 x 16 doubles
  static constexpr int B_DBL = BC_ * TRSM_BK;                 // V tile   BC x 16 doubles
  static constexpr int STAGE_DBL = A_DBL + B_DBL;
  static constexpr int TILE_DBL = TRSM_BM * BC_;              // K* / rhs tile
  static constexpr int P_DBL = (STAGES_ * STAGE_DBL > TILE_DBL) ? STAGES_ * STAGE_DBL : TILE_DBL;
  static constexpr int D_DBL = TRSM_BM * DBK_;                // one Dinv tile
  static constexpr int MINB = (BC_ == 64) ? 2 : 1;
  // region Q holds the two Dinv tiles during phases c-e and the staged training rows + alpha during phase a
  static __host__ __device__ int q_dbl(int d) { int a = 2 * D_DBL, b = TRSM_BM * (d | 1) + TRSM_BM; return a > b ? a : b; }
  static __host__ __device__ size_t smem_bytes(int d) { return (size_t)(P_DBL + q_dbl(d) + BC_ * d + 4 * BC_) * sizeof(double); }
};

struct TrsmParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  double* scratch;
  int* sm_slots;             // [#SMs] zeroed before the launch: arrival order of the CTAs on each SM (stagger)
  long long stagger_cycles;  // delay of the second CTA of every SM before its first item (0 = none)
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) posterior_trsm_kernel(TrsmParams p) {
  constexpr int BC = Cfg::BC, NT = Cfg::NT, STAGES = Cfg::STAGES, DBK = Cfg::DBK, WN = Cfg::WN;
  constexpr int A_DBL = Cfg::A_DBL, STAGE_DBL = Cfg::STAGE_DBL, D_DBL = Cfg::D_DBL;
  extern __shared__ __align__(128) double smem[];
  const int d = p.d, n_pad = p.n_pad;
  double* P = smem;                              // pipeline stages | K* tile [row][cand] | rhs tile [cand][row]
  double* Q = P + Cfg::P_DBL;                    // Dinv tiles (phases c-e) | staged training rows + alpha (phase a)
  const int dp = d | 1;                          // odd row stride: the 8 row-lanes of a fragment hit 8 different bank pairs
  double* xt_s = Q;                              // [128][dp]  training rows of the current block / lengthscale
  double* alpha_s = Q + TRSM_BM * dp;            // [128]
  double* xc_s = Q + Cfg::q_dbl(d);              // [d][BC]    candidates / lengthscale, transposed
  double* red = xc_s + BC * d;                   // [4][BC]    mu halves, ss halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int Rw = wm * 64, Cw = wn * 32;
  const int swz = (g & 1) << 2;
  const int dswz = (DBK == 16) ? swz : 0;
  const int nblk = n_pad / TRSM_BM;
  const int kstart = p.pad / TRSM_BK;

  double* Vt = p.scratch + (size_t)blockIdx.x * BC * n_pad;   // [BC cands][n_pad] solved panel of this CTA

  __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES];
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NT);          // one (noinc) arrival per thread when its copies of the tile complete
      mbar_init(&empty_bar[s], NT / 32);    // one arrival per warp when it is done with the tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned g0 = 0;                          // running index of the pipeline tiles issued so far (stage = g % STAGES)

  // Two CTAs that start an SM together would stay in lock-step (every item costs the same), i.e. run their
  // non-tensor phases at the same time.  Delay the second arrival once, by a fraction of an item: the offset persists
  // for the whole launch and the phases interleave.
  if (Cfg::MINB == 2 && p.stagger_cycles > 0) {
    if (tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      if (atomicAdd(p.sm_slots + smid, 1) & 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.stagger_cycles) __nanosleep(2000);
      }
    }
    __syncthreads();
  }

  for (int item = blockIdx.x; item < p.m * p.ntiles; item += gridDim.x) {
    const int j = item / p.ntiles, tile = item - j * p.ntiles;
    const int c0 = tile * BC;
    const GpDev gp = p.gps[j * p.H + p.hsel];
    const int kind = gp.kind;
    const double variance = gp.variance;

    __syncthreads();
    for (int idx = tid; idx < BC * d; idx += NT) {
      int c = idx / d, q = idx - c * d;
      int cc = min(c0 + c, p.N - 1);
      xc_s[q * BC + c] = p.Xc[(size_t)cc * d + q] / gp.ell[q];
    }

    double mu[4][2], ss[4][2];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) mu[jj][0] = mu[jj][1] = ss[jj][0] = ss[jj][1] = 0.0;
    double acc[8][4][2];

    for (int ib = 0; ib < nblk; ++ib) {
      const int R0 = ib * TRSM_BM;
      for (int idx = tid; idx < TRSM_BM * d; idx += NT) {
        int r = idx / d, q = idx - r * d;
        xt_s[r * dp + q] = gp.Xs[(size_t)R0 * d + idx];
      }
      for (int idx = tid; idx < TRSM_BM; idx += NT) alpha_s[idx] = gp.alpha[R0 + idx];
      __threadfence_block();
      __syncthreads();  // S1: staging visible; previous block's V stores ordered before the loads below; P and Q are free

      // ---- pipeline prologue first (region P is free during phase a): the first L / V tiles of this block row travel
      //      from L2 while K* is being generated
      const int kt0 = kstart, nsteps = (ib == 0) ? 0 : ib * (TRSM_BM / TRSM_BK) - kt0;
      const double* Arow = gp.Lneg + (size_t)R0 * n_pad;
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nsteps) {   // every earlier tile has been consumed by all warps (S1) -> no empty wait
          const unsigned gs = (g0 + s) % STAGES;
          load_tile<TRSM_BM, TRSM_BK, NT>(P + gs * STAGE_DBL, Arow + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
          load_tile<BC, TRSM_BK, NT>(P + gs * STAGE_DBL + A_DBL, Vt + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
          cp_async_mbar_arrive(&full_bar[gs]);
        }
      }

      // ---- phase a: K* for this thread's own accumulator fragments (row Rw+8i+g, candidate Cw+8jj+2t+e), straight
      //      into the DMMA C registers: squared distances first, then the kernel value in place; mean partial sums.
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
      for (int q = 0; q < d; ++q) {
        double xr[8];
        double2 xcq[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) xr[i] = xt_s[(Rw + 8 * i + g) * dp + q];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) xcq[jj] = *reinterpret_cast<const double2*>(xc_s + q * BC + Cw + 8 * jj + 2 * t);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            double d0 = xr[i] - xcq[jj].x, d1 = xr[i] - xcq[jj].y;
            acc[i][jj][0] = fma(d0, d0, acc[i][jj][0]);
            acc[i][jj][1] = fma(d1, d1, acc[i][jj][1]);
          }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = Rw + 8 * i + g;
        const bool live = (R0 + row >= p.pad);
        const double al = alpha_s[row];
        double kv[4][2];
        kern_value8(kind, variance, &acc[i][0][0], &kv[0][0]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = live ? kv[jj][e] : 0.0;
            mu[jj][e] = fma(v, al, mu[jj][e]);
            acc[i][jj][e] = v;
          }
      }
      __syncthreads();  // S3: every warp is done with the staged rows in Q -> Q is free for the Dinv tiles

      // ---- phase c: acc += (-L[ib, k]) * V[k]  over the already-solved rows
      const double* Dblk = gp.Dinv + (size_t)ib * TRSM_BM * TRSM_BM;
      load_tile<TRSM_BM, DBK, NT>(Q, Dblk, TRSM_BM, tid);
      load_tile<TRSM_BM, DBK, NT>(Q + D_DBL, Dblk + DBK, TRSM_BM, tid);
      cp_async_commit();
      for (int it = 0; it < nsteps; ++it) {
        const unsigned gt = g0 + it, st_c = gt % STAGES;
        mbar_wait(&full_bar[st_c], (gt / STAGES) & 1);
        const double* As = P + st_c * STAGE_DBL;
        const double* Bs = As + A_DBL;
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
          if (kk == 0) {
            // mid-iteration refill of the stage consumed one iteration ago: by now every warp has long released it
            const int nxt = it + STAGES - 1;
            if (nxt < nsteps) {
              const unsigned gn = g0 + nxt, st_n = gn % STAGES;
              if (gn >= (unsigned)STAGES && it > 0) mbar_wait(&empty_bar[st_n], ((gn / STAGES) - 1) & 1);
              double* stp = P + st_n * STAGE_DBL;
              load_tile<TRSM_BM, TRSM_BK, NT>(stp, Arow + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
              load_tile<BC, TRSM_BK, NT>(stp + A_DBL, Vt + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
              cp_async_mbar_arrive(&full_bar[st_n]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a[i].y, b[jj].y);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st_c]);   // after this warp's last DMMA on the tile has been issued
      }
      g0 += (unsigned)nsteps;
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

      // ---- phase e: V = Dinv * rhs (lower-triangular 128x128 times 128xBC), 2-stage Dinv pipeline of 128 x DBK tiles
      constexpr int NKS = TRSM_BM / DBK;
      for (int ks = 0; ks < NKS; ++ks) {
        if (ks >= 2) {
          cp_async_wait<1>();
          __syncthreads();
        }
        const double* As = Q + (ks & 1) * D_DBL;
        if (DBK * ks <= Rw + 63) {
#pragma unroll
          for (int kk = 0; kk < DBK / 8; ++kk) {
            const int ch = ((t + 4 * kk) ^ dswz) << 1;
            const int chb = (((DBK / 2) * ks + t + 4 * kk) ^ swz) << 1;
            double2 b[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) b[jj] = *reinterpret_cast<const double2*>(P + (Cw + 8 * jj + g) * TRSM_BM + chb);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (DBK * ks <= Rw + 8 * i + 7) {
                double2 a = *reinterpret_cast<const double2*>(As + (Rw + 8 * i + g) * DBK + ch);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a.x, b[jj].x);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma(acc[i][jj][0], acc[i][jj][1], a.y, b[jj].y);
              }
            }
          }
        }
        __syncthreads();
        if (ks + 2 < NKS) load_tile<TRSM_BM, DBK, NT>(Q + (ks & 1) * D_DBL, Dblk + (size_t)(ks + 2) * DBK, TRSM_BM, tid);
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
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = ss[jj][e], w = mu[jj][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        w += __shfl_xor_sync(0xffffffffu, w, 16);
        if (g == 0) {
          red[wm * BC + Cw + 8 * jj + 2 * t + e] = w;
          red[(2 + wm) * BC + Cw + 8 * jj + 2 * t + e] = v;
        }
      }
    __syncthreads();
    if (tid < BC && c0 + tid < p.N) {
      if (p.mean) p.mean[(size_t)j * p.N + c0 + tid] = (red[tid] + red[BC + tid]) + gp.ymean;
      if (p.var) {
        double v = variance - (red[2 * BC + tid] + red[3 * BC + tid]);
        if (p.flags & BOPL_VAR_INCLUDE_NOISE) v = gp.noise + v;
        if (p.flags & BOPL_VAR_CLIP) v = fmax(v, 1e-10);
        p.var[(size_t)j * p.N + c0 + tid] = v;
      }
    }
  }
}

template <class Cfg>
int launch_cfg(bopl_handle* h, TrsmParams& p, cudaStream_t st) {
  p.ntiles = (p.N + Cfg::BC - 1) / Cfg::BC;
  long long items = (long long)p.ntiles * h->m;
  int grid = (int)std::min<long long>(items, (long long)h->sms * Cfg::MINB);
  size_t need = (size_t)grid * Cfg::BC * h->n_pad * sizeof(double);
  int rc = ensure_bytes(h, (void**)&h->trsm_scratch, &h->trsm_scratch_bytes, need);
  if (rc) return rc;
  p.scratch = h->trsm_scratch;
  p.sm_slots = nullptr;
  p.stagger_cycles = 0;
  if (Cfg::MINB == 2 && items >= 4LL * grid && h->trsm_stagger) {
    if (!h->sm_slots) BOPL_CUDA(h, cudaMalloc((void**)&h->sm_slots, 1024 * sizeof(int)));
    BOPL_CUDA(h, cudaMemsetAsync(h->sm_slots, 0, 1024 * sizeof(int), st));
    p.sm_slots = h->sm_slots;
    // one item = BC * n_pad * (n_pad + 128) / 2 FMAs on half an SM's FP64 tensor rate (32 FMA/clk); offset ~0.4 items
    p.stagger_cycles = (long long)(0.4 * (double)Cfg::BC * h->n_pad * (h->n_pad + 128.0) / 2.0 / 32.0);
  }
  size_t smem = Cfg::smem_bytes(h->d);
  size_t& attr_smem = h->attr_smem[Cfg::MINB == 2 ? 2 : 3];   // function attributes are per device: remember per handle
  if (smem > attr_smem) {
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_trsm_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_trsm_kernel<Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_smem = smem;
  }
  posterior_trsm_kernel<Cfg><<<grid, Cfg::NT, smem, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "posterior_trsm_kernel");
  return BOPL_OK;
}

}  // namespace

int launch_posterior_trsm(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                          cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  if (use_small_path(h, N)) return launch_posterior_small(h, Xc, N, hsel, mean, var, flags, st);
  if (h->trsm_variant == 4 && ozaki_available(h)) return launch_posterior_ozaki(h, Xc, N, hsel, mean, var, flags, st);
  // default: the candidate-owning kernel; wider inputs (d > 20) do not fit its shared-memory plan -> row-owning form
  if (h->trsm_variant != 1 && h->trsm_variant != 2 && trsm_t_fits(h->d))
    return launch_posterior_trsm_t(h, Xc, N, hsel, mean, var, flags, st);
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
  p.flags = flags;
  if (h->trsm_variant == 2) return launch_cfg<TrsmCfg<64, 3, 8>>(h, p, st);   // two CTAs per SM (A/B only)
  return launch_cfg<TrsmCfg<128, 4, 16>>(h, p, st);
}

}  // namespace bopl
