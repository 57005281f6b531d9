// posterior_trsm_t.cu — kernels (1)+(2) of the hot path in fp64 on the FP64 tensor pipe: per attribute, for a tile of 128 candidates,
//   K* = K(X, Xc)                      (generated on the fly, never written to HBM)
//   mu = K*^T alpha + ymean             (GPy posterior.py:299-305, gp.py:398-399)
//   V  = L^{-1} K*                      (LAPACK dtrtrs in the reference: posterior.py:311, linalg.py:92-111)
//   var = variance - colsum(V^2) (+noise, clip)   (posterior.py:312, gp.py:416-417, gpmodel.py:214)
// as a LEFT-LOOKING BLOCKED forward substitution over 128-row blocks of L,
//   V_ib = Dinv_ib * ( K*_ib + sum_{k<ib} (-L_ib,k) V_k ),   Dinv_ib = L_ib,ib^{-1} (precomputed per model update),
// so that all flops are GEMM-shaped and run on DMMA.8x8x4 (mma.sync.m8n8k4.f64 — the only FP64 MMA shape sm_100a SASS has; tcgen05
// has no f64 kind).  Persistent CTAs walk (attribute, candidate-tile) items, attribute-major so that all CTAs stream the same L
// out of L2.  The product is formed TRANSPOSED, V^T[cand][row], and every warp owns 16 candidates x all 128 rows of the block row:
//
//   * main loop     acc[cand][row] += V^T[cand][k] * (-L)^T[k][row]     A = V tile [cand][16k], B = L tile [row][16k]
//   * diagonal      V^T = rhs^T * Dinv^T.  The C fragment of an 8x8 tile holds (cand g, k = 2t, 2t+1): taken as the
//                   k-sets {0,2,4,6} and {1,3,5,7} it IS the A fragment of two DMMAs — so the triangular multiply runs
//                   straight out of the accumulator registers, in place, from the last 8-row tile down to the first.
//                   No shared-memory transposition, no CTA barrier between the main loop and the diagonal solve, and
//                   no load imbalance (in the row-owning form the warps holding rows 0-63 idle through half of it).
//   * Dinv          arrives as a host-packed stream of B fragments (136 lower-triangular 8x8 blocks x 32 lanes x 16 B =
//                   69.6 KB per block row), prefetched with cp.async during the main loop into its own region.
//   * reductions    mean and sum V^2 are per candidate, i.e. per (warp, lane-group): two xor-shuffles, no shared memory.
//
// Shared memory: STAGES x (128x16 L tile + 128x16 V tile) + Dinv fragments + staged inputs; one persistent CTA per SM.
#include "pipeline.cuh"

namespace bopl {

namespace {

constexpr int NT = 256;
constexpr int TILE = TRSM_BM * TRSM_BK;                       // 2048 doubles
constexpr int STAGE_DBL = 2 * TILE;                           // L tile | V tile
constexpr int DFRAG_DBL = 136 * 64;                           // packed Dinv fragments of one block row

struct TParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  double* scratch;
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
  int keep_tiles;   // V k-tiles (16 rows each) below this index are kept in L2 (evict_last); the rest is streamed
};

template <int STAGES>
__host__ __device__ constexpr size_t t_smem_bytes(int d) {
  return (size_t)(STAGES * STAGE_DBL + DFRAG_DBL + TRSM_BM * d + 2 * TRSM_BM * (d + 1)) * sizeof(double);
}

template <int STAGES, int HINTS>
__global__ void __launch_bounds__(NT, 1) posterior_trsm_t_kernel(TParams p) {
  extern __shared__ __align__(128) double smem[];
  const int d = p.d, n_pad = p.n_pad;
  double* P = smem;                                // pipeline stages
  double* DF = P + STAGES * STAGE_DBL;             // Dinv B fragments of the current block row
  double* xc_s = DF + DFRAG_DBL;                   // [d][128]  candidates / lengthscale
  double* SB = xc_s + TRSM_BM * d;                 // [2][(d+1)*128]  staged block: training rows^T / lengthscale, alpha
  const int SB_DBL = TRSM_BM * (d + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Cw = warp * 16;                        // this warp's 16 candidates
  const int swz = (g & 1) << 2;
  const int nblk = n_pad / TRSM_BM;
  const int kstart = p.pad / TRSM_BK;

  double* Vt = p.scratch + (size_t)blockIdx.x * TRSM_BM * n_pad;   // [128 cands][n_pad] solved panel of this CTA

  __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES], df_bar, sb_bar[2];
  if (tid == 0) {
    mbar_init(&sb_bar[0], NT);
    mbar_init(&sb_bar[1], NT);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NT);
      mbar_init(&empty_bar[s], NT / 32);
    }
    mbar_init(&df_bar, NT);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  unsigned g0 = 0;       // running index of the pipeline tiles issued so far
  unsigned dfc = 0;      // running count of Dinv-fragment loads
  unsigned bc = 0;       // running count of staged blocks (buffer = bc & 1)

  for (int item = blockIdx.x; item < p.m * p.ntiles; item += gridDim.x) {
    const int j = item / p.ntiles, tile = item - j * p.ntiles;
    const int c0 = tile * TRSM_BM;
    const GpDev gp = p.gps[j * p.H + p.hsel];
    const int kind = gp.kind;
    const double variance = gp.variance;

    __syncthreads();
    for (int idx = tid; idx < TRSM_BM * d; idx += NT) {
      int c = idx / d, q = idx - c * d;
      int cc = min(c0 + c, p.N - 1);
      xc_s[q * TRSM_BM + c] = p.Xc[(size_t)cc * d + q] / gp.ell[q];
    }

    // staged block 0 of this item (the buffer's last readers are two block rows back, behind the barrier above)
    for (int idx = tid; idx < SB_DBL / 2; idx += NT) cp_async16(SB + (bc & 1) * SB_DBL + 2 * idx, gp.StageBlk + 2 * idx);
    cp_async_mbar_arrive(&sb_bar[bc & 1]);

    double mu[2] = {0.0, 0.0}, ss[2] = {0.0, 0.0};
    double acc[2][16][2];

    for (int ib = 0; ib < nblk; ++ib) {
      const int R0 = ib * TRSM_BM;
      const double* xt_s = SB + (bc & 1) * SB_DBL;          // [d][128] training rows of the block / lengthscale
      const double* alpha_s = xt_s + TRSM_BM * d;           // [128]
      mbar_wait(&sb_bar[bc & 1], (bc >> 1) & 1);            // prefetched during the previous block row
      __threadfence_block();
      __syncthreads();  // S1: the previous block row's V stores are ordered before the loads below; every warp is past
                        //     the previous diagonal solve and K* generation -> DF, all stages and the other SB are free
      if (ib + 1 < nblk) {                                  // next block's rows travel while this block row computes
        const double* src = gp.StageBlk + (size_t)(ib + 1) * SB_DBL;
        double* dst = SB + ((bc + 1) & 1) * SB_DBL;
        for (int idx = tid; idx < SB_DBL / 2; idx += NT) cp_async16(dst + 2 * idx, src + 2 * idx);
        cp_async_mbar_arrive(&sb_bar[(bc + 1) & 1]);
      }

      // ---- prefetch: Dinv fragments of this block row and the first pipeline tiles
      {
        const double* src = gp.DinvFrag + (size_t)ib * DFRAG_DBL;
        for (int idx = tid; idx < DFRAG_DBL / 2; idx += NT) cp_async16(DF + 2 * idx, src + 2 * idx);
        cp_async_mbar_arrive(&df_bar);
      }
      const int kt0 = kstart, nsteps = (ib == 0) ? 0 : ib * (TRSM_BM / TRSM_BK) - kt0;
      const double* Lrow = gp.Lneg + (size_t)R0 * n_pad;
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nsteps) {
          const unsigned gs = (g0 + s) % STAGES;
          if (HINTS) {
            load_tile_hint<TRSM_BM, TRSM_BK, NT>(P + gs * STAGE_DBL, Lrow + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid, pol_keep);
            load_tile_hint<TRSM_BM, TRSM_BK, NT>(P + gs * STAGE_DBL + TILE, Vt + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid,
                                                 (kt0 + s) < p.keep_tiles ? pol_keep : pol_stream);
          } else {
            load_tile<TRSM_BM, TRSM_BK, NT>(P + gs * STAGE_DBL, Lrow + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
            load_tile<TRSM_BM, TRSM_BK, NT>(P + gs * STAGE_DBL + TILE, Vt + (size_t)(kt0 + s) * TRSM_BK, n_pad, tid);
          }
          cp_async_mbar_arrive(&full_bar[gs]);
        }
      }

      // ---- phase a: K*^T for this thread's accumulator fragments: candidate Cw+8mi+g, row 8nt+2t+e
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
      for (int q = 0; q < d; ++q) {
        const double x0 = xc_s[q * TRSM_BM + Cw + g], x1 = xc_s[q * TRSM_BM + Cw + 8 + g];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          const double2 xr = *reinterpret_cast<const double2*>(xt_s + q * TRSM_BM + 8 * nt + 2 * t);
          double a0 = xr.x - x0, a1 = xr.y - x0, b0 = xr.x - x1, b1 = xr.y - x1;
          acc[0][nt][0] = fma(a0, a0, acc[0][nt][0]);
          acc[0][nt][1] = fma(a1, a1, acc[0][nt][1]);
          acc[1][nt][0] = fma(b0, b0, acc[1][nt][0]);
          acc[1][nt][1] = fma(b1, b1, acc[1][nt][1]);
        }
      }
#pragma unroll
      for (int ntq = 0; ntq < 4; ++ntq) {
        double al[4][2];
        bool live[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = 8 * (4 * ntq + u) + 2 * t;
          const double2 a2 = *reinterpret_cast<const double2*>(alpha_s + row);
          al[u][0] = a2.x;
          al[u][1] = a2.y;
          live[u][0] = (R0 + row >= p.pad);
          live[u][1] = (R0 + row + 1 >= p.pad);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          double kv[8];
          kern_value8(kind, variance, &acc[mi][4 * ntq][0], kv);
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double v = live[u][e] ? kv[2 * u + e] : 0.0;
              mu[mi] = fma(v, al[u][e], mu[mi]);
              acc[mi][4 * ntq + u][e] = v;
            }
        }
      }

      // A warp cannot leave a main loop of >= STAGES steps before every thread has issued its share of a refill, i.e.
      // before every warp is past phase a — that orders the next block row's staging writes after these reads.  Short
      // loops (block row 0, tiny n) have no such hand-off: barrier explicitly (the condition is CTA-uniform).
      if (nsteps < STAGES) __syncthreads();

      // ---- main loop: acc[cand][row] += V^T[cand][k] * (-L)^T[k][row] over the already-solved rows
      for (int it = 0; it < nsteps; ++it) {
        const unsigned gt = g0 + it, st_c = gt % STAGES;
        mbar_wait(&full_bar[st_c], (gt / STAGES) & 1);
        const double* Ls = P + st_c * STAGE_DBL;
        const double* Vs = Ls + TILE;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int ch = ((t + 4 * kk) ^ swz) << 1;
          double2 a[2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) a[mi] = *reinterpret_cast<const double2*>(Vs + (Cw + 8 * mi + g) * TRSM_BK + ch);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            double2 b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = *reinterpret_cast<const double2*>(Ls + (8 * (8 * hf + u) + g) * TRSM_BK + ch);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][8 * hf + u][0], acc[mi][8 * hf + u][1], a[mi].x, b[u].x);
            if (kk == 0 && hf == 1) {
              // mid-iteration refill of the stage consumed one iteration ago
              const int nxt = it + STAGES - 1;
              if (nxt < nsteps) {
                const unsigned gn = g0 + nxt, st_n = gn % STAGES;
                if (it > 0) mbar_wait(&empty_bar[st_n], ((gn / STAGES) - 1) & 1);
                double* stp = P + st_n * STAGE_DBL;
                if (HINTS) {
                  load_tile_hint<TRSM_BM, TRSM_BK, NT>(stp, Lrow + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid, pol_keep);
                  load_tile_hint<TRSM_BM, TRSM_BK, NT>(stp + TILE, Vt + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid,
                                                       (kt0 + nxt) < p.keep_tiles ? pol_keep : pol_stream);
                } else {
                  load_tile<TRSM_BM, TRSM_BK, NT>(stp, Lrow + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
                  load_tile<TRSM_BM, TRSM_BK, NT>(stp + TILE, Vt + (size_t)(kt0 + nxt) * TRSM_BK, n_pad, tid);
                }
                cp_async_mbar_arrive(&full_bar[st_n]);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) dmma(acc[mi][8 * hf + u][0], acc[mi][8 * hf + u][1], a[mi].y, b[u].y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st_c]);
      }
      g0 += (unsigned)nsteps;

      // ---- diagonal block, in registers and in place: V^T[:, nt] = sum_{kt <= nt} rhs^T[:, kt] * Dinv^T[kt, nt]
      const bool publish = (ib + 1 < nblk);
      mbar_wait(&df_bar, dfc & 1);
      ++dfc;
      {
        const double2* frag = reinterpret_cast<const double2*>(DF) + lane;
        int fi = 0;
#pragma unroll
        for (int G = 3; G >= 0; --G) {
          double tmp[4][2][2];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) tmp[u][mi][0] = tmp[u][mi][1] = 0.0;
#pragma unroll
          for (int kt = 0; kt <= 4 * G + 3; ++kt) {
            // fragments of the (up to four) blocks of this k-step first, then all the k-even DMMAs, then all the k-odd ones:
            // consecutive DMMAs never hit the same accumulator
            double2 b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (kt <= 4 * G + u) {
                b[u] = frag[32 * fi];
                ++fi;
              }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (kt <= 4 * G + u) {
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) dmma(tmp[u][mi][0], tmp[u][mi][1], acc[mi][kt][0], b[u].x);
              }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (kt <= 4 * G + u) {
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) dmma(tmp[u][mi][0], tmp[u][mi][1], acc[mi][kt][1], b[u].y);
              }
          }
          // write the finished tiles back in place; their sum of squares and their publication for the following block
          // rows (16-byte stores, 64 B per candidate) overlap the DMMAs of the next group
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              const double v0 = tmp[u][mi][0], v1 = tmp[u][mi][1];
              acc[mi][4 * G + u][0] = v0;
              acc[mi][4 * G + u][1] = v1;
              ss[mi] = fma(v0, v0, ss[mi]);
              ss[mi] = fma(v1, v1, ss[mi]);
              if (publish) {
                double* dst = Vt + (size_t)(Cw + 8 * mi + g) * n_pad + R0 + 8 * (4 * G + u) + 2 * t;
                if (HINTS == 2) st_global_v2_hint(dst, v0, v1, (R0 / TRSM_BK) < p.keep_tiles ? pol_keep : pol_stream);
                else *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
              }
            }
        }
      }

      ++bc;
    }

    // ---- finalize the item: the four t-lanes of a candidate hold its partial sums
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      double v = ss[mi], w = mu[mi];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      w += __shfl_xor_sync(0xffffffffu, w, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      w += __shfl_xor_sync(0xffffffffu, w, 2);
      const int c = c0 + Cw + 8 * mi + g;
      if (t == 0 && c < p.N) {
        if (p.mean) p.mean[(size_t)j * p.N + c] = w + gp.ymean;
        if (p.var) {
          double vv = variance - v;
          if (p.flags & BOPL_VAR_INCLUDE_NOISE) vv = gp.noise + vv;
          if (p.flags & BOPL_VAR_CLIP) vv = fmax(vv, 1e-10);
          p.var[(size_t)j * p.N + c] = vv;
        }
      }
    }
  }
}

template <int STAGES, int HINTS>
int launch_t(bopl_handle* h, TParams& p, cudaStream_t st) {
  p.ntiles = (p.N + TRSM_BM - 1) / TRSM_BM;
  long long items = (long long)p.ntiles * h->m;
  int grid = (int)std::min<long long>(items, (long long)h->sms);
  size_t need = (size_t)grid * TRSM_BM * h->n_pad * sizeof(double);
  int rc = ensure_bytes(h, (void**)&h->trsm_scratch, &h->trsm_scratch_bytes, need);
  if (rc) return rc;
  p.scratch = h->trsm_scratch;
  {
    // L2 budget: the shared factor of the running attribute stays resident; of what is left (a fraction of the L2 size
    // the device reports) every CTA keeps the first block rows of its solved panel — the ones re-read most often.
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device);
    const double budget = (h->trsm_l2_keep_frac * (double)l2 - (double)h->n_pad * h->n_pad * 8.0 * 0.55) / grid;
    long long blocks = (long long)(budget / ((double)TRSM_BM * TRSM_BM * 8.0));
    if (h->trsm_keep_blocks >= 0) blocks = h->trsm_keep_blocks;
    blocks = std::max<long long>(0, std::min<long long>(blocks, h->n_pad / TRSM_BM));
    p.keep_tiles = (int)blocks * (TRSM_BM / TRSM_BK);
  }
  size_t smem = t_smem_bytes<STAGES>(h->d);
  size_t& attr_smem = h->attr_smem_t[STAGES - 2][HINTS];   // function attributes are per device: remember per handle
  if (smem > attr_smem) {
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_trsm_t_kernel<STAGES, HINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  posterior_trsm_t_kernel<STAGES, HINTS><<<grid, NT, smem, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "posterior_trsm_t_kernel");
  return BOPL_OK;
}

}  // namespace

// Host-side packing of the Dinv B fragments, in the order the kernel consumes them (see the diagonal-block loop):
// for G = 3..0, for kt = 0..4G+3, for u = 0..3 with kt <= 4G+u: block (nt = 4G+u, kt), lane 4g+t -> Dinv[8nt+g][8kt+2t .. +1].
void pack_dinv_fragments(const double* Dinv_block /*128x128 row-major*/, double* out /*136*64*/) {
  size_t fi = 0;
  for (int G = 3; G >= 0; --G)
    for (int kt = 0; kt <= 4 * G + 3; ++kt)
      for (int u = 0; u < 4; ++u) {
        if (kt > 4 * G + u) continue;
        const int nt = 4 * G + u;
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          out[fi * 64 + 2 * lane + 0] = Dinv_block[(size_t)(8 * nt + g) * TRSM_BM + 8 * kt + 2 * t];
          out[fi * 64 + 2 * lane + 1] = Dinv_block[(size_t)(8 * nt + g) * TRSM_BM + 8 * kt + 2 * t + 1];
        }
        ++fi;
      }
}

// Host-side packing of the per-block staging arrays: [nblk][(d+1)*128] = scaled training rows transposed [q][128], alpha[128].
void pack_stage_blocks(const double* Xs /*n_pad*d*/, const double* alpha /*n_pad*/, int n_pad, int d, double* out) {
  const int nblk = n_pad / TRSM_BM;
  for (int b = 0; b < nblk; ++b) {
    double* o = out + (size_t)b * TRSM_BM * (d + 1);
    for (int r = 0; r < TRSM_BM; ++r) {
      for (int q = 0; q < d; ++q) o[(size_t)q * TRSM_BM + r] = Xs[(size_t)(b * TRSM_BM + r) * d + q];
      o[(size_t)d * TRSM_BM + r] = alpha[b * TRSM_BM + r];
    }
  }
}

int launch_posterior_trsm_t(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                            cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  TParams p;
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
  // The kernel keeps the staged inputs of two block rows in shared memory: three pipeline stages fit up to d = 20, two up to d = 24 (MAX_D).
  if (t_smem_bytes<3>(h->d) + 256 > 227 * 1024) return launch_t<2, 1>(h, p, st);
  if (h->trsm_stages == 4 && t_smem_bytes<4>(h->d) + 256 <= 227 * 1024) return launch_t<4, 0>(h, p, st);
  if (h->trsm_l2_hints == 2) return launch_t<3, 2>(h, p, st);
  if (h->trsm_l2_hints == 1) return launch_t<3, 1>(h, p, st);
  return launch_t<3, 0>(h, p, st);
}

// Posterior mean and variance of a batch: the kernel choice behind bopl_posterior / bopl_uei*.  N <= 32 with the inverse factor loaded:
// matrix-vector path (gp_kernels.cu); otherwise the int8-tensor-pipe kernel (posterior_ozaki.cu) when it is selected and covers the
// shape, else the fp64 DMMA kernel above.  All three are device kernels of this library; there is no other path.
int launch_posterior_trsm(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                          cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  if (use_small_path(h, N)) return launch_posterior_small(h, Xc, N, hsel, mean, var, flags, st);
  if (h->use_ozaki && ozaki_available(h)) return launch_posterior_ozaki(h, Xc, N, hsel, mean, var, flags, st);
  return launch_posterior_trsm_t(h, Xc, N, hsel, mean, var, flags, st);
}

}  // namespace bopl
