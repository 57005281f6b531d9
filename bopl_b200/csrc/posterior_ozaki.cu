// posterior_ozaki.cu — kernels (1)+(2) of the hot path with the many-right-hand-side update on the INT8 tensor pipe
// (tcgen05.mma kind::i8, TMEM accumulators) by an error-free split ("Ozaki scheme") of the fp64 operands.  OPT-IN
// (BOPL_TRSM_VARIANT=4 + a model uploaded with bopl_set_model while BOPL_OZAKI=1); the DMMA kernel stays the default.
//
// Same mathematics as posterior_trsm_t.cu (posterior.py:299-320, gp.py:380-435, linalg.py:92-111): left-looking blocked forward
// substitution V_i = Dinv_i (K*_i - sum_{k<i} L_ik V_k), 64-row blocks, tile = 128 candidates, then mean / variance.
// What changes is the update sum_k L_ik V_k (n^2/2 FMAs per candidate, 97 % of the work).  sm_100a has no fp64 tcgen05 kind,
// but its int8 kind is 128x faster than the fp64 pipe, and an fp64 product can be split into exact int8 products:
//   * every row of the off-diagonal part of L is scaled by a power of two ls_r >= 2 max|L_rk| and written in S = 6 balanced
//     base-256 digits (int8): L_rk = ls_r * sum_s d_s 2^(-7-8s)  (47 fractional bits, done once on the host);
//   * |V| <= sqrt(variance) (because var = variance - |V|^2 >= 0), so V is written in fixed point against vs = 2^ceil(log2 sigma)+1
//     the same way, on the fly, by the warp that solved it;
//   * digit products are accumulated EXACTLY in s32 by the tensor core, one accumulator per digit weight g = s_L + s_V < 6
//     (|acc| <= 6 * n * 2^14 < 2^31 for n <= 16384), and combined once per block row in fp64 (Horner in 2^-8).
// Dropped terms (g >= 6) are below 2^-48 of ls_r * vs * n: measured 5e-12 relative on the variance at the benchmark shape, vs
// 3e-14 for the fp64 kernel (tests/test_gpu_parity.py::test_ozaki_*).
//
// One instruction covers several digit pairs: the six digit planes of an L chunk are stacked as ONE 384-row B operand, so
// V digit a against L digits 0..5-a is a single N = 64 (6 - a) MMA (split at N = 256) whose output columns are exactly the
// accumulators g = a..5 — 8 instructions per 32-deep k-step instead of 21.
//
// Roles (192 threads, one persistent CTA per SM): warp 0 lane 0 streams operand chunks with cp.async.bulk (L digit images from
// the model, V digit images from this CTA's scratch panel) through a 2-stage ring; warp 1 lane 0 issues the MMAs; warps 2-5
// (thread = candidate = TMEM lane) generate K*, read the accumulators, apply the inverted 64x64 diagonal block, accumulate
// mean and sum V^2, and publish the digits of the solved block.
#include "pipeline.cuh"

namespace bopl {

namespace {

constexpr int OB = 64;                       // rows per block row
constexpr int OS = 6;                        // digits
constexpr int OZ_MAXD = 12;                  // candidate coordinates kept in registers
constexpr int LIMG = OS * OB * OB;           // 24576 B: 384 rows x 64 k
constexpr int VIMG1 = 128 * OB;              // 8192 B: one digit plane of V, 128 candidates x 64 k
constexpr int VIMG = OS * VIMG1;             // 49152 B
constexpr int OSTAGE = LIMG + VIMG;          // 73728 B
constexpr int ONSTAGE = 2;
constexpr int ONT = 192;

struct OzParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  signed char* scratch;   // [grid][n_pad/64][VIMG]
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
};

__host__ __device__ inline size_t oz_smem_bytes(int d) {
  return (size_t)ONSTAGE * OSTAGE + (size_t)OB * OB * 8 + (size_t)OB * (d + 2) * 8 + 256;
}

__device__ __forceinline__ uint64_t umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t idesc_s8(int N) {   // M = 128, s8 x s8 -> s32, both K-major
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((unsigned long long)smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

__global__ void __launch_bounds__(ONT, 1) posterior_ozaki_kernel(OzParams p) {
  extern __shared__ __align__(128) unsigned char osm[];
  unsigned char* stage_s = osm;                                              // [2][L image | V image]
  double* Dinv_s = (double*)(osm + (size_t)ONSTAGE * OSTAGE);                // [64][64]
  double* Xt_s = Dinv_s + OB * OB;                                           // [64][d+2]: scaled training row, alpha, row scale
  __shared__ unsigned long long full_bar[ONSTAGE], empty_bar[ONSTAGE], acc_full, tmem_empty, v_pub;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.d, n_pad = p.n_pad, nb = n_pad / OB, XT = d + 2;
  signed char* Vpanel = p.scratch + (size_t)blockIdx.x * nb * VIMG;

  if (tid == 0) {
    for (int s = 0; s < ONSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&tmem_empty, 128);
    mbar_init(&v_pub, 128);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((unsigned long long)smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int nitems = p.m * p.ntiles;

  if (warp == 0) {
    // ================================================================ producer: operand chunks -> shared memory
    if (lane == 0) {
      unsigned gc = 0, brow = 0;          // chunks issued so far; block rows whose digits were waited for so far
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int j = item / p.ntiles;
        const signed char* Limg = p.gps[j * p.H + p.hsel].OzL;
        for (int i = 1; i < nb; ++i) {
          for (int kb = 0; kb < i; ++kb) {
            const unsigned st = gc % ONSTAGE, use = gc / ONSTAGE;
            mbar_wait(&empty_bar[st], (use & 1) ^ 1);
            if (kb == i - 1) {            // the newest block: wait until its digits are in the scratch panel
              mbar_wait(&v_pub, brow & 1);
              ++brow;
              asm volatile("fence.proxy.async;\n" ::: "memory");
            }
            unsigned char* dst = stage_s + (size_t)st * OSTAGE;
            mbar_expect_tx(&full_bar[st], OSTAGE);
            bulk_g2s(dst, Limg + ((size_t)i * (i - 1) / 2 + kb) * LIMG, LIMG, &full_bar[st]);
            bulk_g2s(dst + LIMG, Vpanel + (size_t)kb * VIMG, VIMG, &full_bar[st]);
            ++gc;
          }
        }
        mbar_wait(&v_pub, brow & 1);       // the last block row publishes too: keep the phases in step, and the panel is
        ++brow;                            // not overwritten by the next item before every read of it has been issued
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      unsigned gc = 0, uu = 0;            // chunks consumed; accumulator generations started
      const unsigned lboA = 128 * 16, lboB = OS * OB * 16, sbo = 128;
      const uint32_t id256 = idesc_s8(256), id192 = idesc_s8(192), id128 = idesc_s8(128), id64 = idesc_s8(64);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        for (int i = 1; i < nb; ++i) {
          if (uu > 0) mbar_wait(&tmem_empty, (uu - 1) & 1);     // the previous block row's accumulators have been read
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          for (int kb = 0; kb < i; ++kb) {
            const unsigned st = gc % ONSTAGE, use = gc / ONSTAGE;
            mbar_wait(&full_bar[st], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const unsigned lb = smem_u32(stage_s + (size_t)st * OSTAGE), vb = lb + LIMG;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
              const unsigned lk = lb + ks * 2 * lboB;
              // V digit a (A operand, 128 candidates) x L digits 0..5-a (B operand rows 64 s ..) -> accumulators a..5 (64 columns each)
              {
                const uint64_t a0 = umma_desc(vb + 0 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 0, a0, umma_desc(lk, lboB, sbo), id256, first);
                umma_i8(tmem + 256, a0, umma_desc(lk + 4 * 1024, lboB, sbo), id128, first);
                const uint64_t a1 = umma_desc(vb + 1 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 64, a1, umma_desc(lk, lboB, sbo), id256, 1u);
                umma_i8(tmem + 320, a1, umma_desc(lk + 4 * 1024, lboB, sbo), id64, 1u);
                const uint64_t a2 = umma_desc(vb + 2 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 128, a2, umma_desc(lk, lboB, sbo), id256, 1u);
                const uint64_t a3 = umma_desc(vb + 3 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 192, a3, umma_desc(lk, lboB, sbo), id192, 1u);
                const uint64_t a4 = umma_desc(vb + 4 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 256, a4, umma_desc(lk, lboB, sbo), id128, 1u);
                const uint64_t a5 = umma_desc(vb + 5 * VIMG1 + ks * 2 * lboA, lboA, sbo);
                umma_i8(tmem + 320, a5, umma_desc(lk, lboB, sbo), id64, 1u);
              }
            }
            umma_commit(&empty_bar[st]);
            ++gc;
          }
          umma_commit(&acc_full);
          ++uu;
        }
      }
    }
  } else {
    // ================================================================ compute warps: thread = candidate = TMEM lane
    const int cl = 32 * (warp & 3) + lane;                       // candidate within the tile; warp w owns TMEM lanes 32 (w % 4) ..
    const uint32_t trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    const int ctid = tid - 64;
    unsigned uu = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int j = item / p.ntiles, tile = item - j * p.ntiles;
      const GpDev& gp = p.gps[j * p.H + p.hsel];
      const int kind = gp.kind;
      const double variance = gp.variance;
      const double vs = exp2(ceil(log2(sqrt(variance))) + 1.0);  // fixed-point scale of V: |V| <= sigma <= vs / 2
      const double to_fix = 140737488355328.0 / vs;              // 2^47 / vs
      const double from_fix = vs * (1.0 / 16384.0);              // vs 2^-14
      const int cc = min(tile * 128 + cl, p.N - 1);
      double xs[OZ_MAXD];
#pragma unroll
      for (int q = 0; q < OZ_MAXD; ++q) xs[q] = (q < d) ? p.Xc[(size_t)cc * d + q] / gp.ell[q] : 0.0;
      double mu = 0.0, ss = 0.0;
      for (int i = 0; i < nb; ++i) {
        // ---- stage the inverted diagonal block, the scaled training rows, alpha and the row scales of this block row
        compute_bar();
        {
          const double* src = gp.OzDinv + (size_t)i * OB * OB;
          for (int idx = ctid; idx < OB * OB / 2; idx += 128) cp_async16(Dinv_s + 2 * idx, src + 2 * idx);
          for (int idx = ctid; idx < OB * XT; idx += 128) {
            const int r = idx / XT, q = idx - r * XT, R = i * OB + r;
            Xt_s[idx] = q < d ? gp.Xs[(size_t)R * d + q] : (q == d ? gp.alpha[R] : gp.OzScale[R]);
          }
          cp_async_commit();
          cp_async_wait<0>();
        }
        compute_bar();
        // ---- K* of the block row (rows in the front padding are masked), mean
        double rhs[OB];
#pragma unroll
        for (int r8 = 0; r8 < OB / 8; ++r8) {
          double r2[8], kv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double* xt = Xt_s + (r8 * 8 + u) * XT;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < OZ_MAXD; ++q)
              if (q < d) {
                const double df = xt[q] - xs[q];
                s = fma(df, df, s);
              }
            r2[u] = s;
          }
          kern_value8(kind, variance, r2, kv);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r8 * 8 + u;
            const double k = (i * OB + r >= p.pad) ? kv[u] : 0.0;
            rhs[r] = k;
            mu = fma(k, Xt_s[r * XT + d], mu);
          }
        }
        // ---- subtract sum_k L_ik V_k: six s32 accumulators per (candidate, row), Horner in 2^-8, then the two power-of-two scales
        if (i > 0) {
          mbar_wait(&acc_full, uu & 1);
          ++uu;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
          for (int c8 = 0; c8 < OB / 8; ++c8) {
            int a[OS][8];
#pragma unroll
            for (int g = 0; g < OS; ++g) tmem_ld8(trow + (uint32_t)(g * 64 + c8 * 8), a[g]);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              double t = (double)a[OS - 1][u];
#pragma unroll
              for (int g = OS - 2; g >= 0; --g) t = fma(t, 1.0 / 256.0, (double)a[g][u]);
              const int r = c8 * 8 + u;
              rhs[r] = fma(-(Xt_s[r * XT + d + 1] * from_fix), t, rhs[r]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          mbar_arrive(&tmem_empty);
        }
        // ---- V_i = Dinv_i rhs, in place from the last row up (row r only needs rhs[0..r])
#pragma unroll
        for (int r = OB - 1; r >= 0; --r) {
          const double* drow = Dinv_s + r * OB;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int k = 0; k + 1 <= r; k += 2) {
            const double2 dd = *(const double2*)(drow + k);
            s0 = fma(dd.x, rhs[k], s0);
            s1 = fma(dd.y, rhs[k + 1], s1);
          }
          if ((r & 1) == 0) s0 = fma(drow[r], rhs[r], s0);
          rhs[r] = s0 + s1;
        }
#pragma unroll
        for (int r = 0; r < OB; ++r) ss = fma(rhs[r], rhs[r], ss);
        // ---- publish the digits of V_i for the later block rows
        if (i + 1 < nb) {
          signed char* img = Vpanel + (size_t)i * VIMG;
#pragma unroll
          for (int kc = 0; kc < OB / 16; ++kc) {
            uint32_t w[OS][4];
#pragma unroll
            for (int s = 0; s < OS; ++s)
#pragma unroll
              for (int q = 0; q < 4; ++q) w[s][q] = 0u;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              long long t = __double2ll_rn(rhs[kc * 16 + u] * to_fix);
              t = max(min(t, (1LL << 46)), -(1LL << 46));
#pragma unroll
              for (int s = OS - 1; s >= 0; --s) {
                const int dg = (int)((t + 128) & 255) - 128;
                t = (t - dg) >> 8;
                w[s][u >> 2] |= (uint32_t)(dg & 255) << (8 * (u & 3));
              }
            }
            const size_t off = ((size_t)kc * 16 + (cl >> 3)) * 128 + (size_t)(cl & 7) * 16;
#pragma unroll
            for (int s = 0; s < OS; ++s) *(uint4*)(img + (size_t)s * VIMG1 + off) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
          }
          __threadfence();
          asm volatile("fence.proxy.async;\n" ::: "memory");
        }
        mbar_arrive(&v_pub);
      }
      if (tile * 128 + cl < p.N) {
        const size_t o = (size_t)j * p.N + tile * 128 + cl;
        if (p.mean) p.mean[o] = mu + gp.ymean;
        if (p.var) {
          double v = variance - ss;
          if (p.flags & BOPL_VAR_INCLUDE_NOISE) v = gp.noise + v;
          if (p.flags & BOPL_VAR_CLIP) v = fmax(v, 1e-10);
          p.var[o] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace

bool ozaki_available(const bopl_handle* h) {
  if (!h->has_oz || h->d > OZ_MAXD || h->n_pad > 16384) return false;
  return true;
}

// Host side of the split: digit images of the off-diagonal blocks of the (front-padded, un-negated) factor Lp [n_pad x n_pad],
// the per-row power-of-two scales, and the inverted 64 x 64 diagonal blocks.  img: nb (nb-1)/2 chunks of LIMG bytes.
void ozaki_prepare(const double* Lp, int n_pad, std::vector<signed char>& img, std::vector<double>& scale) {
  const int nb = n_pad / OB;
  img.assign((size_t)nb * (nb - 1) / 2 * LIMG, 0);
  scale.assign((size_t)n_pad, 1.0);
  for (int R = 0; R < n_pad; ++R) {
    const int i = R / OB;
    double mx = 0.0;
    for (int k = 0; k < i * OB; ++k) mx = std::max(mx, std::fabs(Lp[(size_t)R * n_pad + k]));
    if (mx > 0.0) {
      int e;
      std::frexp(mx, &e);                  // mx = f 2^e, 0.5 <= f < 1  ->  2^e > mx
      scale[R] = std::ldexp(1.0, e + 1);   // >= 2 mx: digits of x / scale stay within +-2^46 / 2^47
    }
  }
  for (int i = 1; i < nb; ++i)
    for (int kb = 0; kb < i; ++kb) {
      signed char* base = img.data() + ((size_t)i * (i - 1) / 2 + kb) * LIMG;
      for (int r = 0; r < OB; ++r) {
        const int R = i * OB + r;
        const double inv = 1.0 / scale[R];
        for (int k = 0; k < OB; ++k) {
          long long t = std::llrint(std::ldexp(Lp[(size_t)R * n_pad + kb * OB + k] * inv, 47));
          for (int s = OS - 1; s >= 0; --s) {
            const int dg = (int)((t + 128) & 255) - 128;
            t = (t - dg) >> 8;
            const int rr = s * OB + r;
            base[((size_t)(k / 16) * (OS * OB / 8) + rr / 8) * 128 + (rr % 8) * 16 + (k % 16)] = (signed char)dg;
          }
        }
      }
    }
}

int launch_posterior_ozaki(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags, cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  OzParams p;
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
  p.ntiles = (N + 127) / 128;
  p.flags = flags;
  const long long items = (long long)p.ntiles * h->m;
  const int grid = (int)std::min<long long>(items, h->sms);
  const int nb = h->n_pad / OB;
  int rc = ensure_bytes(h, (void**)&h->oz_scratch, &h->oz_scratch_bytes, (size_t)grid * nb * VIMG);
  if (rc) return rc;
  p.scratch = (signed char*)h->oz_scratch;
  const size_t smem = oz_smem_bytes(h->d);
  if (smem > h->oz_attr_smem) {
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->oz_attr_smem = smem;
  }
  posterior_ozaki_kernel<<<grid, ONT, smem, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "posterior_ozaki_kernel");
  return BOPL_OK;
}

}  // namespace bopl
