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
constexpr int OZ_MAXD = 12;                  // input dimensions the shared-memory plan is sized for
constexpr int LIMG = OS * OB * OB;           // 24576 B: 384 stacked rows x 64 k, one off-diagonal block of L
constexpr int LHALF = LIMG / 2;              // one k-step (32 k) of it
constexpr int VPL = 128 * 32;                // 4096 B: one digit plane of V, 128 candidates x 32 k
constexpr int VHALF = OS * VPL;              // 24576 B: the six planes of one k-step
constexpr int VIMG = 2 * VHALF;              // 49152 B per solved block
constexpr int OSTAGE = LHALF + VHALF;        // 36864 B per pipeline stage = one k-step of one (block row, block) pair
constexpr int ONSTAGE = 4;
constexpr int NFRAG = 36;                    // lower-triangular 8x8 blocks of a 64x64 diagonal block
constexpr int ONT = 320;                     // producer warp, MMA warp, eight compute warps
constexpr int NCW = 8;

struct OzParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  signed char* scratch;   // [grid][n_pad/64][VIMG]
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
  int keep_blocks;        // solved blocks below this index are kept in L2 (evict_last), the rest is streamed
};

// staged per block row: Dinv B fragments, scaled training rows transposed [q][64], alpha[64], row scales[64]
__host__ __device__ inline int oz_sb_doubles(int d) { return NFRAG * 64 + (d + 2) * OB; }
__host__ __device__ inline size_t oz_smem_bytes(int d) {
  return (size_t)ONSTAGE * OSTAGE + (size_t)2 * oz_sb_doubles(d) * 8 + (size_t)d * 128 * 8;
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
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
// 16 TMEM lanes x 8 columns: lane l/4 (+8), columns 2 (l%4), 2 (l%4) + 1 — the C-fragment layout of an 8x8 DMMA tile pair
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}

__global__ void __launch_bounds__(ONT, 1) posterior_ozaki_kernel(OzParams p) {
  extern __shared__ __align__(128) unsigned char osm[];
  const int d = p.d, n_pad = p.n_pad, nb = n_pad / OB;
  const int SBD = oz_sb_doubles(d);
  unsigned char* stage_s = osm;                                              // [4][L k-step | V k-step]
  double* SB = (double*)(osm + (size_t)ONSTAGE * OSTAGE);                    // [2][SBD]
  double* xc_s = SB + 2 * SBD;                                               // [d][128] candidates / lengthscale
  __shared__ unsigned long long full_bar[ONSTAGE], empty_bar[ONSTAGE], sb_full[2], sb_empty[2], acc_full, tmem_empty, v_pub;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  signed char* Vpanel = p.scratch + (size_t)blockIdx.x * nb * VIMG;

  if (tid == 0) {
    for (int s = 0; s < ONSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sb_full[s], 1);
      mbar_init(&sb_empty[s], NCW);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&tmem_empty, NCW);
    mbar_init(&v_pub, NCW * 32);
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
    // ================================================================ producer: operand k-steps and per-block-row staging
    if (lane == 0) {
      unsigned gc = 0, brow = 0, sc = 0;   // k-steps issued; block rows whose digits were waited for; staging blocks issued
      const unsigned sb_bytes = (unsigned)SBD * 8;
      const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
      auto stage_block = [&](const GpDev& g, int i) {
        const unsigned buf = sc & 1;
        if (sc >= 2) mbar_wait(&sb_empty[buf], ((sc >> 1) - 1) & 1);
        mbar_expect_tx(&sb_full[buf], sb_bytes);
        bulk_g2s(SB + (size_t)buf * SBD, g.OzStage + (size_t)i * SBD, sb_bytes, &sb_full[buf]);
        ++sc;
      };
      auto load_block = [&](const signed char* Limg, int i, int kb) {      // both k-steps of block (i, kb)
        const unsigned long long vpol = kb < p.keep_blocks ? pol_keep : pol_stream;
#pragma unroll 1
        for (int ks = 0; ks < 2; ++ks) {
          const unsigned st = gc % ONSTAGE, use = gc / ONSTAGE;
          mbar_wait(&empty_bar[st], (use & 1) ^ 1);
          unsigned char* dst = stage_s + (size_t)st * OSTAGE;
          mbar_expect_tx(&full_bar[st], OSTAGE);
          bulk_g2s_hint(dst, Limg + ((size_t)i * (i - 1) / 2 + kb) * LIMG + (size_t)ks * LHALF, LHALF, &full_bar[st], pol_keep);
          bulk_g2s_hint(dst + LHALF, Vpanel + (size_t)kb * VIMG + (size_t)ks * VHALF, VHALF, &full_bar[st], vpol);
          ++gc;
        }
      };
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const GpDev& gp = p.gps[(item / p.ntiles) * p.H + p.hsel];
        const signed char* Limg = gp.OzL;
        if (item == (int)blockIdx.x) stage_block(gp, 0);
        for (int i = 0; i < nb; ++i) {
          // The staging block of row i+1 reuses the buffer of row i-1, free once that row's diagonal solve is over — which is
          // before its digits are published: issue it behind the wait for those digits, so that the k-steps of the blocks
          // kb < i-1 (needing neither) are prefetched while the previous block row is still being solved.
          for (int kb = 0; kb + 1 < i; ++kb) load_block(Limg, i, kb);
          if (i > 0) {                        // the newest block: wait until its digits are in the scratch panel
            mbar_wait(&v_pub, brow & 1);
            ++brow;
            asm volatile("fence.proxy.async;\n" ::: "memory");
          }
          if (i + 1 < nb) stage_block(gp, i + 1);
          else if (item + (int)gridDim.x < nitems) stage_block(p.gps[((item + (int)gridDim.x) / p.ntiles) * p.H + p.hsel], 0);
          if (i > 0) load_block(Limg, i, i - 1);
        }
        mbar_wait(&v_pub, brow & 1);       // the last block row arrives too: keep the phases in step
        ++brow;
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      unsigned gc = 0, uu = 0;            // k-steps consumed; accumulator generations started
      const unsigned lboA = 128 * 16, lboB = OS * OB * 16, sbo = 128;
      const uint32_t id256 = idesc_s8(256), id192 = idesc_s8(192), id128 = idesc_s8(128), id64 = idesc_s8(64);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        for (int i = 1; i < nb; ++i) {
          if (uu > 0) mbar_wait(&tmem_empty, (uu - 1) & 1);     // the previous block row's accumulators have been read
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          for (int ks2 = 0; ks2 < 2 * i; ++ks2) {
            const unsigned st = gc % ONSTAGE, use = gc / ONSTAGE;
            mbar_wait(&full_bar[st], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const unsigned lk = smem_u32(stage_s + (size_t)st * OSTAGE), vb = lk + LHALF;
            const uint32_t first = (ks2 == 0) ? 0u : 1u;
            // V digit a (A operand, 128 candidates) x L digits 0..5-a (B operand rows 64 s ..) -> accumulators a..5 (64 columns each)
            const uint64_t bl = umma_desc(lk, lboB, sbo), bh = umma_desc(lk + 4 * 1024, lboB, sbo);
            const uint64_t a0 = umma_desc(vb + 0 * VPL, lboA, sbo);
            umma_i8(tmem + 0, a0, bl, id256, first);
            umma_i8(tmem + 256, a0, bh, id128, first);
            const uint64_t a1 = umma_desc(vb + 1 * VPL, lboA, sbo);
            umma_i8(tmem + 64, a1, bl, id256, 1u);
            umma_i8(tmem + 320, a1, bh, id64, 1u);
            umma_i8(tmem + 128, umma_desc(vb + 2 * VPL, lboA, sbo), bl, id256, 1u);
            umma_i8(tmem + 192, umma_desc(vb + 3 * VPL, lboA, sbo), bl, id192, 1u);
            umma_i8(tmem + 256, umma_desc(vb + 4 * VPL, lboA, sbo), bl, id128, 1u);
            umma_i8(tmem + 320, umma_desc(vb + 5 * VPL, lboA, sbo), bl, id64, 1u);
            umma_commit(&empty_bar[st]);
            ++gc;
          }
          umma_commit(&acc_full);
          ++uu;
        }
      }
    }
  } else {
    // ================================================================ compute warps: 16 candidates x all 64 rows of the block row,
    // in the DMMA C-fragment layout: candidate cbase + 8 mi + g, row 8 nt + 2 t + e
    const int g = lane >> 2, t = lane & 3;
    const int cbase = 32 * (warp & 3) + 16 * ((warp - 2) >> 2);   // TMEM lanes of this warp: quadrant warp % 4, half (warp-2)/4
    const uint32_t trow = tmem + ((uint32_t)cbase << 16);
    unsigned uu = 0, sbc = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int j = item / p.ntiles, tile = item - j * p.ntiles;
      const GpDev& gp = p.gps[j * p.H + p.hsel];
      const int kind = gp.kind;
      const double variance = gp.variance;
      const double vs = exp2(ceil(log2(sqrt(variance))) + 1.0);  // fixed-point scale of V: |V| <= sigma <= vs / 2
      const double to_fix = 140737488355328.0 / vs;              // 2^47 / vs
      const double from_fix16 = vs * (1.0 / 1073741824.0);       // vs 2^-14 (digit product weight) 2^-16 (integer Horner)
      const double vmax = 0.5 * vs;
      __syncwarp();
      for (int idx = lane; idx < 16 * d; idx += 32) {
        const int c = idx / d, q = idx - c * d;
        const int cc = min(tile * 128 + cbase + c, p.N - 1);
        xc_s[q * 128 + cbase + c] = p.Xc[(size_t)cc * d + q] / gp.ell[q];
      }
      __syncwarp();
      double mu[2] = {0.0, 0.0}, ss[2] = {0.0, 0.0};
      double acc[2][8][2];
      for (int i = 0; i < nb; ++i) {
        const int R0 = i * OB;
        const unsigned buf = sbc & 1;
        mbar_wait(&sb_full[buf], (sbc >> 1) & 1);
        const double* frag_s = SB + (size_t)buf * SBD;
        const double* xt_s = frag_s + NFRAG * 64;
        const double* alpha_s = xt_s + d * OB;
        const double* scale_s = alpha_s + OB;

        // ---- K*^T of the block row straight into the fragments (rows in the front padding are masked), mean
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) acc[mi][nt][0] = acc[mi][nt][1] = 0.0;
        for (int q = 0; q < d; ++q) {
          const double x0 = xc_s[q * 128 + cbase + g], x1 = xc_s[q * 128 + cbase + 8 + g];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const double2 xr = *reinterpret_cast<const double2*>(xt_s + q * OB + 8 * nt + 2 * t);
            const double a0 = xr.x - x0, a1 = xr.y - x0, b0 = xr.x - x1, b1 = xr.y - x1;
            acc[0][nt][0] = fma(a0, a0, acc[0][nt][0]);
            acc[0][nt][1] = fma(a1, a1, acc[0][nt][1]);
            acc[1][nt][0] = fma(b0, b0, acc[1][nt][0]);
            acc[1][nt][1] = fma(b1, b1, acc[1][nt][1]);
          }
        }
#pragma unroll
        for (int ntq = 0; ntq < 2; ++ntq) {
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
                const double v = live[u][e] ? kv[2 * u + e] : 0.0;
                mu[mi] = fma(v, al[u][e], mu[mi]);
                acc[mi][4 * ntq + u][e] = v;
              }
          }
        }

        // ---- subtract sum_k L_ik V_k: six s32 accumulators per (candidate, row), Horner in 2^-8, then the two power-of-two scales
        if (i > 0) {
          mbar_wait(&acc_full, uu & 1);
          ++uu;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            int a[OS][4];
#pragma unroll
            for (int w = 0; w < OS; ++w) tmem_ld_frag(trow + (uint32_t)(w * 64 + nt * 8), a[w]);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            const double2 rs = *reinterpret_cast<const double2*>(scale_s + 8 * nt + 2 * t);
            const double s0 = -(rs.x * from_fix16), s1 = -(rs.y * from_fix16);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              // weights 0..2 and 3..5 combined exactly in 64-bit integers (< 2^48), converted by the 1.5 * 2^52 trick (I2F.F64 runs
              // at a quarter of the DFMA rate), joined by one FMA
              const long long hi3 = ((long long)a[0][x] << 16) + ((long long)a[1][x] << 8) + (long long)a[2][x];
              const long long lo3 = ((long long)a[3][x] << 16) + ((long long)a[4][x] << 8) + (long long)a[5][x];
              const double dh = __longlong_as_double(0x4338000000000000LL + hi3) - 6755399441055744.0;
              const double dl = __longlong_as_double(0x4338000000000000LL + lo3) - 6755399441055744.0;
              const double h = fma(dl, 1.0 / 16777216.0, dh);
              acc[x >> 1][nt][x & 1] = fma((x & 1) ? s1 : s0, h, acc[x >> 1][nt][x & 1]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty);
        }

        // ---- diagonal block, in registers and in place: V^T[:, nt] = sum_{kt <= nt} rhs^T[:, kt] * Dinv^T[kt, nt]
        {
          const double2* frag = reinterpret_cast<const double2*>(frag_s) + lane;
          int fi = 0;
#pragma unroll
          for (int G = 1; G >= 0; --G) {
            double tmp[4][2][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) tmp[u][mi][0] = tmp[u][mi][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt <= 4 * G + 3; ++kt) {
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
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) {
                const double v0 = tmp[u][mi][0], v1 = tmp[u][mi][1];
                acc[mi][4 * G + u][0] = v0;
                acc[mi][4 * G + u][1] = v1;
                ss[mi] = fma(v0, v0, ss[mi]);
                ss[mi] = fma(v1, v1, ss[mi]);
              }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb_empty[buf]);
        ++sbc;

        // ---- publish the digits of V_i for the later block rows.  Position of row 8 nt + 2 t + e inside the 64-deep block:
        // 16 t + 2 nt + e (the L images carry the same permutation), so a thread's sixteen values are one 16-byte row per plane.
        if (i + 1 < nb) {
          signed char* img = Vpanel + (size_t)i * VIMG + (size_t)(t >> 1) * VHALF + (size_t)(((t & 1) * 16 + (cbase >> 3)) * 128 + g * 16);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            uint32_t lo[8][2], hi[8][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const double v = fmin(fmax(acc[mi][nt][e], -vmax), vmax);
                const double z = fma(v, to_fix, 6755399441055744.0);          // 1.5 * 2^52 + round(v 2^47 / vs)
                const unsigned long long u =
                    (((unsigned long long)(unsigned)__double2hiint(z) << 32) | (unsigned)__double2loint(z)) + 0x0000808080808080ULL;
                lo[nt][e] = (uint32_t)u;             // bytes 0..3: digits 5..2 (+128)
                hi[nt][e] = (uint32_t)(u >> 32);     // bytes 0..1: digits 1..0 (+128)
              }
            uint32_t plw[OS][4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint32_t p01 = __byte_perm(lo[2 * w][0], lo[2 * w][1], 0x5140), r01 = __byte_perm(lo[2 * w + 1][0], lo[2 * w + 1][1], 0x5140);
              const uint32_t p23 = __byte_perm(lo[2 * w][0], lo[2 * w][1], 0x7362), r23 = __byte_perm(lo[2 * w + 1][0], lo[2 * w + 1][1], 0x7362);
              const uint32_t ph = __byte_perm(hi[2 * w][0], hi[2 * w][1], 0x5140), rh = __byte_perm(hi[2 * w + 1][0], hi[2 * w + 1][1], 0x5140);
              plw[5][w] = __byte_perm(p01, r01, 0x5410) ^ 0x80808080u;
              plw[4][w] = __byte_perm(p01, r01, 0x7632) ^ 0x80808080u;
              plw[3][w] = __byte_perm(p23, r23, 0x5410) ^ 0x80808080u;
              plw[2][w] = __byte_perm(p23, r23, 0x7632) ^ 0x80808080u;
              plw[1][w] = __byte_perm(ph, rh, 0x5410) ^ 0x80808080u;
              plw[0][w] = __byte_perm(ph, rh, 0x7632) ^ 0x80808080u;
            }
#pragma unroll
            for (int s = 0; s < OS; ++s) *reinterpret_cast<uint4*>(img + (size_t)s * VPL + (size_t)mi * 128) = make_uint4(plw[s][0], plw[s][1], plw[s][2], plw[s][3]);
          }
          // generic-proxy stores -> the producer's bulk copies (async proxy): proxy fence, then release through the mbarrier.
          // (No gpu-scope __threadfence: its CCTL.IVALL / MEMBAR were 48 % of this phase's stall samples.)
          asm volatile("fence.proxy.async;\n" ::: "memory");
        }
        mbar_arrive(&v_pub);
      }
      // ---- finalize the item: the four t-lanes of a candidate hold its partial sums
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        double v = ss[mi], w = mu[mi];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        w += __shfl_xor_sync(0xffffffffu, w, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        w += __shfl_xor_sync(0xffffffffu, w, 2);
        const int c = tile * 128 + cbase + 8 * mi + g;
        if (t == 0 && c < p.N) {
          const size_t o = (size_t)j * p.N + c;
          if (p.mean) p.mean[o] = w + gp.ymean;
          if (p.var) {
            double vv = variance - v;
            if (p.flags & BOPL_VAR_INCLUDE_NOISE) vv = gp.noise + vv;
            if (p.flags & BOPL_VAR_CLIP) vv = fmax(vv, 1e-10);
            p.var[o] = vv;
          }
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

int ozaki_stage_doubles(int d) { return oz_sb_doubles(d); }

// Host side of the split: digit images of the off-diagonal blocks of the (front-padded, un-negated) factor Lp [n_pad x n_pad]
// and the per-row power-of-two scales.  img: nb (nb-1)/2 blocks of LIMG bytes; column k of a block sits at position
// 16 ((k % 8) / 2) + 2 (k / 8) + (k % 2) of its 64-deep k range (the order the compute warps publish V in).
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
          const int kp = 16 * ((k % 8) / 2) + 2 * (k / 8) + (k % 2);
          long long t = std::llrint(std::ldexp(Lp[(size_t)R * n_pad + kb * OB + k] * inv, 47));
          for (int s = OS - 1; s >= 0; --s) {
            const int dg = (int)((t + 128) & 255) - 128;
            t = (t - dg) >> 8;
            const int rr = s * OB + r;
            base[((size_t)(kp / 16) * (OS * OB / 8) + rr / 8) * 128 + (rr % 8) * 16 + (kp % 16)] = (signed char)dg;
          }
        }
      }
    }
}

// Per-block-row staging array of the kernel: [n_pad/64][36 Dinv B fragments (64 doubles each) | Xs^T [d][64] | alpha[64] | scale[64]].
// Fragment order = consumption order of the diagonal solve: G = 1, 0; kt = 0..4G+3; u = 0..3 with kt <= 4G+u: block (nt = 4G+u, kt),
// lane 4g+t -> Dinv[8nt+g][8kt+2t .. +1].
void ozaki_pack_stage(const double* Dinv64 /*[nb][64][64]*/, const double* Xs /*n_pad*d*/, const double* alpha, const double* scale,
                      int n_pad, int d, std::vector<double>& out) {
  const int nb = n_pad / OB, SBD = oz_sb_doubles(d);
  out.assign((size_t)nb * SBD, 0.0);
  for (int b = 0; b < nb; ++b) {
    double* o = out.data() + (size_t)b * SBD;
    const double* D = Dinv64 + (size_t)b * OB * OB;
    size_t fi = 0;
    for (int G = 1; G >= 0; --G)
      for (int kt = 0; kt <= 4 * G + 3; ++kt)
        for (int u = 0; u < 4; ++u) {
          if (kt > 4 * G + u) continue;
          const int nt = 4 * G + u;
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            o[fi * 64 + 2 * lane + 0] = D[(size_t)(8 * nt + g) * OB + 8 * kt + 2 * t];
            o[fi * 64 + 2 * lane + 1] = D[(size_t)(8 * nt + g) * OB + 8 * kt + 2 * t + 1];
          }
          ++fi;
        }
    double* xt = o + NFRAG * 64;
    for (int r = 0; r < OB; ++r) {
      const int R = b * OB + r;
      for (int q = 0; q < d; ++q) xt[(size_t)q * OB + r] = Xs[(size_t)R * d + q];
      xt[(size_t)d * OB + r] = alpha[R];
      xt[(size_t)(d + 1) * OB + r] = scale[R];
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
  {
    // L2 budget as in posterior_trsm_t.cu: the digit image of the running attribute's factor stays resident; of the rest every CTA
    // keeps the first blocks of its solved panel — the ones re-read most often
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device);
    const double limg = (double)nb * (nb - 1) / 2 * LIMG;
    const double budget = (h->trsm_l2_keep_frac * (double)l2 - 1.5 * limg) / grid;
    long long blocks = (long long)(budget / (double)VIMG);
    if (h->trsm_keep_blocks >= 0) blocks = h->trsm_keep_blocks;
    p.keep_blocks = (int)std::max<long long>(0, std::min<long long>(blocks, nb));
  }
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
