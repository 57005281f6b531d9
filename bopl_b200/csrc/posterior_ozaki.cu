// posterior_ozaki.cu — kernels (1)+(2) of the hot path with the many-right-hand-side update on the INT8 tensor pipe
// (tcgen05.mma kind::i8, TMEM accumulators) by an error-free split ("Ozaki scheme") of the fp64 operands.  Selected with
// bopl_set_posterior_kernel(h, BOPL_POSTERIOR_INT8 | BOPL_POSTERIOR_INT8_7) (or BOPL_POSTERIOR=int8|int8x7) before the model is
// uploaded or fitted; the fp64 DMMA kernel (posterior_trsm_t.cu) stays the library default.  DESIGN.md §3.1b.
//
// Same mathematics as posterior_trsm_t.cu (posterior.py:299-320, gp.py:380-435, linalg.py:92-111): left-looking blocked forward
// substitution V_i = Dinv_i (K*_i - sum_{k<i} L_ik V_k), 64-row blocks, tile = 128 candidates, then mean / variance.
// What changes is the update sum_k L_ik V_k (n^2/2 FMAs per candidate, 97 % of the work).  sm_100a has no fp64 tcgen05 kind,
// but its int8 kind is 128x faster than the fp64 pipe, and an fp64 product can be split into exact int8 products:
//   * every row of the off-diagonal part of L is scaled by the smallest power of two ls_r >= max|L_rk| / 0.995 and written in S = 6 (or 7) balanced
//     base-256 digits (int8): L_rk = ls_r * sum_s d_s 2^(-7-8s)  (47 / 55 fractional bits; once per model, host or device);
//   * |V| <= sqrt(variance) (because var = variance - |V|^2 >= 0), so V is written in fixed point against vs = 2^ceil(log2(sigma / 0.995))
//     the same way, on the fly, by the warp that solved it;
//   * digit products are accumulated EXACTLY in s32 by the tensor core, one accumulator per digit weight g = s_L + s_V < S
//     (|acc| <= S * n * 2^14 < 2^31 for n <= 16384), and combined once per block row in fp64.
// Dropped terms (g >= S) are below 2^(-8S) of ls_r * vs * n: measured against the fp64 kernel on 2^20 candidates of the benchmark
// shape, 1.9e-11 relative on the variance with 6 digits and 1.4e-13 with 7 (tests/test_gpu_int8_posterior.py).
//
// One instruction covers several digit pairs: the S digit planes of an L block are stacked as ONE 64 S-row B operand, so
// V digit a against L digits 0..S-1-a is a single N = 64 (S - a) MMA (split at N = 256) whose output columns are exactly the
// accumulators g = a..S-1 — 8 (10) instructions per 32-deep k-step instead of 21 (28).
//
// Roles (320 threads, one persistent CTA per SM; by default launched as clusters of two CTAs that share one candidate tile and its solved
// panel — see the kernel's header comment): warp 0 lane 0 streams operand k-steps with cp.async.bulk (L digit images from
// the model, V digit images from the scratch panel) through a 3- or 4-stage ring plus one staging block per block row;
// warp 1 lane 0 issues the MMAs; warps 2-9 each own 16 candidates x the 64 rows of the block row in the DMMA C-fragment layout
// (tcgen05.ld.16x256b delivers the accumulators in exactly that layout): K* into the fragments, accumulators recombined and
// subtracted, inverted diagonal block applied in place on DMMA, mean and sum V^2, digits of the solved block published.
// The FP64 pipe and the int8 tensor pipe time-share on this part (tools/fp64_vs_umma_probe.cu), so the launch costs about
// (tensor time) + (fp64 time); profiles/r01_summary.md §A has the measurements.
#include <atomic>
#include <thread>

#include "pipeline.cuh"

namespace bopl {

namespace {

constexpr int OB = 64;                       // rows per block row
constexpr int OZ_MAXD = MAX_D;               // input dimensions the shared-memory plan is sized for (3 ring stages; 4 fit for 6 digits and d <= 12)
// S balanced base-256 digits hold |t| <= 127 (256^S - 1) / 255 = 0.498 * 2^(8S): a value may use up to OZ_FILL of its power-of-two scale
constexpr double OZ_FILL = 0.995;
constexpr int VPL = 128 * 32;                // 4096 B: one digit plane of V, 128 candidates x 32 k
// S = digits per operand: 6 (47 fractional bits; 21 digit pairs) or 7 (55 bits, i.e. every bit of an fp64 mantissa; 28 pairs)
template <int S>
struct Oz {
  static constexpr int LIMG = S * OB * OB;          // S = 6: 24576 B: 384 stacked rows x 64 k, one off-diagonal block of L
  static constexpr int LHALF = LIMG / 2;            // one k-step (32 k) of it
  static constexpr int VHALF = S * VPL;             // the S planes of one k-step
  static constexpr int VIMG = 2 * VHALF;            // per solved block
  static constexpr int OSTAGE = LHALF + VHALF;      // per pipeline stage = one k-step of one (block row, block) pair (36864 / 43008 B)
};
constexpr int NFRAG = 36;                    // lower-triangular 8x8 blocks of a 64x64 diagonal block
constexpr int ONT = 320;                     // producer warp, MMA warp, eight compute warps
constexpr int NCW = 8;
constexpr int PROD_WARP = 0, MMA_WARP = 1;   // lane 0 of each: bulk-copy producer, tcgen05.mma issuer; warps 2..9 compute

struct OzParams {
  const GpDev* gps;
  const double* Xc;
  double* mean;
  double* var;
  signed char* scratch;   // [grid][n_pad/64][VIMG]
  int H, hsel, m, N, d, n_pad, pad, ntiles, flags;
  int keep_blocks;        // solved blocks below this index are kept in L2 (evict_last), the rest is streamed
  long long* trace;       // TRACE instantiation only
};

// staged per block row: Dinv B fragments, scaled training rows transposed [q][64], alpha[64], row scales[64]
__host__ __device__ inline int oz_sb_doubles(int d) { return NFRAG * 64 + (d + 2) * OB; }
constexpr int OZ_XBUF = 2 * 8 * 32 * 8;      // partial sums handed over inside a CTA pair: [2][compute warps][32] doubles
template <int S>
__host__ __device__ inline size_t oz_smem_bytes(int d, int nst) {
  return (size_t)nst * Oz<S>::OSTAGE + (size_t)2 * oz_sb_doubles(d) * 8 + (size_t)d * 128 * 8 + OZ_XBUF;
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
// arrive on the mbarrier at this offset in both CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  const unsigned short mask = 3;
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
// the same into this CTA and its cluster peer (same CTA-relative destination and mbarrier offsets in both)
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
  const unsigned short mask = 3;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask), "l"(pol)
               : "memory");
}
// 16 TMEM lanes x 8 columns: lane l/4 (+8), columns 2 (l%4), 2 (l%4) + 1 — the C-fragment layout of an 8x8 DMMA tile pair
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}

// ---- cluster-scope helpers of the row-paired launch
__device__ __forceinline__ unsigned cluster_peer_addr(const void* p, unsigned rank) {        // same offset in the peer CTA's shared memory
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(unsigned long long* bar) {       // local barrier, visible to waiters in the peer
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.release.cluster.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_remote_f64(unsigned cluster_addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(cluster_addr), "d"(v) : "memory");
}
// waits whose arrivals may come from the peer CTA: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_wait_cluster_backoff(unsigned long long* bar, unsigned parity, unsigned ns) {
  unsigned done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void mbar_arrive_plain(unsigned long long* bar) { mbar_arrive(bar); }

// posterior_ozaki_kernel<OS digits, NST ring stages, TRACE, RP>
//
// RP (row-paired, the default launch): clusters of TWO CTAs work on the SAME 128-candidate tile; CTA r of the pair owns the block rows
// i = 2 I + r of every "super-row" I.  What the pair shares is the solved panel V: for the blocks kb < 2 I both CTAs need the same V
// k-steps, so each fetches HALF of every V k-step and multicasts it into both shared memories, while the L k-steps (its own rows) are
// private.  The launch is bound by L2 bandwidth (profiles/r02_summary.md): per 32-deep k-step an SM pulls 14.3 (L) + 14.3 (half V) KB
// instead of 7.2 (half L, the candidate-paired form of round 1) + 28.7 (V), and the pair keeps ONE panel (74 instead of 148 in
// flight: they now fit the L2, so the DRAM re-reads of round 1 go away).  Block 2 I is produced by CTA 0 while the pair is in super-row
// I and is the last block CTA 1 needs for row 2 I + 1: CTA 1 appends two private k-steps (CTA 0 walks the same two ring slots
// without operands so that both rings stay in step).  Hand-offs inside the pair: the "block published" mbarrier of BOTH CTAs is
// arrived on by the publishing CTA's compute warps (local + remote release.cluster arrive); ring stages are released by both MMA
// threads (multicast tcgen05.commit); the per-candidate partial sums (mean, sum V^2: even block rows in CTA 0, odd ones in CTA 1)
// are combined through the peer's shared memory at the end of the tile.
// !RP: one CTA per tile, every block row; same arithmetic in the same order (partial sums of even and odd block rows are kept apart
// and combined as the pair does), so a candidate's numbers are bit-identical in both launches.
// TRACE (!RP only): CTA 0 records clock64 at its phase boundaries into p.trace (tools/ozaki_trace.py; BOPL_OZ_TRACE=1), slot layout
// [row][24]: 8.. every compute warp's TMEM release, 16.. every compute warp's K* done; 0 K* start, 1 K* done, 2 accumulators full, 3 recombined (TMEM released), 4 diagonal solve done, 5 digits published
// (compute warp 2, lane 0); 6 first MMA of the row issued, 7 last MMA of the row committed (MMA thread).  First work item only.
template <int OS, int NST, bool TRACE, bool RP>
__global__ void __launch_bounds__(ONT, 1) posterior_ozaki_kernel(OzParams p) {
  constexpr int LIMG = Oz<OS>::LIMG, LHALF = Oz<OS>::LHALF, VHALF = Oz<OS>::VHALF, VIMG = Oz<OS>::VIMG, OSTAGE = Oz<OS>::OSTAGE;
  static_assert(!(TRACE && RP), "the phase trace is recorded by the single-CTA launch");
  extern __shared__ __align__(128) unsigned char osm[];
  const int d = p.d, n_pad = p.n_pad, nb = n_pad / OB;
  const int SBD = oz_sb_doubles(d);
  unsigned char* stage_s = osm;                                              // [NST][L k-step | V k-step]
  double* SB = (double*)(osm + (size_t)NST * OSTAGE);                        // [2][SBD]
  double* xc_s = SB + 2 * SBD;                                               // [d][128] candidates / lengthscale
  double* xbuf = xc_s + (size_t)d * 128;                                     // [2][NCW][32] partial sums received from the peer (RP)
  __shared__ unsigned long long full_bar[NST], empty_bar[NST], sb_full[2], sb_empty[2], acc_full, tmem_empty, v_pub[2], x_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RP ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sb_full[s], 1);
      mbar_init(&sb_empty[s], NCW);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&tmem_empty, NCW);
    mbar_init(&v_pub[0], RP ? NCW : NCW * 32);   // block i is published on v_pub[i & 1]: with ONE barrier a producer of the pair could fall two
    mbar_init(&v_pub[1], RP ? NCW : NCW * 32);   // phases behind (blocks 2 I and 2 I + 1 are both out before it asks for 2 I) and the parity would alias
    mbar_init(&x_full, NCW);
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
  unsigned crank = 0;
  if (RP) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");   // the peer's barriers are initialised
  }
  // work items: (attribute, candidate tile), one per CTA (or per pair)
  const int nitems = p.m * p.ntiles;
  const int unit = RP ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, istep = RP ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int item0 = unit;
  const int RS = RP ? 2 : 1, r0 = RP ? (int)crank : 0;          // this CTA's block rows: r0, r0 + RS, ...
  signed char* Vpanel = p.scratch + (size_t)unit * nb * VIMG;

  if (warp == PROD_WARP) {
    // ================================================================ producer: operand k-steps and per-block-row staging
    if (lane == 0) {
      unsigned gc = 0, sc = 0;             // ring slots issued; staging blocks issued
      unsigned pubc[2] = {0, 0};           // publications waited for on v_pub[0] (even blocks) / v_pub[1] (odd blocks)
      const unsigned sb_bytes = (unsigned)SBD * 8;
      const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
      auto stage_block = [&](const GpDev& g, int i) {
        const unsigned buf = sc & 1;
        if (sc >= 2) mbar_wait_backoff(&sb_empty[buf], ((sc >> 1) - 1) & 1, 200);
        mbar_expect_tx(&sb_full[buf], sb_bytes);
        bulk_g2s(SB + (size_t)buf * SBD, g.OzStage + (size_t)i * SBD, sb_bytes, &sb_full[buf]);
        ++sc;
      };
      auto wait_published = [&](int blk) {  // block blk (asked for in block order) has its digits in the panel
        const int e = blk & 1;
        if (RP) mbar_wait_cluster_backoff(&v_pub[e], pubc[e] & 1, 32);
        else mbar_wait_backoff(&v_pub[e], pubc[e] & 1, 100);
        ++pubc[e];
        asm volatile("fence.proxy.async;\n" ::: "memory");
      };
      // both k-steps of block (i, kb): the L k-step of this CTA's row, and the V k-step — whole (private) or this CTA's half, multicast
      auto load_block = [&](const signed char* Limg, int i, int kb, bool shared_v) {
        const unsigned long long vpol = kb < p.keep_blocks ? pol_keep : pol_stream;
#pragma unroll 1
        for (int ks = 0; ks < 2; ++ks) {
          const unsigned st = gc % NST, use = gc / NST;
          mbar_wait_backoff(&empty_bar[st], (use & 1) ^ 1, 100);
          unsigned char* dst = stage_s + (size_t)st * OSTAGE;
          mbar_expect_tx(&full_bar[st], OSTAGE);
          bulk_g2s_hint(dst, Limg + ((size_t)i * (i - 1) / 2 + kb) * LIMG + (size_t)ks * LHALF, LHALF, &full_bar[st], pol_keep);
          const signed char* vsrc = Vpanel + (size_t)kb * VIMG + (size_t)ks * VHALF;
          if (RP && shared_v) bulk_g2s_mc(dst + LHALF + (size_t)crank * (VHALF / 2), vsrc + (size_t)crank * (VHALF / 2), VHALF / 2, &full_bar[st], vpol);
          else bulk_g2s_hint(dst + LHALF, vsrc, VHALF, &full_bar[st], vpol);
          ++gc;
        }
      };
      for (int item = item0; item < nitems; item += istep) {
        const GpDev& gp = p.gps[(item / p.ntiles) * p.H + p.hsel];
        const signed char* Limg = gp.OzL;
        if (item == item0) stage_block(gp, r0);
        int seen = 0;                         // blocks of this item whose publication has been waited for
        if (RP) {
          for (int I = 0; 2 * I < nb; ++I) {
            const int i = 2 * I + r0;
            for (int kb = 0; kb < 2 * I; ++kb) {            // the blocks both CTAs of the pair need
              while (seen <= kb) {
                wait_published(seen);
                ++seen;
              }
              load_block(Limg, i, kb, true);
            }
            // staging block of this CTA's next row: its buffer (that of the previous row) is free by now — every block up to
            // 2 I - 1 has been published, i.e. both CTAs are past the diagonal solves of the previous super-row
            if (i + 2 < nb) stage_block(gp, i + 2);
            else if (item + istep < nitems) stage_block(p.gps[((item + istep) / p.ntiles) * p.H + p.hsel], r0);
            if (r0 == 1) {                                   // block 2 I, just solved by CTA 0: private k-steps of CTA 1
              while (seen <= 2 * I) {
                wait_published(seen);
                ++seen;
              }
              load_block(Limg, i, 2 * I, false);
            } else {                                         // CTA 0 walks the same two ring slots without operands
              for (int ks = 0; ks < 2; ++ks) {
                const unsigned st = gc % NST, use = gc / NST;
                mbar_wait_backoff(&empty_bar[st], (use & 1) ^ 1, 100);
                mbar_arrive(&full_bar[st]);
                ++gc;
              }
            }
          }
          while (seen < nb) {                                // every publication is consumed once: keeps the phases in step
            wait_published(seen);
            ++seen;
          }
        } else {
          for (int i = 0; i < nb; ++i) {
            // The staging block of row i+1 reuses the buffer of row i-1, free once that row's diagonal solve is over — which is
            // before its digits are published: issue it behind the wait for those digits, so that the k-steps of the blocks
            // kb < i-1 (needing neither) are prefetched while the previous block row is still being solved.
            for (int kb = 0; kb + 1 < i; ++kb) load_block(Limg, i, kb, false);
            if (i > 0) {                        // the newest block: wait until its digits are in the scratch panel
              wait_published(seen);
              ++seen;
            }
            if (i + 1 < nb) stage_block(gp, i + 1);
            else if (item + istep < nitems) stage_block(p.gps[((item + istep) / p.ntiles) * p.H + p.hsel], 0);
            if (i > 0) load_block(Limg, i, i - 1, false);
          }
          wait_published(seen);                 // the last block row arrives too: keep the phases in step
          ++seen;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      unsigned gc = 0, uu = 0;            // ring slots consumed; accumulator generations completed
      const unsigned lboA = 128 * 16, lboB = OS * OB * 16, sbo = 128;
      // V digit a (A operand, 128 candidates) x L digits 0..S-1-a (B operand rows 64 s ..) -> accumulators a..S-1 (64 columns
      // each); an instruction covers at most 256 columns.  The a = 0 pair touches every accumulator: it clears them on the first k-step.
      auto issue_kstep = [&](unsigned st, uint32_t first) {
        const unsigned lk = smem_u32(stage_s + (size_t)st * OSTAGE), vb = lk + LHALF;
        const uint64_t bl = umma_desc(lk, lboB, sbo);
#pragma unroll
        for (int a = 0; a < OS; ++a) {
          const uint64_t ad = umma_desc(vb + a * VPL, lboA, sbo);
          const int ncols = 64 * (OS - a);
          const uint32_t acc_flag = (a == 0) ? first : 1u;
          if (ncols > 256) {      // two instructions of (nearly) equal width — 224 + 224, 192 + 192, 160 + 160 — rather than 256 + a narrow rest: N = 64 issues at 65 % of the N = 256 rate
            const int n1 = ((ncols / 2 + 15) / 16) * 16;
            umma_i8(tmem + 64 * a, ad, bl, idesc_s8(n1), acc_flag);
            umma_i8(tmem + 64 * a + n1, ad, umma_desc(lk + 16 * n1, lboB, sbo), idesc_s8(ncols - n1), acc_flag);
          } else {
            umma_i8(tmem + 64 * a, ad, bl, idesc_s8(ncols), acc_flag);
          }
        }
      };
      for (int item = item0; item < nitems; item += istep) {
        if (RP) {
          for (int I = 0; 2 * I < nb; ++I) {
            const int i = 2 * I + r0;
            if (i > 0) {
              if (uu > 0) mbar_wait_backoff(&tmem_empty, (uu - 1) & 1, 50);     // the previous row's accumulators have been read
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            uint32_t done = 0;
            for (int ks2 = 0; ks2 < 4 * I; ++ks2) {
              const unsigned st = gc % NST, use = gc / NST;
              mbar_wait_backoff(&full_bar[st], use & 1, 20);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
              issue_kstep(st, done);
              done = 1u;
              umma_commit_pair(&empty_bar[st]);
              ++gc;
            }
            if (r0 == 0) {
              if (i > 0) {
                umma_commit(&acc_full);
                ++uu;
              }
              for (int ks = 0; ks < 2; ++ks) {             // the peer's private k-steps: release the slots, nothing to multiply
                const unsigned st = gc % NST, use = gc / NST;
                mbar_wait_backoff(&full_bar[st], use & 1, 20);
                umma_commit_pair(&empty_bar[st]);
                ++gc;
              }
            } else {
              for (int ks = 0; ks < 2; ++ks) {
                const unsigned st = gc % NST, use = gc / NST;
                mbar_wait_backoff(&full_bar[st], use & 1, 20);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                issue_kstep(st, done);
                done = 1u;
                umma_commit_pair(&empty_bar[st]);
                ++gc;
              }
              umma_commit(&acc_full);
              ++uu;
            }
          }
        } else {
          for (int i = 1; i < nb; ++i) {
            if (uu > 0) mbar_wait_backoff(&tmem_empty, (uu - 1) & 1, 50);     // the previous block row's accumulators have been read
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (TRACE && blockIdx.x == 0 && item == item0) p.trace[i * 24 + 6] = clock64();
            for (int ks2 = 0; ks2 < 2 * i; ++ks2) {
              const unsigned st = gc % NST, use = gc / NST;
              mbar_wait_backoff(&full_bar[st], use & 1, 20);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
              issue_kstep(st, (ks2 == 0) ? 0u : 1u);
              umma_commit(&empty_bar[st]);
              ++gc;
            }
            umma_commit(&acc_full);
            if (TRACE && blockIdx.x == 0 && item == item0) p.trace[i * 24 + 7] = clock64();
            ++uu;
          }
        }
      }
    }
  } else {
    // ================================================================ compute warps: 16 candidates x all 64 rows of the block row,
    // in the DMMA C-fragment layout: candidate cbase + 8 mi + g, row 8 nt + 2 t + e
    const int g = lane >> 2, t = lane & 3;
    const int cbase = 32 * (warp & 3) + 16 * ((warp - 2) >> 2);   // TMEM lanes of this warp: quadrant warp % 4, half (warp-2)/4
    const uint32_t trow = tmem + ((uint32_t)cbase << 16);
    unsigned uu = 0, sbc = 0, itc = 0;
    const unsigned peer_vpub0 = RP ? cluster_peer_addr(&v_pub[0], crank ^ 1u) : 0u;
    for (int item = item0; item < nitems; item += istep, ++itc) {
      const int j = item / p.ntiles, tile = item - j * p.ntiles;
      const GpDev& gp = p.gps[j * p.H + p.hsel];
      const int kind = gp.kind;
      const double variance = gp.variance;
      const double vs = exp2(ceil(log2(sqrt(variance) * (1.0 / OZ_FILL))));   // fixed-point scale of V: |V| <= sigma <= OZ_FILL vs
      const double to_fix = (OS == 6 ? 140737488355328.0 : 36028797018963968.0) / vs;   // 2^47 / vs (2^55 / vs)
      const double from_fix16 = vs * (1.0 / 1073741824.0);       // vs 2^-14 (digit product weight) 2^-16 (integer Horner)
      const double vmax = OZ_FILL * vs;
      __syncwarp();
      for (int idx = lane; idx < 16 * d; idx += 32) {
        const int c = idx / d, q = idx - c * d;
        const int cc = min(tile * 128 + cbase + c, p.N - 1);
        xc_s[q * 128 + cbase + c] = p.Xc[(size_t)cc * d + q] / gp.ell[q];
      }
      __syncwarp();
      // partial sums of this thread's two candidates over the block rows handled so far: (mu, ss) the rows of the running parity,
      // (mu_o, ss_o) the others.  !RP: the two sets swap after every row, so at the end (nb is even) mu = even rows, mu_o = odd rows.
      double mu[2] = {0.0, 0.0}, ss[2] = {0.0, 0.0}, mu_o[2] = {0.0, 0.0}, ss_o[2] = {0.0, 0.0};
      double acc[2][8][2];
#pragma unroll 1
      for (int i = r0; i < nb; i += RS) {
        const int R0 = i * OB;
        const unsigned buf = sbc & 1;
        mbar_wait(&sb_full[buf], (sbc >> 1) & 1);
        const double* frag_s = SB + (size_t)buf * SBD;
        const double* xt_s = frag_s + NFRAG * 64;
        const double* alpha_s = xt_s + d * OB;
        const double* scale_s = alpha_s + OB;

        const bool tr = TRACE && blockIdx.x == 0 && item == item0 && warp == 2 && lane == 0;
        if (tr) p.trace[i * 24 + 0] = clock64();
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

        // ---- subtract sum_k L_ik V_k: S s32 accumulators per (candidate, row), Horner in 2^-8, then the two power-of-two scales
        if (tr) p.trace[i * 24 + 1] = clock64();
        if (TRACE && blockIdx.x == 0 && item == item0 && lane == 0) p.trace[i * 24 + 16 + (warp - 2)] = clock64();     // every warp: K* done
        if (i > 0) {
          mbar_wait(&acc_full, uu & 1);
          if (tr) p.trace[i * 24 + 2] = clock64();
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
              // weights 0..2, 3..5 (3..4, 5..6 for seven digits) combined exactly in 64-bit integers (< 2^48), converted by the
              // 1.5 * 2^52 trick (I2F.F64 runs at a quarter of the DFMA rate), joined by FMAs
              const long long hi3 = ((long long)a[0][x] << 16) + ((long long)a[1][x] << 8) + (long long)a[2][x];
              const double dh = __longlong_as_double(0x4338000000000000LL + hi3) - 6755399441055744.0;
              double h;
              if constexpr (OS == 6) {
                const long long lo3 = ((long long)a[3][x] << 16) + ((long long)a[4][x] << 8) + (long long)a[5][x];
                const double dl = __longlong_as_double(0x4338000000000000LL + lo3) - 6755399441055744.0;
                h = fma(dl, 1.0 / 16777216.0, dh);
              } else {
                const long long m2 = ((long long)a[3][x] << 8) + (long long)a[4][x], l2 = ((long long)a[5][x] << 8) + (long long)a[6][x];
                const double dm = __longlong_as_double(0x4338000000000000LL + m2) - 6755399441055744.0;
                const double dl = __longlong_as_double(0x4338000000000000LL + l2) - 6755399441055744.0;
                h = fma(fma(dl, 1.0 / 65536.0, dm), 1.0 / 65536.0, dh);
              }
              acc[x >> 1][nt][x & 1] = fma((x & 1) ? s1 : s0, h, acc[x >> 1][nt][x & 1]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty);
        }
        if (tr) p.trace[i * 24 + 3] = clock64();
        if (TRACE && blockIdx.x == 0 && item == item0 && lane == 0) p.trace[i * 24 + 8 + (warp - 2)] = clock64();      // every warp: TMEM released

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
        if (tr) p.trace[i * 24 + 4] = clock64();
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
                unsigned long long u;
                if constexpr (OS == 6) {
                  const double z = fma(v, to_fix, 6755399441055744.0);          // 1.5 * 2^52 + round(v 2^47 / vs)
                  u = (((unsigned long long)(unsigned)__double2hiint(z) << 32) | (unsigned)__double2loint(z)) + 0x0000808080808080ULL;
                } else {
                  u = (unsigned long long)__double2ll_rn(v * to_fix) + 0x0080808080808080ULL;   // 55 fractional bits: beyond the trick
                }
                lo[nt][e] = (uint32_t)u;             // bytes 0..3: digits S-1..S-4 (+128)
                hi[nt][e] = (uint32_t)(u >> 32);     // bytes 0..S-5: digits S-5..0 (+128)
              }
            uint32_t plw[OS][4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const uint32_t p01 = __byte_perm(lo[2 * w][0], lo[2 * w][1], 0x5140), r01 = __byte_perm(lo[2 * w + 1][0], lo[2 * w + 1][1], 0x5140);
              const uint32_t p23 = __byte_perm(lo[2 * w][0], lo[2 * w][1], 0x7362), r23 = __byte_perm(lo[2 * w + 1][0], lo[2 * w + 1][1], 0x7362);
              const uint32_t ph = __byte_perm(hi[2 * w][0], hi[2 * w][1], 0x5140), rh = __byte_perm(hi[2 * w + 1][0], hi[2 * w + 1][1], 0x5140);
              plw[OS - 1][w] = __byte_perm(p01, r01, 0x5410) ^ 0x80808080u;
              plw[OS - 2][w] = __byte_perm(p01, r01, 0x7632) ^ 0x80808080u;
              plw[OS - 3][w] = __byte_perm(p23, r23, 0x5410) ^ 0x80808080u;
              plw[OS - 4][w] = __byte_perm(p23, r23, 0x7632) ^ 0x80808080u;
              plw[OS - 5][w] = __byte_perm(ph, rh, 0x5410) ^ 0x80808080u;
              plw[OS - 6][w] = __byte_perm(ph, rh, 0x7632) ^ 0x80808080u;
              if constexpr (OS == 7) {
                const uint32_t ph2 = __byte_perm(hi[2 * w][0], hi[2 * w][1], 0x7362), rh2 = __byte_perm(hi[2 * w + 1][0], hi[2 * w + 1][1], 0x7362);
                plw[0][w] = __byte_perm(ph2, rh2, 0x5410) ^ 0x80808080u;
              }
            }
#pragma unroll
            for (int s = 0; s < OS; ++s) *reinterpret_cast<uint4*>(img + (size_t)s * VPL + (size_t)mi * 128) = make_uint4(plw[s][0], plw[s][1], plw[s][2], plw[s][3]);
          }
          // generic-proxy stores -> the producers' bulk copies (async proxy): proxy fence, then release through the mbarrier(s).
          // (No gpu-scope __threadfence: its CCTL.IVALL / MEMBAR were 48 % of this phase's stall samples.)
          asm volatile("fence.proxy.async;\n" ::: "memory");
        }
        if (RP) {             // the block is published to BOTH producers of the pair: the local barrier first, then the peer's
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_release_cluster(&v_pub[i & 1]);
            mbar_arrive_remote(peer_vpub0 + 8u * (unsigned)(i & 1));
          }
        } else {
          mbar_arrive(&v_pub[i & 1]);
        }
        if (tr) p.trace[i * 24 + 5] = clock64();
        if (!RP) {            // next row has the other parity
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const double a = mu[mi], b = ss[mi];
            mu[mi] = mu_o[mi];
            ss[mi] = ss_o[mi];
            mu_o[mi] = a;
            ss_o[mi] = b;
          }
        }
      }
      // ---- finalize the item: the four t-lanes of a candidate hold its partial sums; even block rows + odd block rows
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        double v = ss[mi], w = mu[mi];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        w += __shfl_xor_sync(0xffffffffu, w, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        w += __shfl_xor_sync(0xffffffffu, w, 2);
        ss[mi] = v;
        mu[mi] = w;
        if (!RP) {
          double vo = ss_o[mi], wo = mu_o[mi];
          vo += __shfl_xor_sync(0xffffffffu, vo, 1);
          wo += __shfl_xor_sync(0xffffffffu, wo, 1);
          vo += __shfl_xor_sync(0xffffffffu, vo, 2);
          wo += __shfl_xor_sync(0xffffffffu, wo, 2);
          ss_o[mi] = vo;
          mu_o[mi] = wo;
        }
      }
      if (RP) {
        double* xb = xbuf + (size_t)(itc & 1) * (NCW * 32) + (size_t)(warp - 2) * 32;
        if (crank == 1) {               // odd block rows: hand the sums to CTA 0 through its shared memory
          if (t == 0) {
            const unsigned ra = cluster_peer_addr(xb + 4 * g, 0);
            st_remote_f64(ra, mu[0]);
            st_remote_f64(ra + 8, mu[1]);
            st_remote_f64(ra + 16, ss[0]);
            st_remote_f64(ra + 24, ss[1]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(cluster_peer_addr(&x_full, 0));
        } else {
          mbar_wait_cluster(&x_full, itc & 1);
          mu_o[0] = xb[4 * g];
          mu_o[1] = xb[4 * g + 1];
          ss_o[0] = xb[4 * g + 2];
          ss_o[1] = xb[4 * g + 3];
        }
      }
      if (!RP || crank == 0) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const double v = ss[mi] + ss_o[mi], w = mu[mi] + mu_o[mi];
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (RP) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");   // no CTA leaves while its peer may still write into it
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace

bool ozaki_covers(int d, int n_pad) { return d <= OZ_MAXD && n_pad <= 16384; }

bool ozaki_available(const bopl_handle* h) { return h->has_oz && ozaki_covers(h->d, h->n_pad); }

int ozaki_stage_doubles(int d) { return oz_sb_doubles(d); }

// Host side of the split: S-digit images of the off-diagonal blocks of the (front-padded, un-negated) factor Lp [n_pad x n_pad]
// and the per-row power-of-two scales.  img: nb (nb-1)/2 blocks of S * 4096 bytes; column k of a block sits at position
// 16 ((k % 8) / 2) + 2 (k / 8) + (k % 2) of its 64-deep k range (the order the compute warps publish V in).
void ozaki_prepare(const double* Lp, int n_pad, int S, std::vector<signed char>& img, std::vector<double>& scale) {
  const int nb = n_pad / OB;
  const size_t limg = (size_t)S * OB * OB;
  img.assign((size_t)nb * (nb - 1) / 2 * limg, 0);
  scale.assign((size_t)n_pad, 1.0);
  for (int R = 0; R < n_pad; ++R) {
    const int i = R / OB;
    double mx = 0.0;
    for (int k = 0; k < i * OB; ++k) mx = std::max(mx, std::fabs(Lp[(size_t)R * n_pad + k]));
    if (mx > 0.0) {
      int e;
      const double f = std::frexp(mx, &e);                       // mx = f 2^e, 0.5 <= f < 1
      scale[R] = std::ldexp(1.0, f <= OZ_FILL ? e : e + 1);       // |x| / scale <= OZ_FILL: the digits of x / scale 2^(8S-1) fit
    }
  }
  // the block rows are independent: spread them over the host cores (largest first)
  std::atomic<int> next(nb - 1);
  auto work = [&]() {
    for (int i = next.fetch_sub(1); i >= 1; i = next.fetch_sub(1))
      for (int kb = 0; kb < i; ++kb) {
        signed char* base = img.data() + ((size_t)i * (i - 1) / 2 + kb) * limg;
        for (int r = 0; r < OB; ++r) {
          const int R = i * OB + r;
          const double inv = 1.0 / scale[R];
          for (int k = 0; k < OB; ++k) {
            const int kp = 16 * ((k % 8) / 2) + 2 * (k / 8) + (k % 2);
            long long t = std::llrint(std::ldexp(Lp[(size_t)R * n_pad + kb * OB + k] * inv, 8 * S - 1));
            for (int s = S - 1; s >= 0; --s) {
              const int dg = (int)((t + 128) & 255) - 128;
              t = (t - dg) >> 8;
              const int rr = s * OB + r;
              base[((size_t)(kp / 16) * (S * OB / 8) + rr / 8) * 128 + (rr % 8) * 16 + (kp % 16)] = (signed char)dg;
            }
          }
        }
      }
  };
  const int nthr = std::max(1, std::min<int>(nb - 1, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
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

// ---------------------------------------------------------------------------------------------- device-side preparation
// The same images and staging blocks from a model fitted on the device (bopl_fit_model): A holds the factor with its off-diagonal
// 128-blocks negated and the diagonal 128-blocks as they are; the inverse of a 64 x 64 diagonal block is the matching diagonal
// sub-block of the inverted 128 x 128 block.
namespace {

struct OzPrep {
  const double* A;
  const double* Dinv128;
  const double* Xs;
  const double* alpha;
  signed char* img;
  double* stage;
};

// power-of-two row scales (smallest with max |L_Rk| <= OZ_FILL scale, over the columns left of the row's 64 x 64 diagonal block).  grid (n_pad / 8, G) x 256
__global__ void oz_scale_kernel(const OzPrep* pp, int n_pad, int d) {
  const OzPrep& q = pp[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, R = blockIdx.x * 8 + warp;
  const int cols = OB * (R / OB);
  double mx = 0.0;
  for (int k = lane; k < cols; k += 32) mx = fmax(mx, fabs(q.A[(size_t)R * n_pad + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    const int E = (__double2hiint(mx) >> 20) & 0x7ff;      // mx = 1.f 2^(E-1023) = 0.1f 2^(E-1022): frexp exponent E - 1022
    const double p2 = __hiloint2double((E + 1) << 20, 0);  // 2^(E - 1022) > mx
    const double sc = (mx > 0.0 && E > 0) ? (mx <= OZ_FILL * p2 ? p2 : 2.0 * p2) : 1.0;   // as ozaki_prepare
    q.stage[(size_t)(R / OB) * oz_sb_doubles(d) + NFRAG * 64 + (d + 1) * OB + (R % OB)] = sc;
  }
}

// digit image of off-diagonal block (i, kb), built in shared memory, written out with 16-byte stores.  grid (nb (nb-1) / 2, G) x 256
template <int S>
__global__ void oz_digits_kernel(const OzPrep* pp, int n_pad, int d) {
  __shared__ __align__(16) signed char im[S * OB * OB];
  const OzPrep& q = pp[blockIdx.y];
  int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)blockIdx.x)) * 0.5);
  while (i * (i - 1) / 2 > (int)blockIdx.x) --i;
  while ((i + 1) * i / 2 <= (int)blockIdx.x) ++i;
  const int kb = (int)blockIdx.x - i * (i - 1) / 2;
  const int SBD = oz_sb_doubles(d);
  for (int e = threadIdx.x; e < OB * OB; e += 256) {
    const int r = e >> 6, k = e & 63, R = i * OB + r, C = kb * OB + k;
    double v = q.A[(size_t)R * n_pad + C];
    if ((R >> 7) != (C >> 7)) v = -v;
    const double sc = q.stage[(size_t)i * SBD + NFRAG * 64 + (d + 1) * OB + r];
    const long long t = __double2ll_rn(ldexp(v / sc, 8 * S - 1));
    const unsigned long long u = ((unsigned long long)t + (S == 6 ? 0x0000808080808080ULL : 0x0080808080808080ULL)) ^
                                 (S == 6 ? 0x0000808080808080ULL : 0x0080808080808080ULL);
    const int kp = 16 * ((k & 7) >> 1) + 2 * (k >> 3) + (k & 1);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int rr = s * OB + r;
      im[((kp >> 4) * (S * OB / 8) + (rr >> 3)) * 128 + (rr & 7) * 16 + (kp & 15)] = (signed char)(u >> (8 * (S - 1 - s)));
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(q.img + (size_t)blockIdx.x * (S * OB * OB));
  for (int e = threadIdx.x; e < S * OB * OB / 16; e += 256) dst[e] = reinterpret_cast<const uint4*>(im)[e];
}

// staging block of block row b: Dinv fragments, scaled training rows transposed, alpha (the scales are already there).  grid (nb, G) x 256
__global__ void oz_stage_kernel(const OzPrep* pp, int n_pad, int d) {
  const OzPrep& q = pp[blockIdx.y];
  const int b = blockIdx.x, lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  double* o = q.stage + (size_t)b * oz_sb_doubles(d);
  const double* D = q.Dinv128 + (size_t)(b >> 1) * TRSM_BM * TRSM_BM + (size_t)(b & 1) * OB * (TRSM_BM + 1);   // sub-block (b&1, b&1)
  const int g = lane >> 2, t = lane & 3;
  int fi = 0;
  for (int G = 1; G >= 0; --G)
    for (int kt = 0; kt <= 4 * G + 3; ++kt)
      for (int u = 0; u < 4; ++u) {
        if (kt > 4 * G + u) continue;
        if ((fi & 7) == grp) {
          const int nt = 4 * G + u;
          o[fi * 64 + 2 * lane + 0] = D[(size_t)(8 * nt + g) * TRSM_BM + 8 * kt + 2 * t];
          o[fi * 64 + 2 * lane + 1] = D[(size_t)(8 * nt + g) * TRSM_BM + 8 * kt + 2 * t + 1];
        }
        ++fi;
      }
  double* xt = o + NFRAG * 64;
  for (int e = threadIdx.x; e < OB * (d + 1); e += 256) {
    const int r = e / (d + 1), c = e - r * (d + 1), R = b * OB + r;
    xt[(size_t)c * OB + r] = c < d ? q.Xs[(size_t)R * d + c] : q.alpha[R];
  }
}

}  // namespace

// Builds OzL / OzStage of every GP in h->gp_host from its device-resident factor (call before the GpDev array is uploaded).
int ozaki_prepare_device(bopl_handle* h, cudaStream_t st) {
  const int G = (int)h->gp_host.size(), n_pad = h->n_pad, d = h->d, nb = n_pad / OB, S = h->oz_digits;
  if (!ozaki_covers(h->d, n_pad) || nb < 1) return BOPL_OK;   // the kernel does not cover these shapes: fp64 kernel only
  std::vector<OzPrep> prep(G);
  const size_t img_bytes = std::max<size_t>(16, (size_t)nb * (nb - 1) / 2 * S * OB * OB), stage_bytes = (size_t)nb * oz_sb_doubles(d) * 8;
  for (int gi = 0; gi < G; ++gi) {
    void *pi = nullptr, *ps = nullptr;
    BOPL_CUDA(h, cudaMalloc(&pi, img_bytes));
    h->owned.push_back(pi);
    BOPL_CUDA(h, cudaMalloc(&ps, stage_bytes));
    h->owned.push_back(ps);
    GpDev& gp = h->gp_host[gi];
    prep[gi] = OzPrep{gp.Lneg, gp.Dinv, gp.Xs, gp.alpha, (signed char*)pi, (double*)ps};
    gp.OzL = (const signed char*)pi;
    gp.OzStage = (const double*)ps;
  }
  OzPrep* pd = nullptr;
  BOPL_CUDA(h, cudaMalloc((void**)&pd, sizeof(OzPrep) * G));
  h->owned.push_back(pd);
  BOPL_CUDA(h, cudaMemcpyAsync(pd, prep.data(), sizeof(OzPrep) * G, cudaMemcpyHostToDevice, st));
  oz_scale_kernel<<<dim3(n_pad / 8, G), 256, 0, st>>>(pd, n_pad, d);
  BOPL_LAUNCH_CHECK(h, "oz_scale_kernel");
  if (nb > 1) {
    if (S == 7) oz_digits_kernel<7><<<dim3(nb * (nb - 1) / 2, G), 256, 0, st>>>(pd, n_pad, d);
    else oz_digits_kernel<6><<<dim3(nb * (nb - 1) / 2, G), 256, 0, st>>>(pd, n_pad, d);
    BOPL_LAUNCH_CHECK(h, "oz_digits_kernel");
  }
  oz_stage_kernel<<<dim3(nb, G), 256, 0, st>>>(pd, n_pad, d);
  BOPL_LAUNCH_CHECK(h, "oz_stage_kernel");
  BOPL_CUDA(h, cudaStreamSynchronize(st));     // prep (host vector) must outlive the copy
  return BOPL_OK;
}

namespace {
// One instantiation: S digits, NST ring stages.  RP launch (clusters of two CTAs per candidate tile) unless disabled or refused.
template <int S, int NST>
int launch_oz_n(bopl_handle* h, OzParams& p, cudaStream_t st, size_t& attr) {
  const long long items = (long long)p.ntiles * h->m;
  // Launch form (the results are bit-identical): row-paired clusters put TWO SMs on every work item — 1.7x faster per item when there are fewer
  // items than SM pairs (the reference's own anchor batch: 400 candidates = 12 items at m = 3), and the better form for long launches (one solved
  // panel per pair: a quarter of the DRAM traffic, the board runs at its power limit either way).  In between, one CTA per item can need fewer
  // rounds: a pair round costs ~1/1.65 of a single-CTA round (tools/batch_latency.py, profiles/r02_batch_latency_n2000.json).
  bool rp = h->oz_cluster != 0 && h->sms >= 2;
  if (rp && h->oz_cluster == 1) {
    const long long pairs = h->sms / 2;
    if (items > pairs && items < 8LL * h->sms) {
      const long long r_rp = (items + pairs - 1) / pairs, r_single = (items + h->sms - 1) / h->sms;
      if (1.65 * (double)r_single < (double)r_rp) rp = false;
    }
  }
  const int units = (int)std::min<long long>(items, rp ? h->sms / 2 : h->sms);      // CTA pairs or CTAs, one solved panel each
  const int grid = rp ? 2 * units : units;
  const int nb = h->n_pad / OB;
  int rc = ensure_bytes(h, (void**)&h->oz_scratch, &h->oz_scratch_bytes, (size_t)units * nb * Oz<S>::VIMG);
  if (rc) return rc;
  p.scratch = (signed char*)h->oz_scratch;
  {
    // L2 budget as in posterior_trsm_t.cu: the digit image of the running attribute's factor stays resident; of the rest every unit
    // keeps the first blocks of its solved panel — the ones re-read most often
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device);
    const double limg = (double)nb * (nb - 1) / 2 * Oz<S>::LIMG;
    const double budget = (h->trsm_l2_keep_frac * (double)l2 - 1.5 * limg) / units;
    long long blocks = (long long)(budget / (double)Oz<S>::VIMG);
    if (h->trsm_keep_blocks >= 0) blocks = h->trsm_keep_blocks;
    p.keep_blocks = (int)std::max<long long>(0, std::min<long long>(blocks, nb));
  }
  const size_t smem = oz_smem_bytes<S>(h->d, NST);
  if (smem > attr) {
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_ozaki_kernel<S, NST, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_ozaki_kernel<S, NST, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BOPL_CUDA(h, cudaFuncSetAttribute(posterior_ozaki_kernel<S, NST, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  p.trace = nullptr;
  if (const char* tf = getenv("BOPL_OZ_TRACE")) {      // diagnostic: phase timestamps of CTA 0's first work item -> text file (single-CTA launch)
    long long* tb = nullptr;
    const size_t tbytes = (size_t)nb * 24 * sizeof(long long);
    BOPL_CUDA(h, cudaMalloc((void**)&tb, tbytes));
    BOPL_CUDA(h, cudaMemsetAsync(tb, 0, tbytes, st));
    p.trace = tb;
    const int tgrid = (int)std::min<long long>(items, h->sms);
    rc = ensure_bytes(h, (void**)&h->oz_scratch, &h->oz_scratch_bytes, (size_t)tgrid * nb * Oz<S>::VIMG);
    if (rc) return rc;
    p.scratch = (signed char*)h->oz_scratch;
    posterior_ozaki_kernel<S, NST, true, false><<<tgrid, ONT, smem, st>>>(p);
    BOPL_LAUNCH_CHECK(h, "posterior_ozaki_kernel");
    std::vector<long long> host((size_t)nb * 24);
    BOPL_CUDA(h, cudaStreamSynchronize(st));
    BOPL_CUDA(h, cudaMemcpy(host.data(), tb, tbytes, cudaMemcpyDeviceToHost));
    cudaFree(tb);
    if (FILE* f = fopen(tf, "w")) {
      for (int i = 0; i < nb; ++i) {
        for (int c = 0; c < 24; ++c) fprintf(f, "%lld%c", host[(size_t)i * 24 + c], c == 23 ? '\n' : ' ');
      }
      fclose(f);
    }
    return BOPL_OK;
  }
  if (rp) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ONT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t ce = cudaLaunchKernelEx(&cfg, posterior_ozaki_kernel<S, NST, false, true>, p);
    if (ce != cudaSuccess) {          // a device / driver that refuses this cluster shape: same arithmetic, one CTA per tile, from now on
      (void)cudaGetLastError();
      h->oz_cluster = 0;
      return launch_oz_n<S, NST>(h, p, st, attr);
    }
  } else {
    posterior_ozaki_kernel<S, NST, false, false><<<grid, ONT, smem, st>>>(p);
  }
  BOPL_LAUNCH_CHECK(h, "posterior_ozaki_kernel");
  return BOPL_OK;
}

template <int S>
int launch_oz(bopl_handle* h, OzParams& p, cudaStream_t st) {
  // four ring stages when they fit the 227 KB of shared memory (6 digits, d <= 12), else three
  if constexpr (S == 6) {
    if (oz_smem_bytes<S>(h->d, 4) + 512 <= 227 * 1024) return launch_oz_n<S, 4>(h, p, st, h->oz_attr_smem[0]);
  }
  return launch_oz_n<S, 3>(h, p, st, h->oz_attr_smem[S == 6 ? 1 : 2]);
}
}  // namespace

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
  return h->oz_digits == 7 ? launch_oz<7>(h, p, st) : launch_oz<6>(h, p, st);
}

}  // namespace bopl
