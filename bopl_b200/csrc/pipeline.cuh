// pipeline.cuh — device helpers shared by the posterior kernels: cp.async, mbarrier, DMMA.8x8x4, swizzled tile copies.
#pragma once
#include "common.cuh"

namespace bopl {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
// L2 eviction-priority hints (createpolicy): the solved-V panels (148 x 2 MB) do not fit the 126 MB L2; keeping the most
// re-read part (the first block rows of every panel, and the shared factor L) resident and streaming the rest cuts the
// DRAM re-reads of the posterior kernel.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gsrc, unsigned long long pol) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "l"(pol));
}
__device__ __forceinline__ void st_global_v2_hint(double* gdst, double a, double b, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;\n" ::"l"(gdst), "d"(a), "d"(b), "l"(pol));
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// volatile: keeps the DMMA stream ordered against the cp.async / mbarrier instructions around it
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier helpers (shared::cta).  The main-loop pipeline synchronises per stage through two mbarriers instead of a
// CTA-wide __syncthreads per k-step: full[s] (all threads' cp.async of the tile have landed) and empty[s] (every warp
// has issued its last DMMA on the tile), so warps free-run and the DMMA queue never drains at a rendezvous.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Same wait for single-thread roles that sit next to working warps on their SM sub-partition (bulk-copy producer, MMA issuer):
// sleeps between polls so that the spin does not compete with them for issue slots.
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity, unsigned ns) {
  unsigned done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}

// Copy a ROWS x KW tile of doubles (row stride `ld` doubles in global) into shared memory as [row][KW].
// KW == 16: 16-byte chunks XOR-swizzled with (row & 1) << 2 (conflict-free LDS.128 fragments); KW == 8: plain.
template <int ROWS, int KW, int NT>
__device__ __forceinline__ void load_tile_hint(double* dst, const double* src, size_t ld, int tid, unsigned long long pol) {
  constexpr int CPR = KW / 2;
  constexpr int PER = ROWS * CPR / NT;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    int idx = tid + NT * u;
    int row = idx / CPR, c = idx % CPR;
    int cs = (KW == 16) ? (c ^ ((row & 1) << 2)) : c;
    cp_async16_hint(dst + row * KW + (cs << 1), src + (size_t)row * ld + 2 * c, pol);
  }
}

template <int ROWS, int KW, int NT>
__device__ __forceinline__ void load_tile(double* dst, const double* src, size_t ld, int tid) {
  constexpr int CPR = KW / 2;                       // 16-byte chunks per row
  constexpr int PER = ROWS * CPR / NT;
  static_assert(ROWS * CPR % NT == 0, "tile must divide evenly over the threads");
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    int idx = tid + NT * u;
    int row = idx / CPR, c = idx % CPR;
    int cs = (KW == 16) ? (c ^ ((row & 1) << 2)) : c;
    cp_async16(dst + row * KW + (cs << 1), src + (size_t)row * ld + 2 * c);
  }
}

}  // namespace bopl
