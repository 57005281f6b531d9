// pipeline.cuh — device helpers shared by the posterior kernels: cp.async, mbarrier, DMMA.8x8x4, swizzled tile copies.
#pragma once
#include "common.cuh"

namespace bopl {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
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

// Copy a ROWS x KW tile of doubles (row stride `ld` doubles in global) into shared memory as [row][KW].
// KW == 16: 16-byte chunks XOR-swizzled with (row & 1) << 2 (conflict-free LDS.128 fragments); KW == 8: plain.
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
