// probes.cu — the roofline denominator of the int8-tensor-pipe posterior kernel, measured on the device the handle is bound to:
// every SM issues back-to-back tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32, s32 accumulators in TMEM) on operands
// resident in shared memory — no loads, no epilogue.  bench.py calls it in the same run as the kernel it judges
// (`roofline.peak_measured_this_run`), next to a cuBLAS DGEMM for the fp64 kernel's denominator.
#include <algorithm>

#include "pipeline.cuh"

namespace bopl {

namespace {

__device__ __forceinline__ uint64_t probe_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}

// A [128 x 128] and B [256 x 128] int8 images (K-major core-matrix layout, pseudo-random bytes) in shared memory;
// thread 0 issues `reps` rounds of the four 32-deep k-steps, alternating between two accumulator ranges.
__global__ void __launch_bounds__(128, 1) int8_peak_kernel(int reps) {
  extern __shared__ __align__(1024) unsigned char psm[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int K = 128, N = 256;
  // pseudo-random digits (full int8 range, like the kernel's balanced digits): constant operands would understate the pipe's power
  // draw, and with it overstate the clock the board sustains under its power cap
  for (int i = tid; i < (128 + N) * K / 16; i += 128) {
    uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x ^= x << 13;
      x ^= x >> 17;
      x ^= x << 5;
      w[u] = x;
    }
    reinterpret_cast<uint4*>(psm)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((unsigned long long)smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const unsigned a0 = smem_u32(psm), b0 = a0 + 128 * K, lboA = 128 * 16, lboB = N * 16, sbo = 128;
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = probe_desc(a0 + ks * 2 * lboA, lboA, sbo);
      bd[ks] = probe_desc(b0 + ks * 2 * lboB, lboB, sbo);
    }
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acc = (ks > 1 || r > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (uint32_t)((ks & 1) * N)),
            "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((unsigned long long)smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace

}  // namespace bopl

using namespace bopl;

extern "C" int bopl_probe_int8_peak(bopl_handle* h, double target_ms, double* tops_out, double* ms_out) {
  BOPL_CHECK_HANDLE(h);
  if (!tops_out) return fail(h, BOPL_ERR_INVALID, "null output");
  const size_t smem = (size_t)(128 + 256) * 128;
  BOPL_CUDA(h, cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = h->s_compute;
  cudaEvent_t e0, e1;
  BOPL_CUDA(h, cudaEventCreate(&e0));
  BOPL_CUDA(h, cudaEventCreate(&e1));
  auto run = [&](int reps, float* ms) -> int {
    BOPL_CUDA(h, cudaEventRecord(e0, st));
    int8_peak_kernel<<<h->sms, 128, smem, st>>>(reps);
    BOPL_LAUNCH_CHECK(h, "int8_peak_kernel");
    BOPL_CUDA(h, cudaEventRecord(e1, st));
    BOPL_CUDA(h, cudaEventSynchronize(e1));
    BOPL_CUDA(h, cudaEventElapsedTime(ms, e0, e1));
    return BOPL_OK;
  };
  float ms = 0.f;
  int rc = run(2000, &ms);            // warm-up and calibration
  if (!rc) rc = run(20000, &ms);
  if (!rc) {
    const double want = target_ms > 0.0 ? target_ms : 200.0;
    const int reps = (int)std::min(4.0e6, std::max(20000.0, 20000.0 * want / std::max(1e-3, (double)ms)));
    rc = run(reps, &ms);
    if (!rc) {
      *tops_out = 2.0 * 128.0 * 256.0 * 32.0 * 4.0 * (double)reps * (double)h->sms / ((double)ms * 1e-3) * 1e-12;
      if (ms_out) *ms_out = (double)ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}
