// fp64_vs_umma_probe.cu — do the FP64 pipe (DFMA / DMMA) and the int8 tensor pipe (tcgen05.mma kind::i8) run concurrently on a B200 SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fvu tools/fp64_vs_umma_probe.cu && /tmp/fvu
// One CTA per SM: warp 0 issues back-to-back N = 256 int8 MMAs on resident operands, warps 1..8 run eight independent DFMA (or DMMA)
// chains each.  Three runs: FP64 alone, MMA alone, both; every role times itself with clock64.  Prints lanes/clk/SM and MAC/clk/SM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nW_L:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra W_D;\nbra W_L;\nW_D:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(288, 1) probe(int mma_reps, int fp_iters, int fp_mode, long long* out, long long* perwarp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 4; i += 288) { uint32_t x = (uint32_t)i * 2654435761u + 12345u; x ^= x << 13; x ^= x >> 17; x ^= x << 5; ((uint32_t*)smem)[i] = x; }   // pseudo-random operand bytes
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
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
  if (warp == 0) {
    if (tid == 0 && mma_reps > 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const unsigned a0 = smem_u32(smem), b0 = a0 + 128 * 128;
      uint64_t ad[4], bd[4];
      for (int ks = 0; ks < 4; ++ks) { ad[ks] = make_desc(a0 + ks * 2 * 2048, 2048, 128); bd[ks] = make_desc(b0 + ks * 2 * 4096, 4096, 128); }
      const long long t0 = clock64();
      for (int r = 0; r < mma_reps; ++r)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_i8(tmem, ad[ks], bd[ks], idesc, 1u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((unsigned long long)smem_u32(&bar)) : "memory");
      mbar_wait(&bar, 0);
      out[blockIdx.x * 2 + 0] = clock64() - t0;
    }
  } else if (fp_iters > 0) {
    double a[8], x = 1.0 + 1e-9 * tid, y = 1e-9 * warp;
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = u;
    const long long t0 = clock64();
    if (fp_mode == 0) {
      for (int it = 0; it < fp_iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = fma(a[u], x, y);
    } else if (fp_mode == 1) {
      for (int it = 0; it < fp_iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(a[2 * u]), "+d"(a[2 * u + 1]) : "d"(x), "d"(y));
    } else if (fp_mode == 2) {      // FP32 pipe: eight FFMA chains (round 2: does the FP32 pipe time-share with the int8 MMA as the FP64 pipe does?)
      float f[8], fx = 1.0f + 1e-7f * tid, fy = 1e-7f * warp;
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = (float)u;
      for (int it = 0; it < fp_iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = fmaf(f[u], fx, fy);
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = f[u];
    } else {                        // integer pipe: eight IMAD chains
      int q[8], qx = 3 + tid, qy = warp;
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = u;
      for (int it = 0; it < fp_iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = q[u] * qx + qy;
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = q[u];
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u];
    if (s == 123.456) out[0] = 0;
    if (tid == 32) out[blockIdx.x * 2 + 1] = t1 - t0;
    if ((tid & 31) == 0 && blockIdx.x == 0 && perwarp) perwarp[warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}
int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  CK(cudaMalloc(&d, sms * 16 + 16 * 8));
  const size_t sm = (128 + 256) * 128 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  std::vector<long long> h(sms * 2);
  printf("{\"sms\": %d", sms);
  const int reps = 20000, iters = 60000;
  for (int fp_mode = 0; fp_mode < 4; ++fp_mode) {
    const int it = fp_mode == 1 ? iters / 2 : iters;
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int mr = (cfg == 0) ? 0 : reps, fi = (cfg == 1) ? 0 : it;
      if (fp_mode >= 1 && cfg == 1) continue;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(d, 0, sms * 16 + 16 * 8));
        probe<<<sms, 288, sm>>>(mr, fi, fp_mode, d, d + sms * 2);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), d, sms * 16, cudaMemcpyDeviceToHost));
      double cm = 0, cf = 0;
      for (int s = 0; s < sms; ++s) { cm += h[2 * s]; cf += h[2 * s + 1]; }
      cm /= sms; cf /= sms;
      const char* nm = fp_mode == 0 ? "dfma" : fp_mode == 1 ? "dmma" : fp_mode == 2 ? "ffma" : "imad";
      if (mr) printf(", \"%s_cfg%d_mma_mac_per_clk\": %.0f", nm, cfg, 4.0 * reps * 128 * 256 * 32 / cm);
      if (fi) printf(", \"%s_cfg%d_%s_lanes_per_clk\": %.1f", nm, cfg, fp_mode < 2 ? "fp64_fma" : "fma", fp_mode == 1 ? 8.0 * 4 * 256 * fi / cf : 8.0 * 32 * 8 * fi / cf);
      if (fi) {   // clocks of the eight FP64 warps of CTA 0 (warp w sits on sub-partition w % 4; warp 0 issues the MMAs)
        long long pw[16];
        CK(cudaMemcpy(pw, d + sms * 2, sizeof(pw), cudaMemcpyDeviceToHost));
        printf(", \"%s_cfg%d_clocks_warp1to8\": [%lld, %lld, %lld, %lld, %lld, %lld, %lld, %lld]", nm, cfg, pw[1], pw[2], pw[3], pw[4], pw[5], pw[6], pw[7], pw[8]);
      }
    }
  }
  printf("}\n");
  return 0;
}
