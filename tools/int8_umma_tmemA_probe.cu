// int8_umma_tmemA_probe.cu — does taking the A operand of tcgen05.mma kind::i8 from TMEM (filled once by tcgen05.cp) lift the
// ~90-clock floor of the N <= 128 instructions?  Same set-up as int8_umma_probe.cu (M = 128, K = 128 as four K = 32 slabs, operands
// resident, every SM issues back-to-back), A either from shared memory (descriptor) or from TMEM columns 384..415.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/int8_tmemA tools/int8_umma_tmemA_probe.cu && build/tools/int8_tmemA
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__host__ __device__ inline size_t canon(int r, int k, int R) { return ((size_t)(k / 16) * (R / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 16); }
__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nW_L:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra W_D;\nbra W_L;\nW_D:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int a_from_tmem>
__global__ void __launch_bounds__(128, 1) probe(const int8_t* Aimg, const int8_t* Bimg, int N, int reps, int32_t* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int K = 128, tid = threadIdx.x, warp = tid >> 5;
  unsigned char* As = smem;
  unsigned char* Bs = smem + (size_t)128 * K;
  for (int i = tid; i < 128 * K / 16; i += 128) ((uint4*)As)[i] = ((const uint4*)Aimg)[i];
  for (int i = tid; i < N * K / 16; i += 128) ((uint4*)Bs)[i] = ((const uint4*)Bimg)[i];
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
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const unsigned a0 = smem_u32(As), b0 = smem_u32(Bs), lboA = 128 * 16, lboB = (unsigned)N * 16, sbo = 128;
  if (tid == 0) {
    uint64_t ad[4], bd[4];
    uint32_t at[4];
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = make_desc(a0 + ks * 2 * lboA, lboA, sbo);
      bd[ks] = make_desc(b0 + ks * 2 * lboB, lboB, sbo);
      at[ks] = tmem + 384 + ks * 8;
    }
    if (a_from_tmem)
      for (int ks = 0; ks < 4; ++ks)      // one K = 32 slab (128 rows x 32 bytes) per copy
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(at[ks]), "l"(ad[ks]) : "memory");
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (a_from_tmem) mma_i8_ts(tmem, at[ks], bd[ks], idesc, (ks > 0 || r > 0) ? 1u : 0u);
        else mma_i8_ss(tmem, ad[ks], bd[ks], idesc, (ks > 0 || r > 0) ? 1u : 0u);
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((unsigned long long)smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (out) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int c = 0; c < 8; ++c) out[(size_t)tid * N + c0 + c] = (int32_t)v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}
int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("{\"sms\": %d", sms);
  srand(1);
  for (int N : {64, 128, 256}) {
    const int K = 128;
    std::vector<int8_t> A(128 * K), B((size_t)N * K), Ai(128 * K), Bi((size_t)N * K);
    for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) Ai[canon(r, k, 128)] = A[r * K + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) Bi[canon(r, k, N)] = B[(size_t)r * K + k];
    int8_t *dA, *dB; int32_t* dO;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, Ai.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bi.data(), B.size(), cudaMemcpyHostToDevice));
    const size_t sm = (size_t)(128 + N) * K + 1024;
    CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int mode = 0; mode < 2; ++mode) {
      CK(cudaMemset(dO, 0, (size_t)128 * N * 4));
      if (mode) probe<1><<<1, 128, sm>>>(dA, dB, N, 1, dO); else probe<0><<<1, 128, sm>>>(dA, dB, N, 1, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(", \"error_N%d_mode%d\": \"%s\"}\n", N, mode, cudaGetErrorString(e)); return 1; }
      std::vector<int32_t> O((size_t)128 * N);
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      long bad = 0;
      for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
        int32_t s = 0;
        for (int k = 0; k < K; ++k) s += (int32_t)A[i * K + k] * (int32_t)B[(size_t)j * K + k];
        if (s != O[(size_t)i * N + j]) ++bad;
      }
      printf(", \"mismatch_N%d_%s\": %ld", N, mode ? "tmemA" : "smemA", bad);
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      const int reps = 4000;
      if (mode) probe<1><<<sms, 128, sm>>>(dA, dB, N, 64, nullptr); else probe<0><<<sms, 128, sm>>>(dA, dB, N, 64, nullptr); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); if (mode) probe<1><<<sms, 128, sm>>>(dA, dB, N, reps, nullptr); else probe<0><<<sms, 128, sm>>>(dA, dB, N, reps, nullptr); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf(", \"tops_N%d_%s\": %.1f", N, mode ? "tmemA" : "smemA", 2.0 * 128 * N * K * (double)reps * sms / (ms * 1e-3) / 1e12);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  printf("}\n");
  return 0;
}
