// int8_umma_probe.cu — is the INT8 tensor pipe (tcgen05.mma kind::i8, TMEM accumulators) usable for an Ozaki-style split of the
// fp64 triangular-solve update, and how fast is it on this B200?   Stand-alone; built and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/int8_probe tools/int8_umma_probe.cu && /tmp/int8_probe
// (1) correctness: D[128 x N] (s32, TMEM) = A[128 x K] (s8) * B[N x K]^T (s8), operands in shared memory in the canonical
//     K-major no-swizzle layout (8-row x 16-byte core matrices), against the host's integer product;
// (2) rate: every SM issues back-to-back MMAs on resident operands; prints int8 TOP/s for N = 64 / 128 / 256.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
// offset(r, k) = ((k / 16) * (R / 8) + r / 8) * 128 + (r % 8) * 16 + k % 16   ->  SBO (next 8 rows) = 128, LBO (next 16 k) = R * 16
__host__ __device__ inline size_t canon(int r, int k, int R) { return ((size_t)(k / 16) * (R / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 16); }

__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  return d;                // layout type 0: no swizzle
}

__host__ __device__ inline uint32_t make_idesc_s8(int M, int N) {
  // c_format S32 = 2 @ [4,6); a_format, b_format INT8 = 1 @ [7,10), [10,13); K-major both; n_dim = N >> 3 @ [17,23); m_dim = M >> 4 @ [24,29)
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"((unsigned long long)smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// One CTA of 128 threads.  A image [128 x K], B image [N x K] (canonical layout, K a multiple of 32).  reps > 1: rate mode.
__global__ void __launch_bounds__(128, 1) umma_kernel(const int8_t* Aimg, const int8_t* Bimg, int N, int K, int reps, int sync_every, int32_t* out, int rotate = 1) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* As = smem;
  unsigned char* Bs = smem + (size_t)128 * K;
  for (int i = tid; i < 128 * K / 16; i += 128) ((uint4*)As)[i] = ((const uint4*)Aimg)[i];
  for (int i = tid; i < N * K / 16; i += 128) ((uint4*)Bs)[i] = ((const uint4*)Bimg)[i];
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"l"((unsigned long long)smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_s8(128, N);
  const unsigned a0 = smem_u32(As), b0 = smem_u32(Bs);
  const unsigned lboA = 128 * 16, lboB = (unsigned)N * 16, sbo = 128;
  if (tid == 0) {
    // descriptors of the K/32 steps are loop-invariant: build them once (K <= 128 here), the loop is MMA issue only
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = make_desc(a0 + ks * 2 * lboA, lboA, sbo);
      bd[ks] = make_desc(b0 + ks * 2 * lboB, lboB, sbo);
    }
    const int nks = K / 32;
    uint32_t toff[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) toff[ks] = (uint32_t)((ks % rotate) * N);   // rotate divides 4: the offset depends on ks only
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        if (ks < nks) mma_i8(tmem + toff[ks], ad[ks], bd[ks], idesc, (ks > 0 || r > 0) ? 1u : 0u);   // rate mode keeps accumulating (s32 wraps); rotate > 1: consecutive MMAs hit different accumulators
      if (sync_every > 0 && (r + 1) % sync_every == 0 && r + 1 < reps) {   // optional rendezvous (drains the MMA pipe)
        mma_commit(&bar[1]);
        mbar_wait(&bar[1], (((r + 1) / sync_every - 1) & 1));
      }
    }
    mma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (out) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int c = 0; c < 16; ++c) out[(size_t)blockIdx.x * 128 * N + (size_t)tid * N + c0 + c] = (int32_t)v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("{\"sms\": %d", sms);
  srand(1);
  bool all_ok = true;
  for (int N : {64, 128, 256}) {
    const int K = 128;
    std::vector<int8_t> A(128 * K), B((size_t)N * K), Ai(128 * K), Bi((size_t)N * K);
    for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < K; ++k) Ai[canon(r, k, 128)] = A[r * K + k];
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < K; ++k) Bi[canon(r, k, N)] = B[(size_t)r * K + k];
    int8_t *dA, *dB;
    int32_t* dO;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dO, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, Ai.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), B.size(), cudaMemcpyHostToDevice));
    const size_t sm = (size_t)(128 + N) * K + 1024;
    CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    umma_kernel<<<1, 128, sm>>>(dA, dB, N, K, 1, 0, dO);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> O((size_t)128 * N);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < N; ++j) {
        int32_t s = 0;
        for (int k = 0; k < K; ++k) s += (int32_t)A[i * K + k] * (int32_t)B[(size_t)j * K + k];
        if (s != O[(size_t)i * N + j]) ++bad;
      }
    printf(", \"mismatch_N%d\": %ld", N, bad);
    if (bad) {
      all_ok = false;
      printf(", \"first_rows_N%d\": [%d, %d, %d, %d]", N, O[0], O[1], O[N], O[N + 1]);
    }
    // rate: every SM, reps batches of K/32 MMAs; sync_every = 0: never drained, the issuing thread is back-pressured by the hardware queue
    const int reps = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int sync_every : {0, 64, 8}) {
      umma_kernel<<<sms, 128, sm>>>(dA, dB, N, K, 64, sync_every, nullptr);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      umma_kernel<<<sms, 128, sm>>>(dA, dB, N, K, reps, sync_every, nullptr);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * 128 * N * K * (double)reps * sms;
      printf(", \"int8_tops_N%d_sync%d\": %.1f", N, sync_every, ops / (ms * 1e-3) / 1e12);
    }
    // consecutive instructions on DIFFERENT accumulator column ranges (independent): is the small-N rate a dependent-accumulate latency?
    for (int rotate : {2, 4}) {
      if (N * rotate > 512) continue;
      umma_kernel<<<sms, 128, sm>>>(dA, dB, N, K, 64, 0, nullptr, rotate);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      umma_kernel<<<sms, 128, sm>>>(dA, dB, N, K, reps, 0, nullptr, rotate);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = 2.0 * 128 * N * K * (double)reps * sms;
      printf(", \"int8_tops_N%d_rotate%d\": %.1f", N, rotate, ops / (ms * 1e-3) / 1e12);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dO);
  }
  printf(", \"ok\": %s}\n", all_ok ? "true" : "false");
  return 0;
}
