// topk.cu — anchor selection: the k largest acquisition values of a shard, ties -> lowest (global) index.
// Replaces `np.argsort(scores)[:num_anchor]` on scores = -acq: GPyOpt/optimization/anchor_points_generator.py:58-62
// with the sign of GPyOpt/acquisitions/base.py:40.  NaN sorts last, as in np.argsort.
//
// Tournament in levels: every block pulls a 2048-item chunk of (value, index) records into shared memory and
// extracts its k best by k rounds of a block-wide argmax (strict total order: value desc, index asc), writing k
// records; levels repeat until one block is left.  Deterministic regardless of grid size.
#include "common.cuh"

namespace bopl {

namespace {

constexpr int TK_NT = 256;
constexpr int TK_CHUNK = 2048;

__device__ __forceinline__ bool better(double va, long long ia, double vb, long long ib) {
  // NaN is the worst value; among equal values the lower index wins
  const bool na = va != va, nb = vb != vb;
  if (na != nb) return nb;
  if (!na && va != vb) return va > vb;
  return ia < ib;
}

// src_idx == nullptr: level 0, items are acq[i] with index i + index_offset.
__global__ void __launch_bounds__(TK_NT) topk_level_kernel(const double* __restrict__ src_val, const long long* __restrict__ src_idx,
                                                            long long count, int k, long long index_offset,
                                                            double* __restrict__ dst_val, long long* __restrict__ dst_idx) {
  __shared__ double v_s[TK_CHUNK];
  __shared__ long long i_s[TK_CHUNK];
  __shared__ double rv[TK_NT / 32];
  __shared__ long long ri[TK_NT / 32];
  __shared__ int rp[TK_NT / 32];
  const long long base = (long long)blockIdx.x * TK_CHUNK;
  const int cnt = (int)min((long long)TK_CHUNK, count - base);
  for (int t = threadIdx.x; t < TK_CHUNK; t += TK_NT) {
    if (t < cnt) {
      v_s[t] = src_val[base + t];
      i_s[t] = src_idx ? src_idx[base + t] : base + t + index_offset;
    } else {
      v_s[t] = __longlong_as_double(0x7ff8000000000000LL);   // NaN filler: never beats a real record
      i_s[t] = 0x7fffffffffffffffLL;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < k; ++r) {
    double bv = v_s[threadIdx.x];
    long long bi = i_s[threadIdx.x];
    int bp = threadIdx.x;
#pragma unroll
    for (int u = 1; u < TK_CHUNK / TK_NT; ++u) {
      int pos = threadIdx.x + u * TK_NT;
      double v = v_s[pos];
      long long i = i_s[pos];
      if (better(v, i, bv, bi)) {
        bv = v;
        bi = i;
        bp = pos;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double v = __shfl_xor_sync(0xffffffffu, bv, o);
      long long i = __shfl_xor_sync(0xffffffffu, bi, o);
      int pp = __shfl_xor_sync(0xffffffffu, bp, o);
      if (better(v, i, bv, bi)) {
        bv = v;
        bi = i;
        bp = pp;
      }
    }
    if (lane == 0) {
      rv[warp] = bv;
      ri[warp] = bi;
      rp[warp] = bp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < TK_NT / 32; ++w)
        if (better(rv[w], ri[w], bv, bi)) {
          bv = rv[w];
          bi = ri[w];
          bp = rp[w];
        }
      dst_val[(long long)blockIdx.x * k + r] = bv;
      dst_idx[(long long)blockIdx.x * k + r] = bi;
      v_s[bp] = __longlong_as_double(0x7ff8000000000000LL);   // retire the winner
      i_s[bp] = 0x7fffffffffffffffLL;
    }
    __syncthreads();
  }
}

}  // namespace

int launch_topk(bopl_handle* h, const double* acq, int N, int k, long long index_offset, double* val, long long* idx,
                cudaStream_t st) {
  if (k <= 0) return BOPL_OK;
  if (k > MAX_TOPK) return fail(h, BOPL_ERR_UNSUPPORTED, "topk: k > 256");
  if (N < k) return fail(h, BOPL_ERR_INVALID, "topk: fewer candidates than k");
  long long count = N;
  long long blocks = (count + TK_CHUNK - 1) / TK_CHUNK;
  // ping-pong scratch sized for level 0's output
  size_t need = (size_t)2 * blocks * k;
  if (need > h->topk_cap) {
    if (h->topk_val_s) cudaFree(h->topk_val_s);
    if (h->topk_idx_s) cudaFree(h->topk_idx_s);
    h->topk_val_s = nullptr;
    h->topk_idx_s = nullptr;
    h->topk_cap = 0;
    BOPL_CUDA(h, cudaMalloc((void**)&h->topk_val_s, need * sizeof(double)));
    BOPL_CUDA(h, cudaMalloc((void**)&h->topk_idx_s, need * sizeof(long long)));
    h->topk_cap = need;
  }
  const double* sv = acq;
  const long long* si = nullptr;
  int side = 0;
  while (true) {
    double* dv = (blocks == 1) ? val : h->topk_val_s + (size_t)side * (h->topk_cap / 2);
    long long* di = (blocks == 1) ? idx : h->topk_idx_s + (size_t)side * (h->topk_cap / 2);
    topk_level_kernel<<<(unsigned)blocks, TK_NT, 0, st>>>(sv, si, count, k, index_offset, dv, di);
    BOPL_LAUNCH_CHECK(h, "topk_level_kernel");
    if (blocks == 1) break;
    sv = dv;
    si = di;
    count = blocks * k;
    blocks = (count + TK_CHUNK - 1) / TK_CHUNK;
    side ^= 1;
  }
  return BOPL_OK;
}

}  // namespace bopl
