// collective.cu — the one collective of the path: all ranks exchange their k best candidates and run the same merge.
// Candidates shard row-wise over the GPUs of one box; the MC / theta reductions are per candidate, so nothing else travels.
// The reference has no counterpart (its only parallelism is a 4-process pool, uEI.py:91); what this replaces is the single
// `np.argsort(scores)[:k]` over the whole batch in GPyOpt/optimization/anchor_points_generator.py:58-62.
//
//   topk_pack_kernel    k records (value, global index, x[d]) of this rank, fillers (NaN, INT64_MAX) beyond its candidate count
//   ncclAllGather       k (2 + d) 8-byte words per rank on the caller's stream (1.9 KB at k = 20, d = 10) over NVLink / NVSwitch
//   topk_merge_kernel   identical deterministic merge on every rank: value descending, ties -> lowest global index, NaN last
//
// No host synchronisation anywhere in bopl_topk_allgather.  NCCL is bound at run time (dlopen of libnccl.so.2 — inside a PyTorch
// process that is the copy torch already loaded), so the library itself has no link-time dependency on it and single-GPU users never
// touch it.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace bopl {

namespace {

// the five NCCL entry points used, with the ABI of nccl.h 2.x (ncclUniqueId is 128 opaque bytes, ncclFloat64 == 8)
struct NcclUniqueId {
  char internal[128];
};
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_comm_count)(void*, int*);
typedef int (*fn_comm_user_rank)(void*, int*);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);
constexpr int NCCL_FLOAT64 = 8;

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_comm_count comm_count = nullptr;
  fn_comm_user_rank comm_user_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_get_error_string get_error_string = nullptr;
  std::string err;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "") + ")";
      return;
    }
    api.get_unique_id = (fn_get_unique_id)dlsym(api.lib, "ncclGetUniqueId");
    api.comm_init_rank = (fn_comm_init_rank)dlsym(api.lib, "ncclCommInitRank");
    api.comm_destroy = (fn_comm_destroy)dlsym(api.lib, "ncclCommDestroy");
    api.comm_count = (fn_comm_count)dlsym(api.lib, "ncclCommCount");
    api.comm_user_rank = (fn_comm_user_rank)dlsym(api.lib, "ncclCommUserRank");
    api.all_gather = (fn_all_gather)dlsym(api.lib, "ncclAllGather");
    api.get_error_string = (fn_get_error_string)dlsym(api.lib, "ncclGetErrorString");
    if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.comm_count || !api.comm_user_rank || !api.all_gather ||
        !api.get_error_string)
      api.err = "libnccl.so.2 lacks an expected symbol";
  });
  return api;
}

int nccl_fail(bopl_handle* h, const char* what, int code) {
  NcclApi& a = nccl_api();
  return fail(h, BOPL_ERR_CUDA, std::string(what) + ": " + (a.get_error_string ? a.get_error_string(code) : "NCCL error"));
}

constexpr long long IDX_FILL = 0x7fffffffffffffffLL;

__device__ __forceinline__ bool rec_better(double va, long long ia, double vb, long long ib) {   // the order of topk.cu
  const bool na = va != va, nb = vb != vb;
  if (na != nb) return nb;
  if (!na && va != vb) return va > vb;
  return ia < ib;
}

// rec [k][2 + d]: value, global index (bit pattern), x.  Records r >= k_local are fillers.
__global__ void topk_pack_kernel(const double* __restrict__ val, const long long* __restrict__ idx, const double* __restrict__ Xc,
                                 long long index_offset, int k_local, int k, int d, double* __restrict__ rec) {
  const int W = 2 + d;
  for (int e = threadIdx.x; e < k * W; e += blockDim.x) {
    const int r = e / W, c = e - r * W;
    double out;
    if (r >= k_local) {
      out = c == 0 ? __longlong_as_double(0x7ff8000000000000LL) : (c == 1 ? __longlong_as_double(IDX_FILL) : 0.0);
    } else if (c == 0) {
      out = val[r];
    } else if (c == 1) {
      out = __longlong_as_double(idx[r]);
    } else {
      out = Xc[(size_t)(idx[r] - index_offset) * d + (c - 2)];
    }
    rec[e] = out;
  }
}

// all [R][2 + d] records of every rank -> the k best, in order.  One block; R <= 64 * 256.
__global__ void __launch_bounds__(256) topk_merge_kernel(const double* __restrict__ all, int R, int k, int d, double* __restrict__ val,
                                                          long long* __restrict__ idx, double* __restrict__ x) {
  extern __shared__ unsigned char taken[];
  __shared__ double rv[8];
  __shared__ long long ri[8];
  __shared__ int rp[8];
  const int W = 2 + d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < R; e += 256) taken[e] = 0;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    double bv = __longlong_as_double(0x7ff8000000000000LL);
    long long bi = IDX_FILL;
    int bp = -1;
    for (int e = threadIdx.x; e < R; e += 256) {
      if (taken[e]) continue;
      const double v = all[(size_t)e * W];
      const long long i = __double_as_longlong(all[(size_t)e * W + 1]);
      if (bp < 0 || rec_better(v, i, bv, bi)) {
        bv = v;
        bi = i;
        bp = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long i = __shfl_xor_sync(0xffffffffu, bi, o);
      const int pp = __shfl_xor_sync(0xffffffffu, bp, o);
      if (pp >= 0 && (bp < 0 || rec_better(v, i, bv, bi))) {
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
      for (int w = 1; w < 8; ++w)
        if (rp[w] >= 0 && (bp < 0 || rec_better(rv[w], ri[w], bv, bi))) {
          bv = rv[w];
          bi = ri[w];
          bp = rp[w];
        }
      rp[0] = bp;
      val[r] = bv;
      idx[r] = bi;
      if (bp >= 0) taken[bp] = 1;
    }
    __syncthreads();
    const int win = rp[0];
    if (x)
      for (int q = threadIdx.x; q < d; q += 256) x[(size_t)r * d + q] = win >= 0 ? all[(size_t)win * W + 2 + q] : 0.0;
    __syncthreads();
  }
}

}  // namespace

int launch_topk_pack(bopl_handle* h, const double* val, const long long* idx, const double* Xc, long long index_offset, int k_local,
                     int k, int d, double* rec, cudaStream_t st) {
  topk_pack_kernel<<<1, 256, 0, st>>>(val, idx, Xc, index_offset, k_local, k, d, rec);
  BOPL_LAUNCH_CHECK(h, "topk_pack_kernel");
  return BOPL_OK;
}

int launch_topk_merge(bopl_handle* h, const double* all, int R, int k, int d, double* val, long long* idx, double* x, cudaStream_t st) {
  if (R > 64 * MAX_TOPK) return fail(h, BOPL_ERR_UNSUPPORTED, "topk merge: more than 64 ranks x 256 records");
  topk_merge_kernel<<<1, 256, (size_t)R, st>>>(all, R, k, d, val, idx, x);
  BOPL_LAUNCH_CHECK(h, "topk_merge_kernel");
  return BOPL_OK;
}

}  // namespace bopl

using namespace bopl;

extern "C" {

int bopl_nccl_unique_id(void* id128_h) {
  if (!id128_h) return fail(nullptr, BOPL_ERR_INVALID, "null id buffer");
  NcclApi& a = nccl_api();
  if (!a.err.empty()) return fail(nullptr, BOPL_ERR_UNSUPPORTED, a.err);
  NcclUniqueId id;
  const int rc = a.get_unique_id(&id);
  if (rc) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
  memcpy(id128_h, id.internal, sizeof(id.internal));
  return BOPL_OK;
}

int bopl_nccl_comm_init(bopl_handle* h, int nranks, int rank, const void* id128_h, void** comm_out) {
  BOPL_CHECK_HANDLE(h);
  if (!id128_h || !comm_out || nranks <= 0 || rank < 0 || rank >= nranks) return fail(h, BOPL_ERR_INVALID, "bad nranks / rank / id");
  NcclApi& a = nccl_api();
  if (!a.err.empty()) return fail(h, BOPL_ERR_UNSUPPORTED, a.err);
  NcclUniqueId id;
  memcpy(id.internal, id128_h, sizeof(id.internal));
  void* comm = nullptr;
  const int rc = a.comm_init_rank(&comm, nranks, id, rank);
  if (rc) return nccl_fail(h, "ncclCommInitRank", rc);
  *comm_out = comm;
  return BOPL_OK;
}

int bopl_nccl_comm_destroy(void* comm) {
  if (!comm) return BOPL_OK;
  NcclApi& a = nccl_api();
  if (!a.err.empty()) return fail(nullptr, BOPL_ERR_UNSUPPORTED, a.err);
  const int rc = a.comm_destroy(comm);
  return rc ? nccl_fail(nullptr, "ncclCommDestroy", rc) : BOPL_OK;
}

int bopl_topk_allgather(bopl_handle* h, void* comm, const double* val, const long long* idx, int k_local, const double* Xc,
                        long long index_offset, int k, double* val_all, long long* idx_all, double* x_all, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (!comm) return fail(h, BOPL_ERR_INVALID, "null communicator");
  if (k <= 0 || k > MAX_TOPK || k_local < 0 || k_local > k) return fail(h, BOPL_ERR_INVALID, "bad k / k_local");
  if (!val_all || !idx_all || (k_local && (!val || !idx || !Xc))) return fail(h, BOPL_ERR_INVALID, "null buffer");
  NcclApi& a = nccl_api();
  if (!a.err.empty()) return fail(h, BOPL_ERR_UNSUPPORTED, a.err);
  int nranks = 0, rc;
  if ((rc = a.comm_count(comm, &nranks))) return nccl_fail(h, "ncclCommCount", rc);
  const int d = h->d, W = 2 + d;
  const size_t per = (size_t)k * W;
  if ((rc = ensure_bytes(h, (void**)&h->coll_s, &h->coll_bytes, per * (size_t)(nranks + 1) * sizeof(double)))) return rc;
  double* send = h->coll_s;
  double* recv = h->coll_s + per;
  cudaStream_t st = (cudaStream_t)stream;
  StreamOrder order_guard(h, st);
  if ((rc = launch_topk_pack(h, val, idx, Xc, index_offset, k_local, k, d, send, st))) return rc;
  const int nrc = a.all_gather(send, recv, per, NCCL_FLOAT64, comm, st);
  if (nrc) return nccl_fail(h, "ncclAllGather", nrc);
  return launch_topk_merge(h, recv, nranks * k, k, d, val_all, idx_all, x_all, st);
}

int bopl_topk_allgather_host(bopl_handle* h, void* comm, const double* val_h, const long long* idx_h, int k_local, const double* x_h,
                             int k, double* val_all_h, long long* idx_all_h, double* x_all_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (!comm) return fail(h, BOPL_ERR_INVALID, "null communicator");
  if (k <= 0 || k > MAX_TOPK || k_local < 0 || k_local > k) return fail(h, BOPL_ERR_INVALID, "bad k / k_local");
  if (!val_all_h || !idx_all_h || !x_all_h || (k_local && (!val_h || !idx_h || !x_h))) return fail(h, BOPL_ERR_INVALID, "null buffer");
  NcclApi& a = nccl_api();
  if (!a.err.empty()) return fail(h, BOPL_ERR_UNSUPPORTED, a.err);
  int nranks = 0, rc;
  if ((rc = a.comm_count(comm, &nranks))) return nccl_fail(h, "ncclCommCount", rc);
  const int d = h->d, W = 2 + d;
  const size_t per = (size_t)k * W, out_cnt = (size_t)k * W;     // outputs: val [k] | idx [k] | x [k*d]
  if ((rc = ensure_bytes(h, (void**)&h->coll_s, &h->coll_bytes, (per * (size_t)(nranks + 1) + out_cnt) * sizeof(double)))) return rc;
  double* send = h->coll_s;
  double* recv = send + per;
  double* outv = recv + per * nranks;
  long long* outi = (long long*)(outv + k);
  double* outx = outv + 2 * (size_t)k;
  std::vector<double> rec(per);
  for (int r = 0; r < k; ++r) {
    double* o = rec.data() + (size_t)r * W;
    if (r < k_local) {
      o[0] = val_h[r];
      memcpy(o + 1, idx_h + r, sizeof(long long));
      memcpy(o + 2, x_h + (size_t)r * d, (size_t)d * sizeof(double));
    } else {
      const long long nanb = 0x7ff8000000000000LL, fill = IDX_FILL;
      memcpy(o, &nanb, 8);
      memcpy(o + 1, &fill, 8);
      for (int q = 0; q < d; ++q) o[2 + q] = 0.0;
    }
  }
  cudaStream_t st = h->s_compute;
  StreamOrder order_guard(h, st);
  BOPL_CUDA(h, cudaMemcpyAsync(send, rec.data(), per * sizeof(double), cudaMemcpyHostToDevice, st));
  const int nrc = a.all_gather(send, recv, per, NCCL_FLOAT64, comm, st);
  if (nrc) return nccl_fail(h, "ncclAllGather", nrc);
  if ((rc = launch_topk_merge(h, recv, nranks * k, k, d, outv, outi, outx, st))) return rc;
  std::vector<double> out(out_cnt);
  BOPL_CUDA(h, cudaMemcpyAsync(out.data(), outv, out_cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaStreamSynchronize(st));      // also keeps `rec` alive until the upload has been consumed
  memcpy(val_all_h, out.data(), (size_t)k * sizeof(double));
  memcpy(idx_all_h, out.data() + k, (size_t)k * sizeof(long long));
  memcpy(x_all_h, out.data() + 2 * (size_t)k, (size_t)k * d * sizeof(double));
  return BOPL_OK;
}

int bopl_topk_merge(bopl_handle* h, const double* records, int R, int k, double* val_all, long long* idx_all, double* x_all,
                    void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (R <= 0 || k <= 0 || k > MAX_TOPK || !records || !val_all || !idx_all) return fail(h, BOPL_ERR_INVALID, "bad R / k / buffer");
  return launch_topk_merge(h, records, R, k, h->d, val_all, idx_all, x_all, (cudaStream_t)stream);
}

int bopl_topk_pack(bopl_handle* h, const double* val, const long long* idx, int k_local, const double* Xc, long long index_offset,
                   int k, double* records, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (k <= 0 || k > MAX_TOPK || k_local < 0 || k_local > k || !records) return fail(h, BOPL_ERR_INVALID, "bad k / k_local / buffer");
  if (k_local && (!val || !idx || !Xc)) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return launch_topk_pack(h, val, idx, Xc, index_offset, k_local, k, h->d, records, (cudaStream_t)stream);
}

}  // extern "C"
