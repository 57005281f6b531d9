// api.cu — the C-ABI of libbopl_b200.so (include/bopl_b200.h): handle lifetime, model upload and the entry points
// that sequence the kernels.  Plain pointers and sizes only; every entry returns a bopl_status.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>
#include <cstdlib>

#include "common.cuh"

namespace bopl {

thread_local std::string g_err;

int fail(bopl_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_err = msg;
  return code;
}

int ensure_bytes(bopl_handle* h, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return BOPL_OK;
  if (*p) {
    BOPL_CUDA(h, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
  }
  BOPL_CUDA(h, cudaMalloc(p, bytes));
  *cap = bytes;
  return BOPL_OK;
}

void free_model(bopl_handle* h) {
  path_free(h);
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->gp_host.clear();
  h->gp_dev = nullptr;
  h->X_dev = nullptr;
  h->has_model = false;
  h->has_winv = false;
  h->has_linv = false;
  h->has_oz = false;
  h->F_valid = false;
}

namespace {

// Synchronous upload from (pageable) host memory on the legacy stream.  cudaMemcpy may return once the data sits in the driver's staging
// buffer, and the library's own streams are non-blocking (no implicit ordering with the legacy stream): every caller finishes its
// uploads with uploads_done() before a kernel on s_compute / s_copy / the caller's stream may read them.
template <typename T>
int dev_upload(bopl_handle* h, const T* src, size_t count, T** out) {
  void* p = nullptr;
  BOPL_CUDA(h, cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  h->owned.push_back(p);
  if (count) BOPL_CUDA(h, cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *out = (T*)p;
  return BOPL_OK;
}

// Value + gradient entry points: the small-batch posterior and its gradient share one K* / dK generation (gp_kernels.cu).
struct GradFollows {
  bopl_handle* h;
  GradFollows(bopl_handle* h_, bool on) : h(h_) { h->small_grad_follows = on; }
  ~GradFollows() { h->small_grad_follows = false; }
};

// BOPL_POSTERIOR_INT8_AUTO: does the 6-digit split meet the path's bound for THIS model?  The worst place for an absolute error is where
// the variance is smallest — at the training points — so the probe runs (up to) 256 of them through the 6-digit kernel and through the fp64
// DMMA kernel (noiseless variance, no clip, every hyper-sample) and asks for |v6 - v64| <= 1e-9 |v64| + 64 ulp(sigma^2) on every one.
// Returns 1 (keep 6 digits), 0 (use 7), or a negative status.
int int8_probe_passes(bopl_handle* h) {
  const int n = h->n, d = h->d, m = h->m;
  const int N = std::max(SMALL_N + 1, std::min(n, 256));
  std::vector<double> Xh((size_t)n * d), Xp((size_t)N * d);
  BOPL_CUDA(h, cudaMemcpy(Xh.data(), h->X_dev, Xh.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int c = 0; c < N; ++c) {
    const int i = (int)(((long long)c * n) / N) % n;                 // spread over the training set (repeats when n < 33)
    memcpy(Xp.data() + (size_t)c * d, Xh.data() + (size_t)i * d, (size_t)d * sizeof(double));
  }
  double* buf = nullptr;
  BOPL_CUDA(h, cudaMalloc((void**)&buf, ((size_t)N * d + 2 * (size_t)m * N) * sizeof(double)));
  double *Xd = buf, *v6 = buf + (size_t)N * d, *v64 = v6 + (size_t)m * N;
  std::vector<double> a((size_t)m * N), b((size_t)m * N);
  int ok = 1, rc = BOPL_OK;
  cudaStream_t st = h->s_compute;
  cudaError_t e = cudaMemcpy(Xd, Xp.data(), Xp.size() * sizeof(double), cudaMemcpyHostToDevice);
  for (int hh = 0; hh < h->H && ok == 1 && e == cudaSuccess && rc == BOPL_OK; ++hh) {
    if ((rc = launch_posterior_ozaki(h, Xd, N, hh, nullptr, v6, 0, st)) || (rc = launch_posterior_trsm_t(h, Xd, N, hh, nullptr, v64, 0, st))) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    if ((e = cudaMemcpy(a.data(), v6, a.size() * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    if ((e = cudaMemcpy(b.data(), v64, b.size() * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    for (int j = 0; j < m && ok == 1; ++j) {
      const double s2 = h->gp_host[(size_t)j * h->H + hh].variance;
      for (int c = 0; c < N; ++c) {
        const double x = a[(size_t)j * N + c], y = b[(size_t)j * N + c];
        if (!(std::fabs(x - y) <= 1e-9 * std::fabs(y) + 64.0 * 2.220446049250313e-16 * s2)) {
          ok = 0;
          break;
        }
      }
    }
  }
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, BOPL_ERR_CUDA, std::string("int8 probe: ") + cudaGetErrorString(e));
  return ok;
}

int uploads_done(bopl_handle* h) {
  BOPL_CUDA(h, cudaDeviceSynchronize());
  return BOPL_OK;
}

// Inverse of a lower-triangular B x B block (row-major, leading dimension ld) in extended precision.
// Returns false on a non-positive diagonal.
bool invert_lower_block(const double* D, int ld, int B, double* out) {
  std::vector<long double> x((size_t)B * B, 0.0L);
  for (int c = 0; c < B; ++c) {
    long double dcc = D[(size_t)c * ld + c];
    if (!(dcc > 0.0L)) return false;
    x[(size_t)c * B + c] = 1.0L / dcc;
    for (int r = c + 1; r < B; ++r) {
      long double s = 0.0L;
      for (int k = c; k < r; ++k) s += (long double)D[(size_t)r * ld + k] * x[(size_t)k * B + c];
      long double drr = D[(size_t)r * ld + r];
      if (!(drr > 0.0L)) return false;
      x[(size_t)r * B + c] = -s / drr;
    }
  }
  for (size_t i = 0; i < (size_t)B * B; ++i) out[i] = (double)x[i];
  return true;
}

int check_utility(bopl_handle* h, int util_kind, int Lt, int p) {
  if (Lt <= 0) return fail(h, BOPL_ERR_INVALID, "Lt must be positive");
  if (util_kind == BOPL_UTIL_LINEAR || util_kind == BOPL_UTIL_QUADRATIC) {
    if (p != h->m) return fail(h, BOPL_ERR_INVALID, "linear/quadratic utility needs p == m parameters per sample");
  } else if (util_kind == BOPL_UTIL_EXPONENTIAL) {
    if (p != 1) return fail(h, BOPL_ERR_INVALID, "exponential utility needs p == 1");
  } else if (util_kind == BOPL_UTIL_CONSTRAINED) {
    if (p != h->m - 1 || h->m < 2) return fail(h, BOPL_ERR_INVALID, "constrained utility needs m >= 2 and p == m - 1 thresholds");
  } else {
    return fail(h, BOPL_ERR_INVALID, "unknown utility kind (no CPU fallback for arbitrary callables)");
  }
  return BOPL_OK;
}

int ensure_mv(bopl_handle* h, size_t count) {
  int rc = ensure_bytes(h, (void**)&h->mean_s, &h->mean_bytes, count * sizeof(double));
  if (rc) return rc;
  return ensure_bytes(h, (void**)&h->var_s, &h->var_bytes, count * sizeof(double));
}

int ensure_grad(bopl_handle* h, size_t count) {
  int rc = ensure_bytes(h, (void**)&h->dmean_s, &h->dmean_bytes, count * sizeof(double));
  if (rc) return rc;
  return ensure_bytes(h, (void**)&h->dvar_s, &h->dvar_bytes, count * sizeof(double));
}

int train_mean(bopl_handle* h, double* F, cudaStream_t st) {
  for (int hh = 0; hh < h->H; ++hh) {
    int rc = launch_posterior_mean_strided(h, h->X_dev, h->n, hh, F + (size_t)hh * h->m * h->n, h->n, st);
    if (rc) return rc;
  }
  return BOPL_OK;
}

// acq [N] on the device, all inputs on the device.
int uei_device(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind, const double* theta,
               int Lt, int p, const double* wts, const double* best, int nH, double* acq, cudaStream_t st) {
  int rc = ensure_mv(h, (size_t)h->m * N);
  if (rc) return rc;
  const double scale = 1.0 / ((double)nH * (double)nw);
  for (int hh = 0; hh < nH; ++hh) {
    rc = launch_posterior_trsm(h, Xc, N, hh, h->mean_s, h->var_s, BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP, st);
    if (rc) return rc;
    rc = launch_uei_mc(h, h->mean_s, h->var_s, N, Z, nw, util_kind, theta, Lt, p, wts, best + (size_t)hh * Lt, scale,
                       hh > 0, acq, st);
    if (rc) return rc;
  }
  return BOPL_OK;
}

}  // namespace

}  // namespace bopl

using namespace bopl;

extern "C" {

int bopl_create(bopl_handle** out, int device) {
  if (!out) return fail(nullptr, BOPL_ERR_INVALID, "out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, BOPL_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, BOPL_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, BOPL_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, BOPL_ERR_UNSUPPORTED,
                std::string("libbopl_b200 is built for sm_100a only; device is ") + prop.name + " (sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  bopl_handle* h = new bopl_handle();
  h->device = device;
  h->sms = prop.multiProcessorCount;
  if (const char* v = getenv("BOPL_SMALL_PATH")) h->small_path = atoi(v);
  if (const char* v = getenv("BOPL_OZ_CLUSTER")) h->oz_cluster = atoi(v);
  if (const char* v = getenv("BOPL_POSTERIOR")) {        // "int8x7" (default), "int8" (6 digits), "fp64"
    const std::string k(v);
    if (k == "int8" || k == "int8x7" || k == "int8auto") {
      h->use_ozaki = 1;
      h->oz_digits = k == "int8x7" ? 7 : 6;
      h->oz_auto = k == "int8auto";
    } else if (k == "fp64") {
      h->use_ozaki = 0;
    } else {
      delete h;
      return fail(nullptr, BOPL_ERR_INVALID, "BOPL_POSTERIOR must be int8x7, int8, int8auto or fp64");
    }
  }
  if (const char* v = getenv("BOPL_TRSM_STAGES")) h->trsm_stages = atoi(v);
  if (const char* v = getenv("BOPL_TRSM_KEEP_BLOCKS")) h->trsm_keep_blocks = atoi(v);
  if (const char* v = getenv("BOPL_TRSM_L2_HINTS")) h->trsm_l2_hints = atoi(v);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking)) != cudaSuccess) {
    delete h;
    return fail(nullptr, BOPL_ERR_CUDA, cudaGetErrorString(e));
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming);
  *out = h;
  return BOPL_OK;
}

void bopl_destroy(bopl_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_model(h);
  void* bufs[] = {h->trsm_scratch, h->mean_s, h->var_s, h->dmean_s, h->dvar_s, h->grad_scratch, h->acq_s, h->F_s,
                  h->topk_val_s, h->topk_idx_s, h->hx_dev, h->hacq_dev, h->hsmall_dev, h->small_scratch, h->gh_dev, h->fit_ws, h->oz_scratch, h->coll_s};
  for (void* p : bufs)
    if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  delete h;
}

const char* bopl_last_error(const bopl_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int bopl_device_sms(const bopl_handle* h) { return h ? h->sms : 0; }

const char* bopl_build_info(void) { return "libbopl_b200 sm_100a (compute_100a) fp64 DMMA.8x8x4 " __DATE__; }

long long bopl_launch_count(const bopl_handle* h) { return h ? h->launches : 0; }

int bopl_set_posterior_kernel(bopl_handle* h, int kind) {
  if (!h) return fail(nullptr, BOPL_ERR_INVALID, "null handle");
  if (kind == BOPL_POSTERIOR_INT8_AUTO) {       // 6 digits where an upload-time probe of the model allows it, else 7: decided per model
    if (h->has_model) return fail(h, BOPL_ERR_STATE, "select BOPL_POSTERIOR_INT8_AUTO before bopl_set_model / bopl_fit_model: the choice is made at upload");
    h->use_ozaki = 1;
    h->oz_digits = 6;
    h->oz_auto = true;
    return BOPL_OK;
  }
  if (kind != BOPL_POSTERIOR_FP64 && kind != BOPL_POSTERIOR_INT8 && kind != BOPL_POSTERIOR_INT8_7)
    return fail(h, BOPL_ERR_INVALID, "unknown posterior kernel kind");
  h->oz_auto = false;
  if (kind != BOPL_POSTERIOR_FP64 && h->has_model && ozaki_covers(h->d, h->n_pad) && (!h->has_oz || h->oz_digits != kind))
    return fail(h, BOPL_ERR_STATE, "the loaded model carries no digit images of that width: select the kernel before bopl_set_model / bopl_fit_model");
  if (kind == BOPL_POSTERIOR_FP64) {
    h->use_ozaki = 0;
  } else {
    h->use_ozaki = 1;
    h->oz_digits = kind;
  }
  return BOPL_OK;
}

int bopl_get_posterior_kernel(const bopl_handle* h) {
  if (!h) return -1;
  return h->use_ozaki ? h->oz_digits : BOPL_POSTERIOR_FP64;
}

static int set_model_impl(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                          const double* ell_h, const double* var_h, const double* noise_h, const double* ymean_h,
                          const double* L_h, const double* alpha_h, const double* Winv_h, const double* Linv_h) {
  BOPL_CHECK_HANDLE(h);
  if (n <= 0 || d <= 0 || m <= 0 || H <= 0) return fail(h, BOPL_ERR_INVALID, "n, d, m, H must be positive");
  if (d > MAX_D) return fail(h, BOPL_ERR_UNSUPPORTED, "d > 24 input dimensions");
  if (m > MAX_M) return fail(h, BOPL_ERR_UNSUPPORTED, "m > 16 attributes");
  if (!kern_kind_h || !X_h || !ell_h || !var_h || !noise_h || !ymean_h || !L_h || !alpha_h)
    return fail(h, BOPL_ERR_INVALID, "null model array");
  for (int i = 0; i < m * H; ++i) {
    if (kern_kind_h[i] < BOPL_KERN_MAT52 || kern_kind_h[i] > BOPL_KERN_EXPQUAD) return fail(h, BOPL_ERR_INVALID, "unknown kernel kind");
    if (!(var_h[i] > 0.0) || !(noise_h[i] >= 0.0)) return fail(h, BOPL_ERR_INVALID, "kernel variance must be > 0 and noise >= 0");
    for (int q = 0; q < d; ++q)
      if (!(ell_h[(size_t)i * d + q] > 0.0)) return fail(h, BOPL_ERR_INVALID, "lengthscales must be > 0");
  }
  BOPL_CUDA(h, cudaDeviceSynchronize());
  free_model(h);
  h->n = n;
  h->d = d;
  h->m = m;
  h->H = H;
  h->n_pad = ((n + TRSM_BM - 1) / TRSM_BM) * TRSM_BM;
  h->pad = h->n_pad - n;
  const int n_pad = h->n_pad, pad = h->pad, nblk = n_pad / TRSM_BM;
  int rc;
  if ((rc = dev_upload(h, X_h, (size_t)n * d, &h->X_dev))) return rc;

  std::vector<double> Lp((size_t)n_pad * n_pad);
  std::vector<double> Dinv((size_t)nblk * TRSM_BM * TRSM_BM);
  std::vector<double> DinvFrag((size_t)nblk * 136 * 64);
  std::vector<double> Xs((size_t)n_pad * d), al((size_t)n_pad), SBlk((size_t)n_pad * (d + 1));
  h->gp_host.resize((size_t)m * H);
  for (int g = 0; g < m * H; ++g) {
    const double* L = L_h + (size_t)g * n * n;
    // front-padded factor: identity on the pad x pad leading block, the real factor below/right of it
    std::fill(Lp.begin(), Lp.end(), 0.0);
    for (int i = 0; i < pad; ++i) Lp[(size_t)i * n_pad + i] = 1.0;
    for (int i = 0; i < n; ++i) {
      double* row = Lp.data() + (size_t)(pad + i) * n_pad + pad;
      const double* src = L + (size_t)i * n;
      for (int k = 0; k <= i; ++k) row[k] = src[k];
    }
    {  // the nblk diagonal-block inversions (extended precision) are independent: spread them over the host cores
      std::atomic<int> next(0), bad(0);
      auto work = [&]() {
        for (int b = next.fetch_add(1); b < nblk; b = next.fetch_add(1)) {
          if (!invert_lower_block(Lp.data() + (size_t)b * TRSM_BM * n_pad + (size_t)b * TRSM_BM, n_pad, TRSM_BM,
                                  Dinv.data() + (size_t)b * TRSM_BM * TRSM_BM)) {
            bad.store(1);
            continue;
          }
          pack_dinv_fragments(Dinv.data() + (size_t)b * TRSM_BM * TRSM_BM, DinvFrag.data() + (size_t)b * 136 * 64);
        }
      };
      const int nthr = std::max(1, std::min<int>(nblk, (int)std::thread::hardware_concurrency()));
      std::vector<std::thread> pool;
      for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
      work();
      for (auto& th : pool) th.join();
      if (bad.load()) {
        free_model(h);
        return fail(h, BOPL_ERR_INVALID, "Cholesky factor has a non-positive diagonal entry");
      }
    }
    // int8 digit images for the INT8-tensor-pipe variant (posterior_ozaki.cu): from the un-negated factor
    std::vector<signed char> oz_img;
    std::vector<double> oz_scale, oz_d64;
    const bool oz = h->use_ozaki && ozaki_covers(d, n_pad);
    if (oz) {
      ozaki_prepare(Lp.data(), n_pad, h->oz_digits, oz_img, oz_scale);
      const int nb64 = n_pad / 64;
      oz_d64.resize((size_t)nb64 * 64 * 64);
      for (int b = 0; b < nb64; ++b)
        if (!invert_lower_block(Lp.data() + (size_t)b * 64 * n_pad + (size_t)b * 64, n_pad, 64, oz_d64.data() + (size_t)b * 64 * 64)) {
          free_model(h);
          return fail(h, BOPL_ERR_INVALID, "Cholesky factor has a non-positive diagonal entry");
        }
    }
    // negate the off-diagonal blocks so that the kernel's update is a pure accumulate: acc += (-L) V
    for (int i = 0; i < n_pad; ++i) {
      const int bi = i / TRSM_BM;
      double* row = Lp.data() + (size_t)i * n_pad;
      for (int k = 0; k < bi * TRSM_BM; ++k) row[k] = -row[k];
    }
    std::fill(Xs.begin(), Xs.end(), 0.0);
    std::fill(al.begin(), al.end(), 0.0);
    const double* ell = ell_h + (size_t)g * d;
    for (int i = 0; i < n; ++i) {
      for (int q = 0; q < d; ++q) Xs[(size_t)(pad + i) * d + q] = X_h[(size_t)i * d + q] / ell[q];   // stationary.py:161-162
      al[pad + i] = alpha_h[(size_t)g * n + i];
    }
    pack_stage_blocks(Xs.data(), al.data(), n_pad, d, SBlk.data());
    signed char* pOzL = nullptr;
    double* pOzS = nullptr;
    if (oz) {
      std::vector<double> oz_stage;
      ozaki_pack_stage(oz_d64.data(), Xs.data(), al.data(), oz_scale.data(), n_pad, d, oz_stage);
      if ((rc = dev_upload(h, oz_img.data(), oz_img.size(), &pOzL)) || (rc = dev_upload(h, oz_stage.data(), oz_stage.size(), &pOzS))) {
        free_model(h);
        return rc;
      }
    }
    GpDev gp{};
    double *pL, *pD, *pF, *pX, *pa, *pe, *pS, *pW = nullptr;
    if ((rc = dev_upload(h, Lp.data(), Lp.size(), &pL)) || (rc = dev_upload(h, Dinv.data(), Dinv.size(), &pD)) ||
        (rc = dev_upload(h, DinvFrag.data(), DinvFrag.size(), &pF)) || (rc = dev_upload(h, SBlk.data(), SBlk.size(), &pS)) ||
        (rc = dev_upload(h, Xs.data(), Xs.size(), &pX)) || (rc = dev_upload(h, al.data(), al.size(), &pa)) ||
        (rc = dev_upload(h, ell, (size_t)d, &pe))) {
      free_model(h);
      return rc;
    }
    if (Winv_h && (rc = dev_upload(h, Winv_h + (size_t)g * n * n, (size_t)n * n, &pW))) {
      free_model(h);
      return rc;
    }
    double* pLi = nullptr;
    if (Linv_h && (rc = dev_upload(h, Linv_h + (size_t)g * n * n, (size_t)n * n, &pLi))) {
      free_model(h);
      return rc;
    }
    gp.Linv = pLi;
    gp.OzL = pOzL;
    gp.OzStage = pOzS;
    gp.Lneg = pL;
    gp.Dinv = pD;
    gp.DinvFrag = pF;
    gp.StageBlk = pS;
    gp.Xs = pX;
    gp.alpha = pa;
    gp.ell = pe;
    gp.Winv = pW;
    gp.variance = var_h[g];
    gp.noise = noise_h[g];
    gp.ymean = ymean_h[g];
    gp.kind = kern_kind_h[g];
    h->gp_host[g] = gp;
  }
  GpDev* gd;
  if ((rc = dev_upload(h, h->gp_host.data(), h->gp_host.size(), &gd))) {
    free_model(h);
    return rc;
  }
  if ((rc = uploads_done(h))) {
    free_model(h);
    return rc;
  }
  h->gp_dev = gd;
  h->has_winv = Winv_h != nullptr;
  h->has_linv = Linv_h != nullptr;
  h->has_oz = h->use_ozaki != 0 && ozaki_covers(d, n_pad);
  h->has_model = true;
  h->F_valid = false;
  return BOPL_OK;
}

int bopl_cross_cov(bopl_handle* h, const double* Xc, int N, int hsel, int j, double* K, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || hsel < 0 || hsel >= h->H || j < 0 || j >= h->m) return fail(h, BOPL_ERR_INVALID, "bad N / hsel / j");
  if (N && (!Xc || !K)) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return launch_cross_cov(h, Xc, N, hsel, j, K, (cudaStream_t)stream);
}

int bopl_posterior(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                   void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || hsel < 0 || hsel >= h->H) return fail(h, BOPL_ERR_INVALID, "bad N / hsel");
  if (N && !Xc) return fail(h, BOPL_ERR_INVALID, "null candidates");
  if (!var) return mean ? launch_posterior_mean(h, Xc, N, hsel, mean, (cudaStream_t)stream) : BOPL_OK;
  return launch_posterior_trsm(h, Xc, N, hsel, mean, var, flags, (cudaStream_t)stream);
}

int bopl_posterior_mean(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || hsel < 0 || hsel >= h->H) return fail(h, BOPL_ERR_INVALID, "bad N / hsel");
  if (N && (!Xc || !mean)) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return launch_posterior_mean(h, Xc, N, hsel, mean, (cudaStream_t)stream);
}

int bopl_train_mean(bopl_handle* h, double* F, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (!F) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return train_mean(h, F, (cudaStream_t)stream);
}

int bopl_posterior_grad(bopl_handle* h, const double* Xc, int N, int hsel, double* dmean, double* dvar,
                        void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || hsel < 0 || hsel >= h->H) return fail(h, BOPL_ERR_INVALID, "bad N / hsel");
  if (N && !Xc) return fail(h, BOPL_ERR_INVALID, "null candidates");
  if (dvar && !h->has_winv) return fail(h, BOPL_ERR_STATE, "variance gradient needs Winv (bopl_set_model was given NULL)");
  return launch_posterior_grad(h, Xc, N, hsel, dmean, dvar, (cudaStream_t)stream);
}

int bopl_incumbent(bopl_handle* h, int util_kind, const double* theta, int Lt, int p, double* best, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (!theta || !best) return fail(h, BOPL_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  size_t need = (size_t)h->H * h->m * h->n * sizeof(double);
  if (!h->F_s || h->F_bytes < need) {
    rc = ensure_bytes(h, (void**)&h->F_s, &h->F_bytes, need);
    if (rc) return rc;
    h->F_valid = false;
  }
  if (!h->F_valid) {
    rc = train_mean(h, h->F_s, st);
    if (rc) return rc;
    h->F_valid = true;   // later calls on other streams: the caller orders streams (calls on a handle are stream-ordered)
  }
  return launch_incumbent(h, h->F_s, util_kind, theta, Lt, p, best, st);
}

int bopl_uei_mc(bopl_handle* h, const double* mean, const double* var, int N, const double* Z, int nw,
                int util_kind, const double* theta, int Lt, int p, const double* wts, const double* best,
                double scale, int accumulate, double* acq, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (util_kind == BOPL_UTIL_CONSTRAINED) return fail(h, BOPL_ERR_INVALID, "the constrained utility has no MC kernel: use bopl_uei_constrained");
  if (N < 0 || nw <= 0) return fail(h, BOPL_ERR_INVALID, "bad N / nw");
  if (N && (!mean || !var || !Z || !theta || !wts || !best || !acq)) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return launch_uei_mc(h, mean, var, N, Z, nw, util_kind, theta, Lt, p, wts, best, scale, accumulate, acq,
                       (cudaStream_t)stream);
}

int bopl_uei(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
             const double* theta, int Lt, int p, const double* wts, const double* best, int nH, double* acq,
             void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (N < 0 || nw <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / nw / nH");
  if (N == 0) return BOPL_OK;
  if (!Xc || !Z || !theta || !wts || !best || !acq) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return uei_device(h, Xc, N, Z, nw, util_kind, theta, Lt, p, wts, best, nH, acq, (cudaStream_t)stream);
}

int bopl_uei_grad(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
                  const double* theta, int Lt, int p, const double* wts, const double* best, int nH,
                  double* acq, double* dacq, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (N < 0 || nw <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / nw / nH");
  if (N == 0) return BOPL_OK;
  if (!Xc || !Z || !theta || !wts || !best || !acq || !dacq) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (!h->has_winv) return fail(h, BOPL_ERR_STATE, "gradients need Winv (bopl_set_model was given NULL)");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = ensure_mv(h, (size_t)h->m * N)) || (rc = ensure_grad(h, (size_t)h->m * N * h->d))) return rc;
  const double scale = 1.0 / ((double)nH * (double)nw);
  GradFollows fuse(h, true);
  for (int hh = 0; hh < nH; ++hh) {
    if ((rc = launch_posterior_trsm(h, Xc, N, hh, h->mean_s, h->var_s, BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP, st))) return rc;
    if ((rc = launch_posterior_grad(h, Xc, N, hh, h->dmean_s, h->dvar_s, st))) return rc;
    if ((rc = launch_uei_mc_grad(h, h->mean_s, h->var_s, h->dmean_s, h->dvar_s, N, Z, nw, util_kind, theta, Lt, p, wts,
                                 best + (size_t)hh * Lt, scale, hh > 0, acq, dacq, st)))
      return rc;
  }
  return BOPL_OK;
}

int bopl_uei_grad_host(bopl_handle* h, const double* Xc_h, int N, const double* Z_h, int nw, int util_kind,
                       const double* theta_h, int Lt, int p, const double* wts_h, const double* best_h, int nH,
                       double* acq_h, double* dacq_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, h->s_compute);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (N < 0 || nw <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / nw / nH");
  if (N == 0) return BOPL_OK;
  if (!Xc_h || !Z_h || !theta_h || !wts_h || !best_h || !acq_h || !dacq_h) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (!h->has_winv) return fail(h, BOPL_ERR_STATE, "gradients need Winv (bopl_set_model was given NULL)");
  const int d = h->d;
  const size_t nX = (size_t)N * d, nZ = (size_t)nw * h->m, nT = (size_t)Lt * p, nB = (size_t)nH * Lt;
  const size_t total = nX + nZ + nT + Lt + nB + N + nX;     // X | Z | theta | wts | best | acq | dacq
  if ((rc = ensure_bytes(h, (void**)&h->gh_dev, &h->gh_cap, total * sizeof(double)))) return rc;
  double* Xd = h->gh_dev;
  double* Zd = Xd + nX;
  double* thd = Zd + nZ;
  double* wd = thd + nT;
  double* bd = wd + Lt;
  double* ad = bd + nB;
  double* gd = ad + N;
  cudaStream_t st = h->s_compute;
  BOPL_CUDA(h, cudaMemcpyAsync(Xd, Xc_h, nX * sizeof(double), cudaMemcpyHostToDevice, st));
  BOPL_CUDA(h, cudaMemcpyAsync(Zd, Z_h, nZ * sizeof(double), cudaMemcpyHostToDevice, st));
  BOPL_CUDA(h, cudaMemcpyAsync(thd, theta_h, nT * sizeof(double), cudaMemcpyHostToDevice, st));
  BOPL_CUDA(h, cudaMemcpyAsync(wd, wts_h, (size_t)Lt * sizeof(double), cudaMemcpyHostToDevice, st));
  BOPL_CUDA(h, cudaMemcpyAsync(bd, best_h, nB * sizeof(double), cudaMemcpyHostToDevice, st));
  if ((rc = ensure_mv(h, (size_t)h->m * N)) || (rc = ensure_grad(h, (size_t)h->m * N * d))) return rc;
  const double scale = 1.0 / ((double)nH * (double)nw);
  GradFollows fuse(h, true);
  for (int hh = 0; hh < nH; ++hh) {
    if ((rc = launch_posterior_trsm(h, Xd, N, hh, h->mean_s, h->var_s, BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP, st))) return rc;
    if ((rc = launch_posterior_grad(h, Xd, N, hh, h->dmean_s, h->dvar_s, st))) return rc;
    if ((rc = launch_uei_mc_grad(h, h->mean_s, h->var_s, h->dmean_s, h->dvar_s, N, Zd, nw, util_kind, thd, Lt, p, wd,
                                 bd + (size_t)hh * Lt, scale, hh > 0, ad, gd, st)))
      return rc;
  }
  BOPL_CUDA(h, cudaMemcpyAsync(acq_h, ad, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaMemcpyAsync(dacq_h, gd, nX * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaStreamSynchronize(st));
  return BOPL_OK;
}

int bopl_expected_utility(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
                          const double* theta, int Lt, int p, const double* wts, int nH, double* val, double* dval,
                          void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (util_kind == BOPL_UTIL_CONSTRAINED) return fail(h, BOPL_ERR_INVALID, "the constrained utility has no MC kernel");
  if (N < 0 || nw <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / nw / nH");
  if (N == 0) return BOPL_OK;
  if (!Xc || !Z || !theta || !wts || !val) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (dval && !h->has_winv) return fail(h, BOPL_ERR_STATE, "gradients need Winv (bopl_set_model was given NULL)");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = ensure_mv(h, (size_t)h->m * N))) return rc;
  if (dval && (rc = ensure_grad(h, (size_t)h->m * N * h->d))) return rc;
  GradFollows fuse(h, dval != nullptr);
  for (int hh = 0; hh < nH; ++hh) {
    // predict_noiseless: no likelihood variance, clipped at 1e-10 (gpmodel.py:181-190)
    if ((rc = launch_posterior_trsm(h, Xc, N, hh, h->mean_s, h->var_s, BOPL_VAR_CLIP, st))) return rc;
    if (dval) {
      if ((rc = launch_posterior_grad(h, Xc, N, hh, h->dmean_s, h->dvar_s, st))) return rc;
      if ((rc = launch_uei_mc_grad(h, h->mean_s, h->var_s, h->dmean_s, h->dvar_s, N, Z, nw, util_kind, theta, Lt, p, wts,
                                   nullptr, 1.0, hh > 0, val, dval, st, 1)))
        return rc;
    } else if ((rc = launch_uei_mc(h, h->mean_s, h->var_s, N, Z, nw, util_kind, theta, Lt, p, wts, nullptr, 1.0, hh > 0,
                                   val, st, 1))) {
      return rc;
    }
  }
  return BOPL_OK;
}

int bopl_uei_affine(bopl_handle* h, const double* Xc, int N, const double* theta, int Lt, const double* wts,
                    const double* best, int nH, double* acq, double* dacq, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || Lt <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / Lt / nH");
  if (N == 0) return BOPL_OK;
  if (!Xc || !theta || !wts || !best || !acq) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (dacq && !h->has_winv) return fail(h, BOPL_ERR_STATE, "gradients need Winv (bopl_set_model was given NULL)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = ensure_mv(h, (size_t)h->m * N))) return rc;
  if (dacq && (rc = ensure_grad(h, (size_t)h->m * N * h->d))) return rc;
  const double scale = 1.0 / (double)nH;
  GradFollows fuse(h, dacq != nullptr);
  for (int hh = 0; hh < nH; ++hh) {
    if ((rc = launch_posterior_trsm(h, Xc, N, hh, h->mean_s, h->var_s, BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP, st))) return rc;
    if (dacq && (rc = launch_posterior_grad(h, Xc, N, hh, h->dmean_s, h->dvar_s, st))) return rc;
    if ((rc = launch_uei_affine(h, h->mean_s, h->var_s, dacq ? h->dmean_s : nullptr, dacq ? h->dvar_s : nullptr, N, theta,
                                Lt, wts, best + (size_t)hh * Lt, scale, hh > 0, acq, dacq, st)))
      return rc;
  }
  return BOPL_OK;
}

int bopl_uei_constrained(bopl_handle* h, const double* Xc, int N, const double* theta, int Lt, const double* wts,
                         const double* best, int nH, double* acq, double* dacq, void* stream) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || Lt <= 0 || nH <= 0 || nH > h->H) return fail(h, BOPL_ERR_INVALID, "bad N / Lt / nH");
  if (h->m < 2) return fail(h, BOPL_ERR_INVALID, "constrained EI needs at least one constraint attribute (m >= 2)");
  if (N == 0) return BOPL_OK;
  if (!Xc || !theta || !wts || !best || !acq) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (dacq && !h->has_winv) return fail(h, BOPL_ERR_STATE, "gradients need Winv (bopl_set_model was given NULL)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = ensure_mv(h, (size_t)h->m * N))) return rc;
  if (dacq && (rc = ensure_grad(h, (size_t)h->m * N * h->d))) return rc;
  const double scale = 1.0 / (double)nH;
  GradFollows fuse(h, dacq != nullptr);
  for (int hh = 0; hh < nH; ++hh) {
    if ((rc = launch_posterior_trsm(h, Xc, N, hh, h->mean_s, h->var_s, BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP, st))) return rc;
    if (dacq && (rc = launch_posterior_grad(h, Xc, N, hh, h->dmean_s, h->dvar_s, st))) return rc;
    if ((rc = launch_uei_constrained(h, h->mean_s, h->var_s, dacq ? h->dmean_s : nullptr, dacq ? h->dvar_s : nullptr, N,
                                     theta, Lt, wts, best + (size_t)hh * Lt, scale, hh > 0, acq, dacq, st)))
      return rc;
  }
  return BOPL_OK;
}

int bopl_topk(bopl_handle* h, const double* acq, int N, int k, long long index_offset, double* val,
              long long* idx, void* stream) {
  BOPL_CHECK_HANDLE(h);
  StreamOrder order_guard(h, (cudaStream_t)stream);
  if (N < 0 || k < 0) return fail(h, BOPL_ERR_INVALID, "bad N / k");
  if (k == 0) return BOPL_OK;
  if (!acq || !val || !idx) return fail(h, BOPL_ERR_INVALID, "null buffer");
  return launch_topk(h, acq, N, k, index_offset, val, idx, (cudaStream_t)stream);
}

int bopl_uei_host(bopl_handle* h, const double* Xc_h, int N, const double* Z_h, int nw, int util_kind,
                  const double* theta_h, int Lt, int p, const double* wts_h, const double* best_h, int nH,
                  double* acq_h, int k, long long index_offset, double* topk_val_h, long long* topk_idx_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  StreamOrder order_guard(h, h->s_compute);
  int rc = check_utility(h, util_kind, Lt, p);
  if (rc) return rc;
  if (N < 0 || nw <= 0 || nH <= 0 || nH > h->H || k < 0) return fail(h, BOPL_ERR_INVALID, "bad N / nw / nH / k");
  if (N == 0) return k ? fail(h, BOPL_ERR_INVALID, "topk of an empty batch") : BOPL_OK;
  if (!Xc_h || !Z_h || !theta_h || !wts_h || !best_h) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (k && (!topk_val_h || !topk_idx_h)) return fail(h, BOPL_ERR_INVALID, "null top-k buffer");
  if (k > N) return fail(h, BOPL_ERR_INVALID, "topk: fewer candidates than k");
  const int d = h->d;
  // chunk = a whole number of persistent waves of the posterior kernel (sms candidate tiles of 128) so no chunk has a tail
  const int chunk = std::min(N, h->sms * TRSM_BC * 8);
  const size_t small_cnt = (size_t)nw * h->m + (size_t)Lt * p + Lt + (size_t)nH * Lt + k + k;
  if ((rc = ensure_bytes(h, (void**)&h->hx_dev, &h->hx_cap, (size_t)2 * chunk * d * sizeof(double)))) return rc;
  if ((rc = ensure_bytes(h, (void**)&h->hacq_dev, &h->hacq_cap, (size_t)N * sizeof(double)))) return rc;
  if ((rc = ensure_bytes(h, (void**)&h->hsmall_dev, &h->hsmall_cap, small_cnt * sizeof(double)))) return rc;
  double* Zd = h->hsmall_dev;
  double* thd = Zd + (size_t)nw * h->m;
  double* wd = thd + (size_t)Lt * p;
  double* bd = wd + Lt;
  double* tkv = bd + (size_t)nH * Lt;
  long long* tki = (long long*)(tkv + k);
  cudaStream_t sc = h->s_compute, sy = h->s_copy;
  BOPL_CUDA(h, cudaMemcpyAsync(Zd, Z_h, (size_t)nw * h->m * sizeof(double), cudaMemcpyHostToDevice, sc));
  BOPL_CUDA(h, cudaMemcpyAsync(thd, theta_h, (size_t)Lt * p * sizeof(double), cudaMemcpyHostToDevice, sc));
  BOPL_CUDA(h, cudaMemcpyAsync(wd, wts_h, (size_t)Lt * sizeof(double), cudaMemcpyHostToDevice, sc));
  BOPL_CUDA(h, cudaMemcpyAsync(bd, best_h, (size_t)nH * Lt * sizeof(double), cudaMemcpyHostToDevice, sc));
  int ci = 0;
  for (int s = 0; s < N; s += chunk, ++ci) {
    const int nc = std::min(chunk, N - s), buf = ci & 1;
    double* xb = h->hx_dev + (size_t)buf * chunk * d;
    if (ci >= 2) BOPL_CUDA(h, cudaStreamWaitEvent(sy, h->ev_done[buf], 0));
    BOPL_CUDA(h, cudaMemcpyAsync(xb, Xc_h + (size_t)s * d, (size_t)nc * d * sizeof(double), cudaMemcpyHostToDevice, sy));
    BOPL_CUDA(h, cudaEventRecord(h->ev_in[buf], sy));
    BOPL_CUDA(h, cudaStreamWaitEvent(sc, h->ev_in[buf], 0));
    if ((rc = uei_device(h, xb, nc, Zd, nw, util_kind, thd, Lt, p, wd, bd, nH, h->hacq_dev + s, sc))) return rc;
    BOPL_CUDA(h, cudaEventRecord(h->ev_done[buf], sc));
  }
  if (acq_h) BOPL_CUDA(h, cudaMemcpyAsync(acq_h, h->hacq_dev, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, sc));
  if (k) {
    if ((rc = launch_topk(h, h->hacq_dev, N, k, index_offset, tkv, tki, sc))) return rc;
    BOPL_CUDA(h, cudaMemcpyAsync(topk_val_h, tkv, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, sc));
    BOPL_CUDA(h, cudaMemcpyAsync(topk_idx_h, tki, (size_t)k * sizeof(long long), cudaMemcpyDeviceToHost, sc));
  }
  BOPL_CUDA(h, cudaStreamSynchronize(sc));
  return BOPL_OK;
}

// ---- model fit on the device (fit_kernels.cu) ---------------------------------------------------------------------------
static int fit_model_impl(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                          const double* Y_h, const double* ell_h, const double* var_h, const double* noise_h,
                          const double* jitter_h, int want_winv, int want_linv, double* logdet_h) {
  BOPL_CHECK_HANDLE(h);
  if (n <= 0 || d <= 0 || m <= 0 || H <= 0) return fail(h, BOPL_ERR_INVALID, "n, d, m, H must be positive");
  if (d > MAX_D) return fail(h, BOPL_ERR_UNSUPPORTED, "d > 24 input dimensions");
  if (m > MAX_M) return fail(h, BOPL_ERR_UNSUPPORTED, "m > 16 attributes");
  if (!kern_kind_h || !X_h || !Y_h || !ell_h || !var_h || !noise_h) return fail(h, BOPL_ERR_INVALID, "null model array");
  const int G = m * H;
  for (int i = 0; i < G; ++i) {
    if (kern_kind_h[i] < BOPL_KERN_MAT52 || kern_kind_h[i] > BOPL_KERN_EXPQUAD) return fail(h, BOPL_ERR_INVALID, "unknown kernel kind");
    if (!(var_h[i] > 0.0) || !(noise_h[i] >= 0.0)) return fail(h, BOPL_ERR_INVALID, "kernel variance must be > 0 and noise >= 0");
    if (jitter_h && !(jitter_h[i] >= 0.0)) return fail(h, BOPL_ERR_INVALID, "jitter must be >= 0");
    for (int q = 0; q < d; ++q)
      if (!(ell_h[(size_t)i * d + q] > 0.0)) return fail(h, BOPL_ERR_INVALID, "lengthscales must be > 0");
  }
  BOPL_CUDA(h, cudaDeviceSynchronize());
  const bool timing = getenv("BOPL_FIT_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  free_model(h);
  const double t1 = now();
  h->n = n;
  h->d = d;
  h->m = m;
  h->H = H;
  h->n_pad = ((n + TRSM_BM - 1) / TRSM_BM) * TRSM_BM;
  h->pad = h->n_pad - n;
  const int n_pad = h->n_pad, nblk = n_pad / TRSM_BM;
  const bool inverse = want_winv || want_linv;
  int rc;
  if ((rc = dev_upload(h, X_h, (size_t)n * d, &h->X_dev))) return rc;
  // centred targets per attribute (this fork's normaliser subtracts the mean only: util/normalizer.py:57-72)
  std::vector<double> yc((size_t)m * n), ymean(m);
  for (int j = 0; j < m; ++j) {
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += Y_h[(size_t)j * n + i];
    ymean[j] = (double)(s / n);
    for (int i = 0; i < n; ++i) yc[(size_t)j * n + i] = Y_h[(size_t)j * n + i] - ymean[j];
  }
  double *yc_dev, *logdiag_dev, *quad_dev;
  int* info_dev;
  if ((rc = dev_upload(h, yc.data(), yc.size(), &yc_dev))) { free_model(h); return rc; }
  std::vector<double> zeros((size_t)G * nblk + G, 0.0);
  std::vector<int> izeros(G, 0);
  if ((rc = dev_upload(h, zeros.data(), (size_t)G * nblk, &logdiag_dev)) || (rc = dev_upload(h, zeros.data(), (size_t)G, &quad_dev)) ||
      (rc = dev_upload(h, izeros.data(), (size_t)G, &info_dev))) { free_model(h); return rc; }
  const size_t fsz = fit_dev_struct_bytes();
  std::vector<char> fits((size_t)G * fsz);
  std::vector<double*> Us;
  h->gp_host.resize((size_t)G);
  auto dalloc = [&](size_t count, double** out, bool owned) -> int {
    void* p = nullptr;
    BOPL_CUDA(h, cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(double)));
    if (owned) h->owned.push_back(p);
    *out = (double*)p;
    return BOPL_OK;
  };
  auto cleanup = [&]() {
    for (double* u : Us) cudaFree(u);
    Us.clear();
  };
  for (int g = 0; g < G; ++g) {
    double *pL, *pD, *pF, *pS, *pX, *pa, *pe, *pW = nullptr, *pLi = nullptr, *pU = nullptr;
    if ((rc = dalloc((size_t)n_pad * n_pad, &pL, true)) || (rc = dalloc((size_t)nblk * TRSM_BM * TRSM_BM, &pD, true)) ||
        (rc = dalloc((size_t)nblk * 136 * 64, &pF, true)) || (rc = dalloc((size_t)n_pad * (d + 1), &pS, true)) ||
        (rc = dalloc((size_t)n_pad * d, &pX, true)) || (rc = dalloc((size_t)n_pad, &pa, true)) ||
        (rc = dev_upload(h, ell_h + (size_t)g * d, (size_t)d, &pe)) ||
        (want_winv && (rc = dalloc((size_t)n * n, &pW, true))) || (want_linv && (rc = dalloc((size_t)n * n, &pLi, true))) ||
        (inverse && (rc = dalloc((size_t)n_pad * n_pad, &pU, false)))) {
      cleanup();
      free_model(h);
      return rc;
    }
    if (pU) Us.push_back(pU);
    cudaMemsetAsync(pL, 0, (size_t)n_pad * n_pad * sizeof(double), h->s_compute);
    const double diag_add = noise_h[g] + 1e-8 + (jitter_h ? jitter_h[g] : 0.0);   // exact_gaussian_inference.py:46
    fit_dev_fill(fits.data() + (size_t)g * fsz, pL, pD, pF, pU, pW, pLi, pX, pa, pS, pe, yc_dev + (size_t)(g / H) * n,
                 logdiag_dev + (size_t)g * nblk, quad_dev + g, info_dev + g, var_h[g], diag_add, kern_kind_h[g]);
    GpDev gp{};
    gp.Linv = pLi;
    gp.Lneg = pL;
    gp.Dinv = pD;
    gp.DinvFrag = pF;
    gp.StageBlk = pS;
    gp.Xs = pX;
    gp.alpha = pa;
    gp.ell = pe;
    gp.Winv = pW;
    gp.variance = var_h[g];
    gp.noise = noise_h[g];
    gp.ymean = ymean[g / H];
    gp.kind = kern_kind_h[g];
    h->gp_host[g] = gp;
  }
  char* fits_dev;
  const double t2 = now();
  if ((rc = dev_upload(h, fits.data(), fits.size(), &fits_dev)) || (rc = uploads_done(h)) ||
      (rc = fit_factorize(h, fits_dev, G, h->X_dev, n, d, inverse, want_winv != 0, want_linv != 0, h->s_compute))) {
    cleanup();
    free_model(h);
    return rc;
  }
  cudaError_t e = cudaStreamSynchronize(h->s_compute);
  const double t3 = now();
  cleanup();
  if (timing)
    fprintf(stderr, "bopl_fit_model: free %.2f ms, alloc %.2f ms, kernels %.2f ms, scratch free %.2f ms\n", t1 - t0, t2 - t1, t3 - t2,
            now() - t3);
  if (e != cudaSuccess) {
    free_model(h);
    return fail(h, BOPL_ERR_CUDA, std::string("device fit: ") + cudaGetErrorString(e));
  }
  std::vector<int> info(G);
  std::vector<double> logdiag((size_t)G * nblk);
  BOPL_CUDA(h, cudaMemcpy(info.data(), info_dev, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost));
  BOPL_CUDA(h, cudaMemcpy(logdiag.data(), logdiag_dev, logdiag.size() * sizeof(double), cudaMemcpyDeviceToHost));
  h->fit_info = info;
  for (int g = 0; g < G; ++g)
    if (info[g]) {
      free_model(h);
      return fail(h, BOPL_ERR_NUMERIC, "covariance matrix of attribute " + std::to_string(g / H) + ", hyper-sample " +
                                           std::to_string(g % H) + " is not positive definite (pivot " +
                                           std::to_string(info[g] - 1 - h->pad) + ")");
    }
  if (logdet_h)
    for (int g = 0; g < G; ++g) {
      double s = 0.0;
      for (int b = 0; b < nblk; ++b) s += logdiag[(size_t)g * nblk + b];
      logdet_h[g] = 2.0 * s;
    }
  bool oz_ok = false;
  if (h->use_ozaki) {                 // int8 digit images + staging blocks of the INT8-tensor-pipe posterior kernel
    if ((rc = ozaki_prepare_device(h, h->s_compute))) {
      free_model(h);
      return rc;
    }
    oz_ok = h->gp_host[0].OzL != nullptr;
  }
  GpDev* gd;
  if ((rc = dev_upload(h, h->gp_host.data(), h->gp_host.size(), &gd))) {
    free_model(h);
    return rc;
  }
  if ((rc = uploads_done(h))) {
    free_model(h);
    return rc;
  }
  h->gp_dev = gd;
  h->has_winv = want_winv != 0;
  h->has_linv = want_linv != 0;
  h->has_oz = oz_ok;
  h->has_model = true;
  h->F_valid = false;
  return BOPL_OK;
}

// BOPL_POSTERIOR_INT8_AUTO: upload with 6 digits, probe at the training points, upload again with 7 when the probe fails.
static int int8_auto_keeps_six(bopl_handle* h, int rc) {
  if (rc != BOPL_OK || !h->oz_auto || !h->has_oz || h->oz_digits != 6) return 1;
  return int8_probe_passes(h);
}

int bopl_set_model(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                   const double* ell_h, const double* var_h, const double* noise_h, const double* ymean_h,
                   const double* L_h, const double* alpha_h, const double* Winv_h, const double* Linv_h) {
  if (h && h->oz_auto) h->oz_digits = 6;
  int rc = set_model_impl(h, n, d, m, H, kern_kind_h, X_h, ell_h, var_h, noise_h, ymean_h, L_h, alpha_h, Winv_h, Linv_h);
  if (rc != BOPL_OK) return rc;
  const int keep = int8_auto_keeps_six(h, rc);
  if (keep < 0) {
    free_model(h);
    return keep;
  }
  if (keep == 0) {
    h->oz_digits = 7;
    rc = set_model_impl(h, n, d, m, H, kern_kind_h, X_h, ell_h, var_h, noise_h, ymean_h, L_h, alpha_h, Winv_h, Linv_h);
  }
  return rc;
}

int bopl_fit_model(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                   const double* Y_h, const double* ell_h, const double* var_h, const double* noise_h,
                   const double* jitter_h, int want_winv, int want_linv, double* logdet_h) {
  if (h && h->oz_auto) h->oz_digits = 6;
  int rc = fit_model_impl(h, n, d, m, H, kern_kind_h, X_h, Y_h, ell_h, var_h, noise_h, jitter_h, want_winv, want_linv, logdet_h);
  if (rc != BOPL_OK) return rc;
  const int keep = int8_auto_keeps_six(h, rc);
  if (keep < 0) {
    free_model(h);
    return keep;
  }
  if (keep == 0) {
    h->oz_digits = 7;
    rc = fit_model_impl(h, n, d, m, H, kern_kind_h, X_h, Y_h, ell_h, var_h, noise_h, jitter_h, want_winv, want_linv, logdet_h);
  }
  return rc;
}

int bopl_fit_info(const bopl_handle* h, int count, int* info_h) {
  if (!h || !info_h || count < 0) return fail(nullptr, BOPL_ERR_INVALID, "null handle / buffer");
  for (int g = 0; g < count; ++g) info_h[g] = g < (int)h->fit_info.size() ? h->fit_info[g] : 0;
  return BOPL_OK;
}

int bopl_log_marginal(bopl_handle* h, int n, int d, int G, const int* kern_kind_h, const double* X_h, const double* Yc_h,
                      const double* ell_h, const double* var_h, const double* noise_h, double* ll_h, double* grad_h,
                      int* info_h) {
  BOPL_CHECK_HANDLE(h);
  if (n <= 0 || d <= 0 || G <= 0) return fail(h, BOPL_ERR_INVALID, "n, d, G must be positive");
  if (d > MAX_D) return fail(h, BOPL_ERR_UNSUPPORTED, "d > 24 input dimensions");
  if (!kern_kind_h || !X_h || !Yc_h || !ell_h || !var_h || !noise_h || !ll_h || !info_h) return fail(h, BOPL_ERR_INVALID, "null array");
  for (int i = 0; i < G; ++i) {
    if (kern_kind_h[i] < BOPL_KERN_MAT52 || kern_kind_h[i] > BOPL_KERN_EXPQUAD) return fail(h, BOPL_ERR_INVALID, "unknown kernel kind");
    if (!(var_h[i] > 0.0) || !(noise_h[i] >= 0.0)) return fail(h, BOPL_ERR_INVALID, "kernel variance must be > 0 and noise >= 0");
    for (int q = 0; q < d; ++q)
      if (!(ell_h[(size_t)i * d + q] > 0.0)) return fail(h, BOPL_ERR_INVALID, "lengthscales must be > 0");
  }
  const int n_pad = ((n + TRSM_BM - 1) / TRSM_BM) * TRSM_BM, nblk = n_pad / TRSM_BM, nstrips = (n + 31) / 32;
  const bool grad = grad_h != nullptr;
  const size_t fsz = fit_dev_struct_bytes();
  // carve one slab: per instance A, Dinv, DinvFrag, Xs, alpha, ell, y (+ U, Winv for the gradient), then the shared pieces
  auto al = [](size_t x) { return (x + 31) & ~(size_t)31; };
  const size_t cA = al((size_t)n_pad * n_pad), cD = al((size_t)nblk * TRSM_BM * TRSM_BM), cF = al((size_t)nblk * 136 * 64),
               cX = al((size_t)n_pad * d), ca = al((size_t)n_pad), ce = al((size_t)d), cy = al((size_t)n), cW = al((size_t)n * n);
  const size_t per = cA + cD + cF + cX + ca + ce + cy + (grad ? cA + cW : 0);
  const size_t shared = al((size_t)n * d) + al((size_t)G * nblk) + al((size_t)G) + al((size_t)G) /*info as doubles' space*/ +
                        al((size_t)G * nstrips * (d + 2)) + al((size_t)G * (d + 2)) + al((G * fsz + 7) / 8);
  int rc;
  if ((rc = ensure_bytes(h, &h->fit_ws, &h->fit_ws_bytes, (per * G + shared) * sizeof(double)))) return rc;
  double* base = (double*)h->fit_ws;
  double* sh = base + per * G;
  double* X_dev = sh;                      sh += al((size_t)n * d);
  double* logdiag_dev = sh;                sh += al((size_t)G * nblk);
  double* quad_dev = sh;                   sh += al((size_t)G);
  int* info_dev = (int*)sh;                sh += al((size_t)G);
  double* partial_dev = sh;                sh += al((size_t)G * nstrips * (d + 2));
  double* out_dev = sh;                    sh += al((size_t)G * (d + 2));
  char* fits_dev = (char*)sh;
  cudaStream_t st = h->s_compute;
  std::vector<char> fits((size_t)G * fsz);
  for (int g = 0; g < G; ++g) {
    double* p = base + per * g;
    double* A = p;      p += cA;
    double* D = p;      p += cD;
    double* F = p;      p += cF;
    double* Xs = p;     p += cX;
    double* a = p;      p += ca;
    double* e = p;      p += ce;
    double* y = p;      p += cy;
    double *U = nullptr, *W = nullptr;
    if (grad) {
      U = p;  p += cA;
      W = p;
    }
    BOPL_CUDA(h, cudaMemcpyAsync(e, ell_h + (size_t)g * d, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st));
    BOPL_CUDA(h, cudaMemcpyAsync(y, Yc_h + (size_t)g * n, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    fit_dev_fill(fits.data() + (size_t)g * fsz, A, D, F, U, W, nullptr, Xs, a, nullptr, e, y, logdiag_dev + (size_t)g * nblk,
                 quad_dev + g, info_dev + g, var_h[g], noise_h[g] + 1e-8, kern_kind_h[g]);
  }
  BOPL_CUDA(h, cudaMemcpyAsync(X_dev, X_h, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, st));
  BOPL_CUDA(h, cudaMemcpyAsync(fits_dev, fits.data(), fits.size(), cudaMemcpyHostToDevice, st));
  if ((rc = fit_factorize(h, fits_dev, G, X_dev, n, d, grad, grad, false, st))) return rc;
  if (grad && (rc = fit_lml_grad(h, fits_dev, G, n, d, partial_dev, out_dev, st))) return rc;
  std::vector<double> logdiag((size_t)G * nblk), quad(G);
  BOPL_CUDA(h, cudaMemcpyAsync(logdiag.data(), logdiag_dev, logdiag.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaMemcpyAsync(quad.data(), quad_dev, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaMemcpyAsync(info_h, info_dev, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (grad) BOPL_CUDA(h, cudaMemcpyAsync(grad_h, out_dev, (size_t)G * (d + 2) * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaStreamSynchronize(st));
  for (int g = 0; g < G; ++g) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += logdiag[(size_t)g * nblk + b];
    // exact_gaussian_inference.py:51: 0.5 (-n log 2pi - log|K| - y^T alpha)
    ll_h[g] = 0.5 * (-(double)n * 1.8378770664093454836 - 2.0 * s - quad[g]);
  }
  return BOPL_OK;
}

int bopl_get_factor(bopl_handle* h, int j, int hsel, double* L_h, double* alpha_h, double* Winv_h, double* Linv_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (j < 0 || j >= h->m || hsel < 0 || hsel >= h->H) return fail(h, BOPL_ERR_INVALID, "bad j / hsel");
  if (Winv_h && !h->has_winv) return fail(h, BOPL_ERR_STATE, "Winv was not loaded");
  if (Linv_h && !h->has_linv) return fail(h, BOPL_ERR_STATE, "Linv was not loaded");
  BOPL_CUDA(h, cudaDeviceSynchronize());
  const GpDev& gp = h->gp_host[(size_t)j * h->H + hsel];
  const int n = h->n, n_pad = h->n_pad, pad = h->pad;
  if (L_h) {
    std::vector<double> Lp((size_t)n_pad * n_pad);
    BOPL_CUDA(h, cudaMemcpy(Lp.data(), gp.Lneg, Lp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
      const int R = pad + i, bi = R / TRSM_BM;
      for (int k = 0; k < n; ++k) {
        const int C = pad + k;
        double v = (k <= i) ? Lp[(size_t)R * n_pad + C] : 0.0;
        if (k <= i && C / TRSM_BM < bi) v = -v;      // off-diagonal blocks are stored negated
        L_h[(size_t)i * n + k] = v;
      }
    }
  }
  if (alpha_h) BOPL_CUDA(h, cudaMemcpy(alpha_h, gp.alpha + pad, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  if (Winv_h) BOPL_CUDA(h, cudaMemcpy(Winv_h, gp.Winv, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
  if (Linv_h) BOPL_CUDA(h, cudaMemcpy(Linv_h, gp.Linv, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
  return BOPL_OK;
}

}  // extern "C"
