// path_kernels.cu — utility Thompson sampling: one posterior sample PATH grown point by point (sampling_policies/uTS.py:40-50).
//
// The reference's objective draws f(x) ~ N(mu(x), var(x)) from a copy of every attribute's GP, appends (x, f(x)) to the
// copy's data and RE-FITS it (`set_XY` -> mean-centring normaliser, Ky, dpotrf, dpotrs: O((n+t)^3) per evaluation), so
// that later draws are conditioned on the earlier ones.  Appending one point to an exact GP is a rank-1 growth of its
// Cholesky factor; here the grown factor is kept as
//        L' = [ L0  0 ]      L0 : the handle's factor (n x n, static, used through its explicit inverse Linv0)
//             [ B   C ]      B  : t x n, row i = (Linv0 k(X, x_i))^T;   C : t x t, held as its inverse Cinv
// and one evaluation is matrix-vector work only:   v0 = Linv0 k(X,x);  rhs = k(Xpath,x) - B v0;  u = Cinv rhs;
//   mu = [v0;u].(a - ybar b) + ybar   (a = L'^-1 Y', b = L'^-1 1, ybar = mean of ALL targets incl. the path: normalizer.py:57-72)
//   var = variance - |v0|^2 - |u|^2   (raw, no likelihood noise: gp.py:846 `_raw_predict(full_cov=True)`)
//   y = mu + sqrt(var) z              (np.random.multivariate_normal of a 1x1 covariance; z is the CALLER's draw)
//   append: l = sqrt(variance + noise + 1e-8 - |v0|^2 - |u|^2)  (exact_gaussian_inference.py:46-47 diagonal),
//           B[t] = v0, Cinv[t] = [-(u^T Cinv)/l, 1/l], a[t] = (y - [v0;u].a)/l, b[t] = (1 - [v0;u].b)/l.
// All attributes advance in the same launches (grid.y = attribute).  Six short launches per evaluation.
#include "common.cuh"

namespace bopl {

int launch_small_solve1(bopl_handle* h, const double* x_dev, int hsel, const double** K0, const double** V0, cudaStream_t st);

namespace {

struct PathAttr {
  double* B;     // [cap][n]
  double* Cinv;  // [cap][cap] lower triangular
  double* a;     // [n + cap]
  double* b;     // [n + cap]
};

struct PathState {
  int hsel = 0, cap = 0, t = 0, n = 0, m = 0, d = 0;
  PathAttr attr_host[MAX_M];
  PathAttr* attr_dev = nullptr;
  double* Xp = nullptr;     // [cap*d] path points (unscaled)
  double* rhs = nullptr;    // [m*cap]
  double* u = nullptr;      // [m*cap]
  double* scal = nullptr;   // [m*8]: 0 |v0|^2, 1 v0.a0, 2 v0.b0, 3 l, 4 sum of all targets
  double* io_dev = nullptr; // x[d] | z[m] | y[m] | mu[m] | var[m]
  double* io_pin = nullptr; // pinned mirror
  int* info_dev = nullptr;  // [m]
  int* info_pin = nullptr;
};

struct PathParams {
  const GpDev* gps;
  const PathAttr* attrs;
  const double* x;    // [d]
  const double* z;    // [m]
  const double* V0;   // [m][n]
  double* Xp;
  double* rhs;
  double* u;
  double* scal;
  double* out;        // y[m] | mu[m] | var[m]
  int* info;
  int H, hsel, m, d, n, t, cap;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// a0 = Linv0 Y_j, b0 = Linv0 1.  grid (ceil(n/8), m), one warp per row.
__global__ void __launch_bounds__(256) path_init_kernel(const GpDev* gps, const PathAttr* attrs, const double* Y, int H, int hsel,
                                                        int n) {
  const int j = blockIdx.y, lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const double* row = gps[j * H + hsel].Linv + (size_t)i * n;
  const double* y = Y + (size_t)j * n;
  double sa = 0.0, sb = 0.0;
  for (int k = lane; k <= i; k += 32) {
    const double l = row[k];
    sa = fma(l, y[k], sa);
    sb += l;
  }
  sa = warp_sum(sa);
  sb = warp_sum(sb);
  if (lane == 0) {
    attrs[j].a[i] = sa;
    attrs[j].b[i] = sb;
  }
}

// rows 0..t-1: rhs_i = k(x_i, x) - B_i . v0;  rows t, t+1, t+2: |v0|^2, v0.a0, v0.b0.  grid (ceil((t+3)/8), m).
__global__ void __launch_bounds__(256) path_rhs_kernel(PathParams p) {
  const int j = blockIdx.y, lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = p.n, t = p.t;
  if (row >= t + 3) return;
  const PathAttr at = p.attrs[j];
  const double* v0 = p.V0 + (size_t)j * n;
  const double* vec = row < t ? at.B + (size_t)row * n : row == t ? v0 : row == t + 1 ? at.a : at.b;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = lane;
  for (; k + 96 < n; k += 128) {
    s0 = fma(vec[k], v0[k], s0);
    s1 = fma(vec[k + 32], v0[k + 32], s1);
    s2 = fma(vec[k + 64], v0[k + 64], s2);
    s3 = fma(vec[k + 96], v0[k + 96], s3);
  }
  for (; k < n; k += 32) s0 = fma(vec[k], v0[k], s0);
  const double s = warp_sum((s0 + s1) + (s2 + s3));
  if (lane != 0) return;
  if (row < t) {
    const GpDev& gp = p.gps[j * p.H + p.hsel];
    double r2 = 0.0;
    for (int q = 0; q < p.d; ++q) {
      const double df = p.Xp[(size_t)row * p.d + q] / gp.ell[q] - p.x[q] / gp.ell[q];
      r2 = fma(df, df, r2);
    }
    p.rhs[(size_t)j * p.cap + row] = kern_value(gp.kind, gp.variance, r2) - s;
  } else {
    p.scal[j * 8 + (row - t)] = s;
  }
}

// u = Cinv rhs.  grid (ceil(t/8), m), one warp per row.
__global__ void __launch_bounds__(256) path_u_kernel(PathParams p) {
  const int j = blockIdx.y, lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= p.t) return;
  const double* row = p.attrs[j].Cinv + (size_t)i * p.cap;
  const double* rhs = p.rhs + (size_t)j * p.cap;
  double s = 0.0;
  for (int k = lane; k <= i; k += 32) s = fma(row[k], rhs[k], s);
  s = warp_sum(s);
  if (lane == 0) p.u[(size_t)j * p.cap + i] = s;
}

// The scalars of the step: mean, variance, the draw, the new diagonal entry, the new entries of a and b.  grid (m), 256 threads.
__global__ void __launch_bounds__(256) path_finish_kernel(PathParams p) {
  __shared__ double red[3][8];
  const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n, t = p.t;
  const PathAttr at = p.attrs[j];
  const double* u = p.u + (size_t)j * p.cap;
  double uu = 0.0, ua = 0.0, ub = 0.0;
  for (int i = threadIdx.x; i < t; i += 256) {
    const double ui = u[i];
    uu = fma(ui, ui, uu);
    ua = fma(ui, at.a[n + i], ua);
    ub = fma(ui, at.b[n + i], ub);
  }
  uu = warp_sum(uu);
  ua = warp_sum(ua);
  ub = warp_sum(ub);
  if (lane == 0) {
    red[0][warp] = uu;
    red[1][warp] = ua;
    red[2][warp] = ub;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  uu = ua = ub = 0.0;
  for (int w = 0; w < 8; ++w) {
    uu += red[0][w];
    ua += red[1][w];
    ub += red[2][w];
  }
  const GpDev& gp = p.gps[j * p.H + p.hsel];
  double* sc = p.scal + j * 8;
  const double vv = sc[0] + uu, va = sc[1] + ua, vb = sc[2] + ub;
  const double ybar = sc[4] / (double)(n + t);
  const double mu = (va - ybar * vb) + ybar;
  const double var = gp.variance - vv;
  // a 1x1 "covariance" that rounding pushed below zero: multivariate_normal factors it by SVD, i.e. uses |var|
  const double y = fma(sqrt(fabs(var)), p.z[j], mu);
  const double l2 = (gp.variance + (gp.noise + 1e-8)) - vv;
  const bool ok = l2 > 0.0;
  const double l = ok ? sqrt(l2) : 1.0;
  p.info[j] = ok ? 0 : 1;
  sc[3] = l;
  sc[4] += y;
  at.a[n + t] = (y - va) / l;
  at.b[n + t] = (1.0 - vb) / l;
  p.out[j] = y;
  p.out[p.m + j] = mu;
  p.out[2 * p.m + j] = var;
}

// New row of Cinv (columns 0..t), new row of B (= v0), the point itself.  grid (ceil((t+1)/32) + ceil(n/256), m).
// A Cinv block owns 32 columns (one per lane, 256-byte coalesced row segments); its 8 warps take the rows i = k0 + w, k0 + w + 8, ...
// and the 8 partial sums are added in a fixed order.
__global__ void __launch_bounds__(256) path_append_kernel(PathParams p) {
  __shared__ double part[8][32];
  const int j = blockIdx.y, n = p.n, t = p.t;
  const PathAttr at = p.attrs[j];
  const int cblocks = (t + 1 + 31) / 32;
  if ((int)blockIdx.x < cblocks) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.x * 32, k = k0 + lane;
    const double* u = p.u + (size_t)j * p.cap;
    double s = 0.0;
    if (k < t)
      for (int i = k0 + warp; i < t; i += 8)
        if (i >= k) s = fma(u[i], at.Cinv[(size_t)i * p.cap + k], s);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
      const double l = p.scal[j * 8 + 3];
      if (k < t) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += part[w][lane];
        at.Cinv[(size_t)t * p.cap + k] = -a / l;
      } else if (k == t) {
        at.Cinv[(size_t)t * p.cap + t] = 1.0 / l;
      }
    }
    if (j == 0 && blockIdx.x == 0 && threadIdx.x < p.d) p.Xp[(size_t)t * p.d + threadIdx.x] = p.x[threadIdx.x];
  } else {
    const int i = (blockIdx.x - cblocks) * 256 + threadIdx.x;
    if (i < n) at.B[(size_t)t * n + i] = p.V0[(size_t)j * n + i];
  }
}

void release(PathState* ps) {
  if (!ps) return;
  for (int j = 0; j < ps->m; ++j) {
    PathAttr& a = ps->attr_host[j];
    if (a.B) cudaFree(a.B);
    if (a.Cinv) cudaFree(a.Cinv);
    if (a.a) cudaFree(a.a);
    if (a.b) cudaFree(a.b);
  }
  void* bufs[] = {ps->attr_dev, ps->Xp, ps->rhs, ps->u, ps->scal, ps->io_dev, ps->info_dev};
  for (void* q : bufs)
    if (q) cudaFree(q);
  if (ps->io_pin) cudaFreeHost(ps->io_pin);
  if (ps->info_pin) cudaFreeHost(ps->info_pin);
  delete ps;
}

// (Re)allocate the capacity-dependent buffers, keeping the first ps->t rows.
int reserve(bopl_handle* h, PathState* ps, int cap, cudaStream_t st) {
  const int n = ps->n, m = ps->m, d = ps->d, t = ps->t;
  for (int j = 0; j < m; ++j) {
    PathAttr old = ps->attr_host[j], nw{};
    BOPL_CUDA(h, cudaMalloc((void**)&nw.B, (size_t)cap * n * sizeof(double)));
    ps->attr_host[j].B = nw.B;
    if (old.B) {
      BOPL_CUDA(h, cudaMemcpyAsync(nw.B, old.B, (size_t)t * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    BOPL_CUDA(h, cudaMalloc((void**)&nw.Cinv, (size_t)cap * cap * sizeof(double)));
    ps->attr_host[j].Cinv = nw.Cinv;
    if (old.Cinv && t > 0) {
      BOPL_CUDA(h, cudaMemcpy2DAsync(nw.Cinv, (size_t)cap * sizeof(double), old.Cinv, (size_t)ps->cap * sizeof(double),
                                     (size_t)t * sizeof(double), t, cudaMemcpyDeviceToDevice, st));
    }
    BOPL_CUDA(h, cudaMalloc((void**)&nw.a, (size_t)(n + cap) * sizeof(double)));
    ps->attr_host[j].a = nw.a;
    BOPL_CUDA(h, cudaMalloc((void**)&nw.b, (size_t)(n + cap) * sizeof(double)));
    ps->attr_host[j].b = nw.b;
    if (old.a) {
      BOPL_CUDA(h, cudaMemcpyAsync(nw.a, old.a, (size_t)(n + t) * sizeof(double), cudaMemcpyDeviceToDevice, st));
      BOPL_CUDA(h, cudaMemcpyAsync(nw.b, old.b, (size_t)(n + t) * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    BOPL_CUDA(h, cudaStreamSynchronize(st));
    if (old.B) cudaFree(old.B);
    if (old.Cinv) cudaFree(old.Cinv);
    if (old.a) cudaFree(old.a);
    if (old.b) cudaFree(old.b);
  }
  double* Xp = nullptr;
  BOPL_CUDA(h, cudaMalloc((void**)&Xp, (size_t)cap * d * sizeof(double)));
  if (ps->Xp) {
    BOPL_CUDA(h, cudaMemcpy(Xp, ps->Xp, (size_t)t * d * sizeof(double), cudaMemcpyDeviceToDevice));
    cudaFree(ps->Xp);
  }
  ps->Xp = Xp;
  if (ps->rhs) cudaFree(ps->rhs);
  if (ps->u) cudaFree(ps->u);
  ps->rhs = ps->u = nullptr;
  BOPL_CUDA(h, cudaMalloc((void**)&ps->rhs, (size_t)m * cap * sizeof(double)));
  BOPL_CUDA(h, cudaMalloc((void**)&ps->u, (size_t)m * cap * sizeof(double)));
  if (!ps->attr_dev) BOPL_CUDA(h, cudaMalloc((void**)&ps->attr_dev, MAX_M * sizeof(PathAttr)));
  BOPL_CUDA(h, cudaMemcpy(ps->attr_dev, ps->attr_host, MAX_M * sizeof(PathAttr), cudaMemcpyHostToDevice));
  ps->cap = cap;
  return BOPL_OK;
}

}  // namespace

void path_free(bopl_handle* h) {
  release((PathState*)h->path);
  h->path = nullptr;
}

}  // namespace bopl

using namespace bopl;

extern "C" {

int bopl_path_begin(bopl_handle* h, int hsel, int capacity, const double* Y_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  if (hsel < 0 || hsel >= h->H) return fail(h, BOPL_ERR_INVALID, "hsel out of range");
  if (!Y_h) return fail(h, BOPL_ERR_INVALID, "null buffer");
  if (!h->has_linv) return fail(h, BOPL_ERR_STATE, "the sample path needs the inverse factor (Linv_h / want_linv)");
  path_free(h);
  PathState* ps = new PathState();
  h->path = ps;
  ps->hsel = hsel;
  ps->n = h->n;
  ps->m = h->m;
  ps->d = h->d;
  for (int j = 0; j < MAX_M; ++j) ps->attr_host[j] = PathAttr{};
  cudaStream_t st = h->s_compute;
  const int m = h->m, n = h->n, d = h->d;
  int rc = reserve(h, ps, capacity > 0 ? capacity : 256, st);
  if (rc) return rc;
  BOPL_CUDA(h, cudaMalloc((void**)&ps->scal, (size_t)m * 8 * sizeof(double)));
  BOPL_CUDA(h, cudaMalloc((void**)&ps->io_dev, (size_t)(d + 4 * m) * sizeof(double)));
  BOPL_CUDA(h, cudaMalloc((void**)&ps->info_dev, (size_t)m * sizeof(int)));
  BOPL_CUDA(h, cudaMallocHost((void**)&ps->io_pin, (size_t)(d + 4 * m) * sizeof(double)));
  BOPL_CUDA(h, cudaMallocHost((void**)&ps->info_pin, (size_t)m * sizeof(int)));
  // targets -> a0, b0, sum
  std::vector<double> sc((size_t)m * 8, 0.0);
  for (int j = 0; j < m; ++j) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += Y_h[(size_t)j * n + i];
    sc[(size_t)j * 8 + 4] = s;
  }
  BOPL_CUDA(h, cudaMemcpyAsync(ps->scal, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  double* Yd = nullptr;
  BOPL_CUDA(h, cudaMalloc((void**)&Yd, (size_t)m * n * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(Yd, Y_h, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    path_init_kernel<<<dim3((n + 7) / 8, m), 256, 0, st>>>(h->gp_dev, ps->attr_dev, Yd, h->H, hsel, n);
    e = cudaGetLastError();
    h->launches++;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(Yd);
  if (e != cudaSuccess) return fail(h, BOPL_ERR_CUDA, std::string("path_init_kernel: ") + cudaGetErrorString(e));
  return BOPL_OK;
}

int bopl_path_eval_host(bopl_handle* h, const double* x_h, const double* z_h, double* y_h, double* mean_h, double* var_h) {
  BOPL_CHECK_HANDLE(h);
  BOPL_NEED_MODEL(h);
  PathState* ps = (PathState*)h->path;
  if (!ps) return fail(h, BOPL_ERR_STATE, "bopl_path_begin has not been called");
  if (!x_h || !z_h || !y_h) return fail(h, BOPL_ERR_INVALID, "null buffer");
  cudaStream_t st = h->s_compute;
  StreamOrder order_guard(h, st);
  int rc;
  if (ps->t == ps->cap && (rc = reserve(h, ps, 2 * ps->cap, st))) return rc;
  const int m = ps->m, d = ps->d, n = ps->n, t = ps->t;
  memcpy(ps->io_pin, x_h, (size_t)d * sizeof(double));
  memcpy(ps->io_pin + d, z_h, (size_t)m * sizeof(double));
  BOPL_CUDA(h, cudaMemcpyAsync(ps->io_dev, ps->io_pin, (size_t)(d + m) * sizeof(double), cudaMemcpyHostToDevice, st));
  const double *K0 = nullptr, *V0 = nullptr;
  if ((rc = launch_small_solve1(h, ps->io_dev, ps->hsel, &K0, &V0, st))) return rc;
  PathParams p{};
  p.gps = h->gp_dev;
  p.attrs = ps->attr_dev;
  p.x = ps->io_dev;
  p.z = ps->io_dev + d;
  p.V0 = V0;
  p.Xp = ps->Xp;
  p.rhs = ps->rhs;
  p.u = ps->u;
  p.scal = ps->scal;
  p.out = ps->io_dev + d + m;
  p.info = ps->info_dev;
  p.H = h->H;
  p.hsel = ps->hsel;
  p.m = m;
  p.d = d;
  p.n = n;
  p.t = t;
  p.cap = ps->cap;
  path_rhs_kernel<<<dim3((t + 3 + 7) / 8, m), 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "path_rhs_kernel");
  if (t > 0) {
    path_u_kernel<<<dim3((t + 7) / 8, m), 256, 0, st>>>(p);
    BOPL_LAUNCH_CHECK(h, "path_u_kernel");
  }
  path_finish_kernel<<<m, 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "path_finish_kernel");
  path_append_kernel<<<dim3((t + 1 + 31) / 32 + (n + 255) / 256, m), 256, 0, st>>>(p);
  BOPL_LAUNCH_CHECK(h, "path_append_kernel");
  BOPL_CUDA(h, cudaMemcpyAsync(ps->io_pin + d + m, ps->io_dev + d + m, (size_t)3 * m * sizeof(double), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaMemcpyAsync(ps->info_pin, ps->info_dev, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, st));
  BOPL_CUDA(h, cudaStreamSynchronize(st));
  for (int j = 0; j < m; ++j)
    if (ps->info_pin[j]) {
      path_free(h);
      return fail(h, BOPL_ERR_NUMERIC, "sample path: grown covariance of attribute " + std::to_string(j) +
                                           " is not positive definite at point " + std::to_string(t) + " (path dropped)");
    }
  memcpy(y_h, ps->io_pin + d + m, (size_t)m * sizeof(double));
  if (mean_h) memcpy(mean_h, ps->io_pin + d + 2 * m, (size_t)m * sizeof(double));
  if (var_h) memcpy(var_h, ps->io_pin + d + 3 * m, (size_t)m * sizeof(double));
  ps->t = t + 1;
  return BOPL_OK;
}

int bopl_path_length(const bopl_handle* h) {
  if (!h || !h->path) return -1;
  return ((const PathState*)h->path)->t;
}

int bopl_path_end(bopl_handle* h) {
  BOPL_CHECK_HANDLE(h);
  cudaStreamSynchronize(h->s_compute);
  path_free(h);
  return BOPL_OK;
}

}  // extern "C"
