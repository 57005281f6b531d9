// uei_kernels.cu — kernel (3) of the hot path and its siblings:
//   incumbent_kernel     best[h,l] = max_i U(F^h[:,i], theta_l)                       (uEI.py:77,112,153)
//   uei_mc_kernel        fused reparameterised samples -> utility -> improvement -> MC/theta reduction
//                        (uEI.py:72-81,105-115,56-59), one warp per candidate, warp-shuffle reduction
//   uei_mc_grad_kernel   the same + analytic gradient (uEI.py:152-164,126-131)
//   uei_affine_kernel    closed-form EI of a linear utility (+ gradient)                 (uEI_affine.py:73-142)
// The posterior samples are mu + sigma * Z_k with Z the caller's common random numbers (uEI.py:31,172): nothing is
// drawn on the device.  All reductions run in a fixed order => results are bit-stable run to run.
#include "common.cuh"

namespace bopl {

namespace {

constexpr int MC_NT = 256;
constexpr int MC_WARPS = MC_NT / 32;

// U(a, theta_l) for a in registers.  Utilities: experiments/test_dtlz1a_linear.py:51-55, test_dtlz2a_quad.py:51-57,
// test_vlmop3_exp.py:41-52.
template <int KIND, int M = 0>
__device__ __forceinline__ double utility_value(const double* a, const double* th, int m) {
  double u = 0.0;
  if (M > 0) {   // compile-time attribute count: no predicated dead iterations in the (sample, theta) inner loop
    if (KIND == BOPL_UTIL_LINEAR) {
#pragma unroll
      for (int j = 0; j < M; ++j) u = fma(th[j], a[j], u);
    } else if (KIND == BOPL_UTIL_QUADRATIC) {
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double df = a[j] - th[j];
        u = fma(df, df, u);
      }
      u = -u;
    } else {
      const double t0 = th[0];
#pragma unroll
      for (int j = 0; j < M; ++j) u += 1.0 - exp_any(-t0 * a[j]);
      u = u * (1.0 / (t0 * (double)M));
    }
    return u;
  }
  if (KIND == BOPL_UTIL_CONSTRAINED) {
    bool ok = true;
#pragma unroll
    for (int j = 1; j < MAX_M; ++j)
      if (j < m) ok = ok && (th[j - 1] < a[j]);
    return ok ? a[0] : -INFINITY;
  } else if (KIND == BOPL_UTIL_LINEAR) {
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) u = fma(th[j], a[j], u);
  } else if (KIND == BOPL_UTIL_QUADRATIC) {
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) {
        double df = a[j] - th[j];
        u = fma(df, df, u);
      }
    u = -u;
  } else {
    const double t0 = th[0];
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) u += 1.0 - exp_any(-t0 * a[j]);
    u = u * (1.0 / (t0 * (double)m));
  }
  return u;
}

// dU/dy_j
template <int KIND>
__device__ __forceinline__ double utility_grad(double aj, const double* th, int m, int j) {
  if (KIND == BOPL_UTIL_LINEAR) return th[j];
  if (KIND == BOPL_UTIL_QUADRATIC) return -2.0 * (aj - th[j]);
  return exp_any(-th[0] * aj) / (double)m;
}

struct McParams {
  const double* mean;   // [m*N]
  const double* var;    // [m*N]
  const double* dmean;  // [m*N*d] (grad kernel)
  const double* dvar;   // [m*N*d]
  const double* Z;      // [nw*m]
  const double* theta;  // [Lt*p]
  const double* wts;    // [Lt]
  const double* best;   // [Lt]
  double* acq;          // [N]
  double* dacq;         // [N*d]
  double scale;
  int N, m, d, nw, Lt, p, accumulate;
  int plain;            // 1: expected utility, sum of U itself (no incumbent, no max(.,0)); acquisition_optimizer.py:211-247
};

// Shared staging: Z [nw*m], theta [Lt*p], wts [Lt], best [Lt]
__device__ __forceinline__ void stage_small(const McParams& p, double* zs, double* ths, double* ws, double* bs) {
  for (int i = threadIdx.x; i < p.nw * p.m; i += blockDim.x) zs[i] = p.Z[i];
  for (int i = threadIdx.x; i < p.Lt * p.p; i += blockDim.x) ths[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.Lt; i += blockDim.x) {
    ws[i] = p.wts[i];
    bs[i] = p.best ? p.best[i] : 0.0;
  }
  __syncthreads();
}

// One warp per candidate.  The 32 lanes form a 4 x 8 grid over (sample group, theta group): a lane keeps TL utility parameters
// (with their weights and incumbents) in REGISTERS and walks the samples k = kq, kq + 4, ...; per sample it forms the posterior
// draw a = mu + sigma * Z_k once and evaluates its TL utilities.  The first version read theta / weight / incumbent from shared
// memory for every (sample, theta) pair — five 8-byte shared loads per six fp64 operations, i.e. bound by the shared-memory
// crossbar at a third of the FP64 pipe (2.4 ms per launch at the benchmark shape); now it is three broadcast loads per TL pairs.
// Lanes then sum in a fixed order (per-lane partial sums, butterfly shuffle): bit-stable run to run.
template <int KIND, int M>
__global__ void __launch_bounds__(MC_NT, (M > 0 && M <= 4) ? 2 : 1) uei_mc_kernel(McParams p) {
  constexpr int MM = (M > 0) ? M : MAX_M;
  constexpr int TL = (M > 0 && M <= 4) ? 4 : 2;                       // utility parameters per lane and pass
  constexpr int LQ = 8, KQ = 4;                                        // lane = kq * 8 + lq
  constexpr int PT = (KIND == BOPL_UTIL_EXPONENTIAL) ? 2 : MM;         // registers per parameter: (-theta, 1 / (theta m)) or theta[m]
  extern __shared__ double sm[];
  double* zs = sm;
  double* ths = zs + p.nw * p.m;
  double* ws = ths + p.Lt * p.p;
  double* bs = ws + p.Lt;
  stage_small(p, zs, ths, ws, bs);
  const int lane = threadIdx.x & 31, lq = lane & (LQ - 1), kq = lane / LQ;
  const int warp_global = blockIdx.x * MC_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * MC_WARPS;
  const int m = (M > 0) ? M : p.m;
  // the next candidate's posterior is fetched while the current one is being evaluated: with one warp per candidate and four warps per
  // scheduler the load + sqrt latency at the top of every candidate was exposed
  double mu_n[MM], var_n[MM];
#pragma unroll
  for (int j = 0; j < MM; ++j)
    if (j < m && warp_global < p.N) {
      mu_n[j] = p.mean[(size_t)j * p.N + warp_global];
      var_n[j] = p.var[(size_t)j * p.N + warp_global];
    }
  for (int c = warp_global; c < p.N; c += nwarps) {
    double mu[MM], sg[MM];
#pragma unroll
    for (int j = 0; j < MM; ++j)
      if (j < m) {
        mu[j] = mu_n[j];
        sg[j] = sqrt(var_n[j]);
      }
    const int cn = c + nwarps;
#pragma unroll
    for (int j = 0; j < MM; ++j)
      if (j < m && cn < p.N) {
        mu_n[j] = p.mean[(size_t)j * p.N + cn];
        var_n[j] = p.var[(size_t)j * p.N + cn];
      }
    double acc = 0.0;
    for (int l0 = 0; l0 < p.Lt; l0 += LQ * TL) {
      double th[TL][PT], wl[TL], bl[TL], part[TL];
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        const int l = l0 + lq * TL + t;
        const bool valid = l < p.Lt;                    // a parameter past the end: weight 0 on a harmless theta
        if (KIND == BOPL_UTIL_EXPONENTIAL) {
          const double t0 = valid ? ths[l] : 1.0;
          th[t][0] = -t0;
          th[t][1] = 1.0 / (t0 * (double)m);
        } else {
#pragma unroll
          for (int j = 0; j < MM; ++j)
            if (j < PT && j < m) th[t][j] = valid ? ths[l * p.p + j] : 0.0;
        }
        wl[t] = valid ? ws[l] : 0.0;
        bl[t] = valid ? bs[l] : 0.0;
        part[t] = 0.0;
      }
      for (int k = kq; k < p.nw; k += KQ) {
        double a[MM];
#pragma unroll
        for (int j = 0; j < MM; ++j)
          if (j < m) a[j] = fma(sg[j], zs[k * m + j], mu[j]);
#pragma unroll
        for (int t = 0; t < TL; ++t) {
          double u = 0.0;
          if (KIND == BOPL_UTIL_LINEAR) {
#pragma unroll
            for (int j = 0; j < MM; ++j)
              if (j < m) u = fma(th[t][j < PT ? j : 0], a[j], u);
          } else if (KIND == BOPL_UTIL_QUADRATIC) {
#pragma unroll
            for (int j = 0; j < MM; ++j)
              if (j < m) {
                const double df = a[j] - th[t][j < PT ? j : 0];
                u = fma(df, df, u);
              }
            u = -u;
          } else {
#pragma unroll
            for (int j = 0; j < MM; ++j)
              if (j < m) u += 1.0 - exp_any(th[t][0] * a[j]);
            u = u * th[t][1];
          }
          part[t] = fma(wl[t], p.plain ? u : fmax(u - bl[t], 0.0), part[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < TL; ++t) acc += part[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      double v = acc * p.scale;
      p.acq[c] = p.accumulate ? p.acq[c] + v : v;
    }
  }
}

template <int KIND>
void launch_mc_m(const McParams& mp, int grid, size_t smem, cudaStream_t st) {
  switch (mp.m) {
    case 1: uei_mc_kernel<KIND, 1><<<grid, MC_NT, smem, st>>>(mp); break;
    case 2: uei_mc_kernel<KIND, 2><<<grid, MC_NT, smem, st>>>(mp); break;
    case 3: uei_mc_kernel<KIND, 3><<<grid, MC_NT, smem, st>>>(mp); break;
    case 4: uei_mc_kernel<KIND, 4><<<grid, MC_NT, smem, st>>>(mp); break;
    case 5: uei_mc_kernel<KIND, 5><<<grid, MC_NT, smem, st>>>(mp); break;
    case 6: uei_mc_kernel<KIND, 6><<<grid, MC_NT, smem, st>>>(mp); break;
    default: uei_mc_kernel<KIND, 0><<<grid, MC_NT, smem, st>>>(mp); break;
  }
}

// Value + gradient.  Per candidate the lanes also accumulate, per attribute j,
//   cm[j] = sum w_l [U > best_l] U'_j            (coefficient of d mu_j / dx)
//   cv[j] = sum w_l [U > best_l] U'_j * 0.5 Z_kj / sigma_j   (coefficient of d var_j / dx)   uEI.py:160-164
// and the warp then forms dacq[q] = scale * sum_j (cm[j] dmu[j][q] + cv[j] dvar[j][q]), lanes over q.
template <int KIND>
__global__ void __launch_bounds__(MC_NT) uei_mc_grad_kernel(McParams p) {
  extern __shared__ double sm[];
  double* zs = sm;
  double* ths = zs + p.nw * p.m;
  double* ws = ths + p.Lt * p.p;
  double* bs = ws + p.Lt;
  stage_small(p, zs, ths, ws, bs);
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * MC_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * MC_WARPS;
  const int m = p.m, d = p.d, npairs = p.nw * p.Lt;
  for (int c = warp_global; c < p.N; c += nwarps) {
    double mu[MAX_M], sg[MAX_M], cm[MAX_M], cv[MAX_M];
#pragma unroll
    for (int j = 0; j < MAX_M; ++j) {
      cm[j] = 0.0;
      cv[j] = 0.0;
      if (j < m) {
        mu[j] = p.mean[(size_t)j * p.N + c];
        sg[j] = sqrt(p.var[(size_t)j * p.N + c]);
      }
    }
    double acc = 0.0;
    for (int pr = lane; pr < npairs; pr += 32) {
      const int k = pr / p.Lt, l = pr - k * p.Lt;
      double a[MAX_M];
#pragma unroll
      for (int j = 0; j < MAX_M; ++j)
        if (j < m) a[j] = fma(sg[j], zs[k * m + j], mu[j]);
      const double* th = ths + l * p.p;
      double u = utility_value<KIND>(a, th, m);
      acc = fma(ws[l], p.plain ? u : fmax(u - bs[l], 0.0), acc);
      if (p.plain || u > bs[l]) {   // strict, uEI.py:160
        const double w = ws[l];
#pragma unroll
        for (int j = 0; j < MAX_M; ++j)
          if (j < m) {
            double gj = w * utility_grad<KIND>(a[j], th, m, j);
            cm[j] += gj;
            cv[j] = fma(gj, 0.5 * zs[k * m + j] / sg[j], cv[j]);
          }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          cm[j] += __shfl_xor_sync(0xffffffffu, cm[j], o);
          cv[j] += __shfl_xor_sync(0xffffffffu, cv[j], o);
        }
      }
    if (lane == 0) {
      double v = acc * p.scale;
      p.acq[c] = p.accumulate ? p.acq[c] + v : v;
    }
    for (int q = lane; q < d; q += 32) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < MAX_M; ++j)
        if (j < m) {
          size_t o = ((size_t)j * p.N + c) * d + q;
          g = fma(cm[j], p.dmean[o], g);
          g = fma(cv[j], p.dvar[o], g);
        }
      g *= p.scale;
      size_t o = (size_t)c * d + q;
      p.dacq[o] = p.accumulate ? p.dacq[o] + g : g;
    }
  }
}

// Closed-form EI of theta . f (uEI_affine.py).  One thread per candidate, loop over theta.
// value path (dacq == nullptr): sigma floored at 1e-10, Phi = 0.5 erfc(-u/sqrt2)   (uEI_affine.py:127-142, 84-87)
// gradient path: no floor; (mu-best) Phi + sigma phi; d = dmu Phi + phi dsigma      (uEI_affine.py:99-113)
__global__ void __launch_bounds__(128) uei_affine_kernel(McParams p) {
  extern __shared__ double sm[];
  double* ths = sm;
  double* ws = ths + p.Lt * p.m;
  double* bs = ws + p.Lt;
  for (int i = threadIdx.x; i < p.Lt * p.m; i += blockDim.x) ths[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.Lt; i += blockDim.x) {
    ws[i] = p.wts[i];
    bs[i] = p.best[i];
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.N) return;
  const int m = p.m, d = p.d;
  const bool with_grad = p.dacq != nullptr;
  const double inv_sqrt2pi = 0.39894228040143267794, inv_sqrt2 = 0.70710678118654752440;
  double mu[MAX_M], vr[MAX_M];
#pragma unroll
  for (int j = 0; j < MAX_M; ++j)
    if (j < m) {
      mu[j] = p.mean[(size_t)j * p.N + c];
      vr[j] = p.var[(size_t)j * p.N + c];
    }
  double acc = 0.0;
  double cm[MAX_M], cv[MAX_M];
#pragma unroll
  for (int j = 0; j < MAX_M; ++j) cm[j] = cv[j] = 0.0;
  for (int l = 0; l < p.Lt; ++l) {
    const double* th = ths + l * m;
    double mm = 0.0, vv = 0.0;
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) {
        mm = fma(th[j], mu[j], mm);
        vv = fma(th[j] * th[j], vr[j], vv);
      }
    double s = sqrt(vv);
    if (!with_grad) {
      // the floor is local to _get_quantiles (scalar branch, uEI_affine.py:137-138): the outer sigma stays un-floored
      double sf = (s < 1e-10) ? 1e-10 : s;
      double u = (mm - bs[l]) / sf;
      double phi = exp(-0.5 * u * u) * inv_sqrt2pi;
      double Phi = 0.5 * erfc(-u * inv_sqrt2);
      acc = fma(ws[l], s * (u * Phi + phi), acc);
    } else {
      double u = (mm - bs[l]) / s;
      double phi = exp(-0.5 * u * u) * inv_sqrt2pi;
      double Phi = 0.5 * erfc(-u * inv_sqrt2);
      acc = fma(ws[l], (mm - bs[l]) * Phi + s * phi, acc);
#pragma unroll
      for (int j = 0; j < MAX_M; ++j)
        if (j < m) {
          cm[j] = fma(ws[l] * Phi, th[j], cm[j]);
          cv[j] = fma(ws[l] * phi * 0.5 / s, th[j] * th[j], cv[j]);
        }
    }
  }
  double v = acc * p.scale;
  p.acq[c] = p.accumulate ? p.acq[c] + v : v;
  if (with_grad) {
    for (int q = 0; q < d; ++q) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < MAX_M; ++j)
        if (j < m) {
          size_t o = ((size_t)j * p.N + c) * d + q;
          g = fma(cm[j], p.dmean[o], g);
          g = fma(cv[j], p.dvar[o], g);
        }
      g *= p.scale;
      size_t o = (size_t)c * d + q;
      p.dacq[o] = p.accumulate ? p.dacq[o] + g : g;
    }
  }
}

// Closed-form constrained EI (uEI_constrained.py:49-73 value, :92-126,149-164 gradient).  One thread per candidate.
// Per theta_l: EI0 = (mu0 - best) Phi0 + s0 phi0 (norm.pdf / norm.cdf, no floor), P = prod_j Phi_j, value = EI0 * P.
// Gradient by the product rule (the reference's recursive aux_function_for_gradient unrolled):
//   d = P (Phi0 dmu0 + phi0 (0.5/s0) dvar0) + EI0 sum_j phi_j Q_j (dmu_j / s_j - 0.5 (mu_j - theta_j) / var_j^1.5 dvar_j),
//   Q_j = prod_{k != j} Phi_k.
__global__ void __launch_bounds__(128) uei_constrained_kernel(McParams p) {
  extern __shared__ double sm[];
  const int pc = p.m - 1;
  double* ths = sm;
  double* ws = ths + p.Lt * pc;
  double* bs = ws + p.Lt;
  for (int i = threadIdx.x; i < p.Lt * pc; i += blockDim.x) ths[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.Lt; i += blockDim.x) {
    ws[i] = p.wts[i];
    bs[i] = p.best[i];
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.N) return;
  const int m = p.m, d = p.d;
  const bool with_grad = p.dacq != nullptr;
  const double inv_sqrt2pi = 0.39894228040143267794, inv_sqrt2 = 0.70710678118654752440;
  double mu[MAX_M], sg[MAX_M], vr[MAX_M], cm[MAX_M], cv[MAX_M];
#pragma unroll
  for (int j = 0; j < MAX_M; ++j) {
    cm[j] = cv[j] = 0.0;
    if (j < m) {
      mu[j] = p.mean[(size_t)j * p.N + c];
      vr[j] = p.var[(size_t)j * p.N + c];
      sg[j] = sqrt(vr[j]);
    }
  }
  double acc = 0.0;
  for (int l = 0; l < p.Lt; ++l) {
    const double* th = ths + l * pc;
    const double u0 = (mu[0] - bs[l]) / sg[0];
    const double phi0 = exp(-0.5 * u0 * u0) * inv_sqrt2pi;
    const double Phi0 = 0.5 * erfc(-u0 * inv_sqrt2);
    const double ei0 = (mu[0] - bs[l]) * Phi0 + sg[0] * phi0;
    double Phi[MAX_M], phi[MAX_M];
    double P = 1.0;
#pragma unroll
    for (int j = 1; j < MAX_M; ++j)
      if (j < m) {
        const double u = (mu[j] - th[j - 1]) / sg[j];
        Phi[j] = 0.5 * erfc(-u * inv_sqrt2);
        phi[j] = exp(-0.5 * u * u) * inv_sqrt2pi;
        P *= Phi[j];
      }
    acc = fma(ws[l], ei0 * P, acc);
    if (with_grad) {
      cm[0] = fma(ws[l] * P, Phi0, cm[0]);
      cv[0] = fma(ws[l] * P, phi0 * 0.5 / sg[0], cv[0]);
#pragma unroll
      for (int j = 1; j < MAX_M; ++j)
        if (j < m) {
          double Q = 1.0;
#pragma unroll
          for (int k = 1; k < MAX_M; ++k)
            if (k < m && k != j) Q *= Phi[k];
          const double w = ws[l] * ei0 * phi[j] * Q;
          cm[j] = fma(w, 1.0 / sg[j], cm[j]);
          cv[j] = fma(w, -0.5 * (mu[j] - th[j - 1]) / (vr[j] * sg[j]), cv[j]);
        }
    }
  }
  double v = acc * p.scale;
  p.acq[c] = p.accumulate ? p.acq[c] + v : v;
  if (with_grad) {
    for (int q = 0; q < d; ++q) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < MAX_M; ++j)
        if (j < m) {
          size_t o = ((size_t)j * p.N + c) * d + q;
          g = fma(cm[j], p.dmean[o], g);
          g = fma(cv[j], p.dvar[o], g);
        }
      g *= p.scale;
      size_t o = (size_t)c * d + q;
      p.dacq[o] = p.accumulate ? p.dacq[o] + g : g;
    }
  }
}

// best[h*Lt + l] = max_i U(F[h][:, i], theta_l);  F [H][m][n].  One block per (l, h).
template <int KIND>
__global__ void __launch_bounds__(256) incumbent_kernel(const double* __restrict__ F, int m, int n, const double* __restrict__ theta,
                                                         int Lt, int p, double* __restrict__ best) {
  __shared__ double red[8];
  const int l = blockIdx.x, hh = blockIdx.y;
  const double* Fh = F + (size_t)hh * m * n;
  double th[MAX_M];
#pragma unroll
  for (int j = 0; j < MAX_M; ++j) th[j] = (j < p) ? theta[l * p + j] : 0.0;
  double best_v = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) {
    double a[MAX_M];
#pragma unroll
    for (int j = 0; j < MAX_M; ++j)
      if (j < m) a[j] = Fh[(size_t)j * n + i];
    best_v = fmax(best_v, utility_value<KIND>(a, th, m));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best_v = fmax(best_v, __shfl_xor_sync(0xffffffffu, best_v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best_v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = red[0];
    for (int w = 1; w < 8; ++w) b = fmax(b, red[w]);
    best[hh * Lt + l] = b;
  }
}

}  // namespace

static int mc_grid(const bopl_handle* h, int N) {
  long long warps_needed = N;
  long long blocks = (warps_needed + MC_WARPS - 1) / MC_WARPS;
  long long cap = (long long)h->sms * 8;
  return (int)std::max<long long>(1, std::min(blocks, cap));
}

int launch_uei_mc(bopl_handle* h, const double* mean, const double* var, int N, const double* Z, int nw, int util_kind,
                  const double* theta, int Lt, int p, const double* wts, const double* best, double scale,
                  int accumulate, double* acq, cudaStream_t st, int plain) {
  if (N <= 0) return BOPL_OK;
  McParams mp{};
  mp.plain = plain;
  mp.mean = mean; mp.var = var; mp.Z = Z; mp.theta = theta; mp.wts = wts; mp.best = best; mp.acq = acq;
  mp.scale = scale; mp.N = N; mp.m = h->m; mp.d = h->d; mp.nw = nw; mp.Lt = Lt; mp.p = p; mp.accumulate = accumulate;
  size_t smem = (size_t)(nw * h->m + Lt * p + 2 * Lt) * sizeof(double);
  if (smem > 48 * 1024) return fail(h, BOPL_ERR_UNSUPPORTED, "uei_mc: nw*m + Lt*(p+2) exceeds 6144 doubles of shared staging");
  int grid = mc_grid(h, N);
  switch (util_kind) {
    case BOPL_UTIL_LINEAR: launch_mc_m<BOPL_UTIL_LINEAR>(mp, grid, smem, st); break;
    case BOPL_UTIL_QUADRATIC: launch_mc_m<BOPL_UTIL_QUADRATIC>(mp, grid, smem, st); break;
    case BOPL_UTIL_EXPONENTIAL: launch_mc_m<BOPL_UTIL_EXPONENTIAL>(mp, grid, smem, st); break;
    default: return fail(h, BOPL_ERR_INVALID, "unknown utility kind");
  }
  BOPL_LAUNCH_CHECK(h, "uei_mc_kernel");
  return BOPL_OK;
}

int launch_uei_mc_grad(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                       int N, const double* Z, int nw, int util_kind, const double* theta, int Lt, int p,
                       const double* wts, const double* best, double scale, int accumulate, double* acq, double* dacq,
                       cudaStream_t st, int plain) {
  if (N <= 0) return BOPL_OK;
  McParams mp{};
  mp.plain = plain;
  mp.mean = mean; mp.var = var; mp.dmean = dmean; mp.dvar = dvar; mp.Z = Z; mp.theta = theta; mp.wts = wts; mp.best = best;
  mp.acq = acq; mp.dacq = dacq;
  mp.scale = scale; mp.N = N; mp.m = h->m; mp.d = h->d; mp.nw = nw; mp.Lt = Lt; mp.p = p; mp.accumulate = accumulate;
  size_t smem = (size_t)(nw * h->m + Lt * p + 2 * Lt) * sizeof(double);
  if (smem > 48 * 1024) return fail(h, BOPL_ERR_UNSUPPORTED, "uei_mc_grad: nw*m + Lt*(p+2) exceeds 6144 doubles of shared staging");
  int grid = mc_grid(h, N);
  switch (util_kind) {
    case BOPL_UTIL_LINEAR: uei_mc_grad_kernel<BOPL_UTIL_LINEAR><<<grid, MC_NT, smem, st>>>(mp); break;
    case BOPL_UTIL_QUADRATIC: uei_mc_grad_kernel<BOPL_UTIL_QUADRATIC><<<grid, MC_NT, smem, st>>>(mp); break;
    case BOPL_UTIL_EXPONENTIAL: uei_mc_grad_kernel<BOPL_UTIL_EXPONENTIAL><<<grid, MC_NT, smem, st>>>(mp); break;
    default: return fail(h, BOPL_ERR_INVALID, "unknown utility kind");
  }
  BOPL_LAUNCH_CHECK(h, "uei_mc_grad_kernel");
  return BOPL_OK;
}

int launch_uei_affine(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                      int N, const double* theta, int Lt, const double* wts, const double* best, double scale,
                      int accumulate, double* acq, double* dacq, cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  McParams mp{};
  mp.mean = mean; mp.var = var; mp.dmean = dmean; mp.dvar = dvar; mp.theta = theta; mp.wts = wts; mp.best = best;
  mp.acq = acq; mp.dacq = dacq;
  mp.scale = scale; mp.N = N; mp.m = h->m; mp.d = h->d; mp.Lt = Lt; mp.p = h->m; mp.accumulate = accumulate;
  size_t smem = (size_t)(Lt * h->m + 2 * Lt) * sizeof(double);
  if (smem > 48 * 1024) return fail(h, BOPL_ERR_UNSUPPORTED, "uei_affine: Lt*(m+2) exceeds 6144 doubles of shared staging");
  uei_affine_kernel<<<(N + 127) / 128, 128, smem, st>>>(mp);
  BOPL_LAUNCH_CHECK(h, "uei_affine_kernel");
  return BOPL_OK;
}

int launch_uei_constrained(bopl_handle* h, const double* mean, const double* var, const double* dmean, const double* dvar,
                           int N, const double* theta, int Lt, const double* wts, const double* best, double scale,
                           int accumulate, double* acq, double* dacq, cudaStream_t st) {
  if (N <= 0) return BOPL_OK;
  McParams mp{};
  mp.mean = mean; mp.var = var; mp.dmean = dmean; mp.dvar = dvar; mp.theta = theta; mp.wts = wts; mp.best = best;
  mp.acq = acq; mp.dacq = dacq;
  mp.scale = scale; mp.N = N; mp.m = h->m; mp.d = h->d; mp.Lt = Lt; mp.p = h->m - 1; mp.accumulate = accumulate;
  size_t smem = (size_t)(Lt * (h->m - 1) + 2 * Lt) * sizeof(double);
  if (smem > 48 * 1024) return fail(h, BOPL_ERR_UNSUPPORTED, "uei_constrained: Lt*(m+1) exceeds 6144 doubles of shared staging");
  uei_constrained_kernel<<<(N + 127) / 128, 128, smem, st>>>(mp);
  BOPL_LAUNCH_CHECK(h, "uei_constrained_kernel");
  return BOPL_OK;
}

int launch_incumbent(bopl_handle* h, const double* F, int util_kind, const double* theta, int Lt, int p, double* best,
                     cudaStream_t st) {
  dim3 grid(Lt, h->H);
  switch (util_kind) {
    case BOPL_UTIL_LINEAR: incumbent_kernel<BOPL_UTIL_LINEAR><<<grid, 256, 0, st>>>(F, h->m, h->n, theta, Lt, p, best); break;
    case BOPL_UTIL_QUADRATIC: incumbent_kernel<BOPL_UTIL_QUADRATIC><<<grid, 256, 0, st>>>(F, h->m, h->n, theta, Lt, p, best); break;
    case BOPL_UTIL_EXPONENTIAL: incumbent_kernel<BOPL_UTIL_EXPONENTIAL><<<grid, 256, 0, st>>>(F, h->m, h->n, theta, Lt, p, best); break;
    case BOPL_UTIL_CONSTRAINED: incumbent_kernel<BOPL_UTIL_CONSTRAINED><<<grid, 256, 0, st>>>(F, h->m, h->n, theta, Lt, p, best); break;
    default: return fail(h, BOPL_ERR_INVALID, "unknown utility kind");
  }
  BOPL_LAUNCH_CHECK(h, "incumbent_kernel");
  return BOPL_OK;
}

}  // namespace bopl
