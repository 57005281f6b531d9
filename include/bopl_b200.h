/*
 * bopl_b200.h — C-ABI of the B200-native uEI hot path (libbopl_b200.so).
 *
 * This is the drop-in boundary for RaulAstudillo06/BOPL's data-parallel hot path: the multi-output GP
 * posterior and the uEI / uEI_affine acquisition over large candidate batches.  The reference is pure
 * Python/NumPy with no FFI of its own; each entry point below names the reference function(s) it replaces
 * (paths relative to /root/reference/bopl/, `GPy/`, `GPyOpt/` = aux_software/GPy, aux_software/GPyOpt).
 * The ctypes binding a maintainer would add is shown in INTEGRATION.md and shipped in bopl_b200/_lib.py.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *  - every function returns 0 on success or a negative bopl_status; the message is available through
 *    bopl_last_error(handle) (or bopl_last_error(NULL) when no handle exists yet).  No exceptions.
 *  - all arrays are fp64, row-major (C order).  Pointers are DEVICE pointers unless the parameter name ends
 *    in `_h` (host).  The caller owns every input/output buffer; the handle owns its copy of the model and
 *    its scratch.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - calls on one handle are not thread-safe; different handles are independent.  Calls on one stream are ordered by the stream;
 *    consecutive calls on DIFFERENT streams (the host-buffer entry points run on the library's own stream) are ordered by the
 *    library: the later stream waits for an event recorded behind the earlier call, because all calls share the handle's scratch.
 *  - there is no CPU fallback: without a CUDA device bopl_create fails.
 */
#ifndef BOPL_B200_H
#define BOPL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bopl_handle bopl_handle;

enum bopl_status {
  BOPL_OK = 0,
  BOPL_ERR_INVALID = -1,    /* bad argument */
  BOPL_ERR_CUDA = -2,       /* CUDA runtime error (message has the cudaGetErrorString text) */
  BOPL_ERR_STATE = -3,      /* call order: model not set, gradients requested without Winv, ... */
  BOPL_ERR_UNSUPPORTED = -4, /* shape outside what the kernels are built for */
  BOPL_ERR_NUMERIC = -5      /* covariance matrix not positive definite (the caller may retry with jitter, linalg.py:53-78) */
};

/* GPy kernel `.name` -> tag.  GPy/kern/src/stationary.py:516 (Mat52), :427 (Mat32), :581 (ExpQuad), rbf.py:22 (rbf). */
enum bopl_kernel_kind { BOPL_KERN_MAT52 = 0, BOPL_KERN_RBF = 1, BOPL_KERN_MAT32 = 2, BOPL_KERN_EXPQUAD = 3 };

/* Utility callables of the named experiment configs (Python callables in the reference, a tag here):
 * linear      U = theta . y                       experiments/test_dtlz1a_linear.py:51-55   (p = m)
 * quadratic   U = -sum_j (y_j - theta_j)^2        experiments/test_dtlz2a_quad.py:51-57     (p = m)
 * exponential U = (1/m) sum_j (1-exp(-theta y_j))/theta   experiments/test_vlmop3_exp.py:41-52 (p = 1) */
enum bopl_utility_kind { BOPL_UTIL_LINEAR = 0, BOPL_UTIL_QUADRATIC = 1, BOPL_UTIL_EXPONENTIAL = 2,
                         /* y_0 if theta_{j-1} < y_j for every j >= 1, else -inf: experiments/test_portfolio2.py:64-79 (p = m-1);
                          * accepted by bopl_incumbent only (uEI_constrained has its own closed-form entry point) */
                         BOPL_UTIL_CONSTRAINED = 3 };

/* bopl_posterior flags */
#define BOPL_VAR_INCLUDE_NOISE 1 /* add the Gaussian likelihood variance: GPy/core/gp.py:416-417, likelihoods/gaussian.py:110-111 */
#define BOPL_VAR_CLIP 2          /* clip at 1e-10: GPyOpt/models/gpmodel.py:214 */

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int bopl_create(bopl_handle** out, int device);
void bopl_destroy(bopl_handle* h);
const char* bopl_last_error(const bopl_handle* h);
/* SM count of the handle's device and the library's build tag ("sm_100a ..."). */
int bopl_device_sms(const bopl_handle* h);
const char* bopl_build_info(void);

/* Kernel behind bopl_posterior / bopl_uei* for batches above the small-batch limit (the reference has one NumPy path:
 * GPy/inference/latent_function_inference/posterior.py:299-320 — dtrtrs on the cached factor, then the column square-sum).
 *   BOPL_POSTERIOR_INT8_7 (default) blocked triangular solve whose n^2/2 update runs on the int8 tensor pipe (tcgen05.mma kind::i8,
 *                         TMEM accumulators): both operands are split exactly into 7 base-256 digits (55 fractional bits against a
 *                         power-of-two row scale — every bit of an fp64 mantissa), the 28 digit products with weight above 2^-56 are
 *                         accumulated exactly in s32 and recombined in fp64; K*, the 64-row diagonal solves, mean and variance are fp64
 *                         (posterior_ozaki.cu).  Passes the fp64 kernel's parity tolerances unchanged, including at candidates next
 *                         to training points (tests/test_gpu_low_variance.py).  Shapes it does not cover (d > 12, n > 16384) run the
 *                         fp64 kernel — a kernel choice inside this library, not a fallback path.
 *   BOPL_POSTERIOR_FP64   fp64 DMMA.8x8x4 blocked triangular solve (posterior_trsm_t.cu), 0.95 of the cuBLAS DGEMM peak.
 *   BOPL_POSTERIOR_INT8   opt-in REDUCED-PRECISION mode: 6 digits (47 fractional bits), 21 digit products.  Absolute error of the
 *                         variance ~2e-11 sigma^2 at n = 2000, i.e. NOT within 1e-9 relative where the variance is below ~1e-2 sigma^2.
 *   BOPL_POSTERIOR_INT8_AUTO  the 6-digit split where it is safe: every bopl_set_model / bopl_fit_model first builds the 6-digit images and
 *                         probes the model where an absolute error hurts most — up to 256 training points, every hyper-sample,
 *                         noiseless variance against the fp64 kernel, |v6 - v64| <= 1e-9 |v64| + 64 ulp(sigma^2) — and builds the
 *                         7-digit images instead when the probe fails (bopl_get_posterior_kernel then reports INT8_7).
 * Select BEFORE bopl_set_model / bopl_fit_model (the digit images of the factor are built at upload); switching to FP64 is
 * always allowed.  Environment default: BOPL_POSTERIOR=int8x7|fp64|int8|int8auto. */
#define BOPL_POSTERIOR_FP64 0
#define BOPL_POSTERIOR_INT8 6
#define BOPL_POSTERIOR_INT8_7 7
#define BOPL_POSTERIOR_INT8_AUTO 60
int bopl_set_posterior_kernel(bopl_handle* h, int kind);
int bopl_get_posterior_kernel(const bopl_handle* h);

/* ---- model state ----------------------------------------------------------------------------------------
 * Replaces the cached state of m x H `GPy.models.GPRegression` instances that the reference keeps as live
 * Python objects: GPyOpt/models/gpmodel.py:123-140 (model_instances[h]) per attribute of
 * models/multi_outputGP.py:46-54; the arrays are those of GPy/inference/latent_function_inference/
 * exact_gaussian_inference.py:46-65 and posterior.py:171-191.  All HOST pointers; copied in (synchronous).
 *   kern_kind_h [m*H]      bopl_kernel_kind per (attribute j, hyper-sample h), index j*H + h (same for all below)
 *   X_h         [n*d]      training inputs (shared by the attributes)
 *   ell_h       [m*H*d]    ARD lengthscales (repeat the scalar d times for a non-ARD kernel)
 *   var_h, noise_h, ymean_h [m*H]   kernel variance, Gaussian noise variance, normaliser mean
 *   L_h         [m*H*n*n]  lower Cholesky factor of K + (noise+1e-8) I, row-major (upper part ignored)
 *   alpha_h     [m*H*n]    woodbury vector
 *   Winv_h      [m*H*n*n]  woodbury inverse (symmetric) or NULL when gradients are never requested
 *   Linv_h      [m*H*n*n]  explicit inverse of the Cholesky factor (lower triangular, row-major; LAPACK dtrtri) or NULL.
 *                          When given, batches of <= 32 candidates (the L-BFGS inner loop of
 *                          optimization_services/acquisition_optimizer.py:112) are solved as matrix-vector products
 *                          instead of through the blocked many-right-hand-side kernel (microseconds instead of one
 *                          2.4 ms tile time at n = 2000).                                                         */
int bopl_set_model(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                   const double* ell_h, const double* var_h, const double* noise_h, const double* ymean_h,
                   const double* L_h, const double* alpha_h, const double* Winv_h, const double* Linv_h);

/* ---- model fit on the device ------------------------------------------------------------------
 * Same result as bopl_set_model, but the per-(attribute, hyper-sample) posterior state is COMPUTED on the device from the
 * data and the hyper-parameters: K = K(X,X) + (noise + 1e-8 + jitter) I, blocked Cholesky on the FP64 tensor pipe,
 * alpha = K^-1 (y - mean(y)), and (optionally) K^-1 and L^-1 — what the reference does on the host per model update in
 * GPy/core/gp.py:52-62,247-260 (mean-centring normaliser of this fork, util/normalizer.py:57-72),
 * inference/latent_function_inference/exact_gaussian_inference.py:44-51 (jitchol/dpotrf, dpotrs) and posterior.py:171-191
 * (dpotri), called from GPyOpt/models/gpmodel.py:96-107 for each of models/multi_outputGP.py:57-62's attributes.
 * All HOST pointers.  Y_h [m*n] raw targets of the m attributes; jitter_h [m*H] extra diagonal or NULL;
 * want_winv / want_linv as the Winv_h / Linv_h arguments of bopl_set_model; logdet_h [m*H] (log|K|) may be NULL.
 * Returns BOPL_ERR_NUMERIC when a matrix is not positive definite (the message names the instance); the handle then has
 * no model. */
int bopl_fit_model(bopl_handle* h, int n, int d, int m, int H, const int* kern_kind_h, const double* X_h,
                   const double* Y_h, const double* ell_h, const double* var_h, const double* noise_h,
                   const double* jitter_h, int want_winv, int want_linv, double* logdet_h);

/* Per-instance outcome of the last bopl_fit_model on this handle: info_h[j*H + h] = 0, or 1 + the index of the pivot that failed
 * (that matrix is not positive definite).  Lets the caller add jitter to the FAILING matrices only, as the reference's jitchol does
 * per matrix (GPy/util/linalg.py:53-78).  count = m*H.  HOST pointer. */
int bopl_fit_info(const bopl_handle* h, int count, int* info_h);

/* Log marginal likelihood of G independent exact GPs on the same inputs and its gradient with respect to the kernel
 * hyper-parameters — one likelihood evaluation of MAP / HMC hyper-parameter inference (GPyOpt/models/gpmodel.py:109-149 ->
 * GPy/core/gp.py:247-266: exact_gaussian_inference.py:44-65 for log p(y) and dL/dK = 0.5 (alpha alpha^T - K^-1);
 * kern/src/stationary.py:191-215,236-245 for dL/d variance and dL/d lengthscale; likelihoods/gaussian.py:64-71 for
 * dL/d noise).  Does not touch the handle's model.  All HOST pointers.
 *   kern_kind_h [G]; X_h [n*d]; Yc_h [G*n] CENTRED targets; ell_h [G*d]; var_h, noise_h [G] (1e-8 is added as the reference does)
 *   ll_h [G]; grad_h [G*(d+2)] = (d/d variance, d/d ell_0..d-1, d/d noise); info_h [G] = 0 or 1 + index of the failed pivot
 *   (ll and grad of such an instance are not meaningful). */
int bopl_log_marginal(bopl_handle* h, int n, int d, int G, const int* kern_kind_h, const double* X_h, const double* Yc_h,
                      const double* ell_h, const double* var_h, const double* noise_h, double* ll_h, double* grad_h,
                      int* info_h);

/* Read back the posterior state of attribute j under hyper-sample hsel in the reference's own layout (tests, checkpoints):
 * L_h [n*n] lower Cholesky factor (row-major, zeros above the diagonal), alpha_h [n], Winv_h [n*n], Linv_h [n*n].
 * Any pointer may be NULL; Winv / Linv must have been loaded or fitted.  HOST pointers. */
int bopl_get_factor(bopl_handle* h, int j, int hsel, double* L_h, double* alpha_h, double* Winv_h, double* Linv_h);

/* ---- kernel (1): cross-covariance ------------------------------------------------------------------------
 * K(Xc, X) for attribute j under hyper-sample hsel -> K [N*n].  Replaces Stationary.K / _scaled_dist /
 * K_of_r: GPy/kern/src/stationary.py:104-113,128-166,529-530,440-441,594-595; rbf.py:42-43.             */
int bopl_cross_cov(bopl_handle* h, const double* Xc, int N, int hsel, int j, double* K, void* stream);

/* ---- kernels (1)+(2): posterior mean and variance -----------------------------------------------------------
 * mean, var [m*N] (either may be NULL).  Replaces MultiOutputGP.posterior_mean / posterior_variance / predict /
 * predict_noiseless: models/multi_outputGP.py:91-141 -> GPyOpt/models/gpmodel.py:170-223 -> GPy/core/gp.py:286-326,
 * 380-435 -> posterior.py:268-320 (K(X,Xc), dtrtrs against the cached Cholesky factor, column square-sum).
 * flags: BOPL_VAR_INCLUDE_NOISE | BOPL_VAR_CLIP.                                                             */
int bopl_posterior(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, double* var, int flags,
                   void* stream);

/* Posterior mean only (no triangular solve), [m*N].  multi_outputGP.py:117-125. */
int bopl_posterior_mean(bopl_handle* h, const double* Xc, int N, int hsel, double* mean, void* stream);

/* Posterior mean at the evaluated (training) points for every hyper-sample: F [H*m*n].
 * multi_outputGP.py:127-131. */
int bopl_train_mean(bopl_handle* h, double* F, void* stream);

/* Gradients of posterior mean and variance wrt the candidate: dmean, dvar [m*N*d] (either may be NULL).
 * Replaces MultiOutputGP.posterior_mean_gradient / posterior_variance_gradient: multi_outputGP.py:225-246 ->
 * GPy/core/gp.py:438-490 -> stationary.py:247-254,312-342,227-234 (+ C loop stationary_utils.c:1-14).
 * Needs Winv for dvar. */
int bopl_posterior_grad(bopl_handle* h, const double* Xc, int N, int hsel, double* dmean, double* dvar,
                        void* stream);

/* ---- incumbents ---------------------------------------------------------------------------------------------
 * best[h*Lt + l] = max_i U(F^h[:, i], theta_l) for h < H.  theta [Lt*p].
 * uEI.py:77,112,153; uEI_affine.py:118-125 (linear). */
int bopl_incumbent(bopl_handle* h, int util_kind, const double* theta, int Lt, int p, double* best, void* stream);

/* ---- kernel (3): fused MC expected-utility improvement --------------------------------------------------------
 * acq[i] (+)= scale * sum_l wts[l] * sum_k max(U(mean[:,i] + sqrt(var[:,i]) * Z[k,:], theta_l) - best[l], 0)
 * given mean, var [m*N].  Z [nw*m], theta [Lt*p], wts [Lt], best [Lt].  accumulate != 0 adds into acq.
 * Replaces the h-loop body of uEI._marginal_acq / _parallel_acq_helper + the theta aggregation of _compute_acq:
 * sampling_policies/acquisition_functions/uEI.py:72-81,105-115,56-59.                                           */
int bopl_uei_mc(bopl_handle* h, const double* mean, const double* var, int N, const double* Z, int nw,
                int util_kind, const double* theta, int Lt, int p, const double* wts, const double* best,
                double scale, int accumulate, double* acq, void* stream);

/* Full batch acquisition value: for hh < nH: posterior under hyper-sample hh, then the MC kernel with
 * best + hh*Lt, scale 1/(nH*nw).  acq [N].  Replaces uEI._compute_acq: uEI.py:40-62 (batch and N==1 paths; the
 * caller chooses per-h or stale incumbents by what it puts in `best` [nH*Lt]).                                   */
int bopl_uei(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
             const double* theta, int Lt, int p, const double* wts, const double* best, int nH, double* acq,
             void* stream);

/* Value and analytic gradient: acq [N], dacq [N*d].  Replaces uEI._compute_acq_withGradients /
 * _marginal_acq_with_gradient: uEI.py:119-168 (strict `>` mask, b = dmu + 0.5 W/sigma * dvar). Needs Winv. */
int bopl_uei_grad(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
                  const double* theta, int Lt, int p, const double* wts, const double* best, int nH,
                  double* acq, double* dacq, void* stream);

/* Expected utility under the noiseless posterior (no incumbent, no improvement clamp, NOT normalised — the reference sums):
 *   val[i] = sum_{h < nH} sum_l wts[l] sum_k U(mean^h[:,i] + sqrt(var_noiseless^h[:,i]) * Z[k,:], theta_l),  dval [N*d] or NULL.
 * Replaces the objective of AcquisitionOptimizer._current_marginal_argmax: optimization_services/
 * acquisition_optimizer.py:211-247 (MC branch, Z = 50 fixed draws); the affine branch (:145-187) is the same call with
 * the linear utility and Z = one row of zeros. */
int bopl_expected_utility(bopl_handle* h, const double* Xc, int N, const double* Z, int nw, int util_kind,
                          const double* theta, int Lt, int p, const double* wts, int nH, double* val, double* dval,
                          void* stream);

/* Closed-form EI of a linear utility: acq [N], dacq [N*d] or NULL.  theta [Lt*m], best [nH*Lt].
 * value path (dacq == NULL): uEI_affine._marginal_acq + _get_quantiles, uEI_affine.py:73-89,127-142 (sigma floor 1e-10);
 * gradient path: uEI_affine._marginal_acq_with_gradient, uEI_affine.py:91-116 (no floor).                       */
int bopl_uei_affine(bopl_handle* h, const double* Xc, int N, const double* theta, int Lt, const double* wts,
                    const double* best, int nH, double* acq, double* dacq, void* stream);

/* Closed-form constrained EI: attribute 0 is the objective, attributes 1..m-1 are constraints y_j > theta_{l,j-1}:
 *   acq = (1/nH) sum_h sum_l wts[l] * EI(mean_0, sigma_0; best[h,l]) * prod_{j>=1} Phi((mean_j - theta_{l,j-1}) / sigma_j)
 * theta [Lt*(m-1)], best [nH*Lt]; acq [N], dacq [N*d] or NULL.  Replaces uEI_constrained._compute_acq /
 * _compute_acq_withGradients: sampling_policies/acquisition_functions/uEI_constrained.py:36-126,149-164. */
int bopl_uei_constrained(bopl_handle* h, const double* Xc, int N, const double* theta, int Lt, const double* wts,
                         const double* best, int nH, double* acq, double* dacq, void* stream);

/* ---- anchor selection ---------------------------------------------------------------------------------------
 * The k largest acq values of this shard, ties -> lowest index; idx are GLOBAL indices (local + index_offset).
 * val [k], idx [k] (device).  Replaces np.argsort(-acq)[:k]: GPyOpt/optimization/anchor_points_generator.py:58-62. */
int bopl_topk(bopl_handle* h, const double* acq, int N, int k, long long index_offset, double* val,
              long long* idx, void* stream);

/* ---- multi-GPU: all-gather of the best candidates only -------------------------------------------
 * Candidates shard row-wise over the GPUs of one box (one process and one handle per GPU); every rank scores its block and keeps its
 * local top-k (bopl_topk).  bopl_topk_allgather then makes every rank hold the GLOBAL top-k — what the reference's single
 * `np.argsort(scores)[:k]` over the whole batch returns (GPyOpt/optimization/anchor_points_generator.py:58-62): it packs k records
 * (value, global index, x[d]), runs ONE ncclAllGather of k (2 + d) 8-byte words per rank on `stream`, and merges with the same
 * deterministic kernel on every rank (value descending, ties -> lowest global index, NaN last).  Device pointers, no host
 * synchronisation.  k_local <= k: how many of this rank's records are real (a rank with fewer than k candidates sends fillers).
 *   val, idx [k_local]   this rank's bopl_topk output;   Xc: this rank's candidate block [N_local*d], index_offset its first global row
 *   val_all [k], idx_all [k], x_all [k*d] (x_all may be NULL)
 * `comm` is an ncclComm_t (as void*) spanning the ranks: the caller's own, or one made with the helpers below — rank 0 calls
 * bopl_nccl_unique_id, ships the 128 bytes to the other ranks by any means (torch.distributed broadcast in sharding.py), every rank
 * calls bopl_nccl_comm_init.  NCCL is bound at run time (libnccl.so.2; in a PyTorch process the copy torch loaded).
 * bopl_topk_allgather_host is the same with HOST records in and out (x_h [k_local*d]: the rows already gathered by the caller), for
 * callers of bopl_uei_host; it synchronises once.  bopl_topk_pack / bopl_topk_merge are the two device stages on their own (records
 * [R*(2+d)] as bopl_topk_pack lays them out), for callers that bring their own transport. */
int bopl_nccl_unique_id(void* id128_h);
int bopl_nccl_comm_init(bopl_handle* h, int nranks, int rank, const void* id128_h, void** comm_out);
int bopl_nccl_comm_destroy(void* comm);
int bopl_topk_allgather(bopl_handle* h, void* comm, const double* val, const long long* idx, int k_local, const double* Xc,
                        long long index_offset, int k, double* val_all, long long* idx_all, double* x_all, void* stream);
int bopl_topk_allgather_host(bopl_handle* h, void* comm, const double* val_h, const long long* idx_h, int k_local, const double* x_h,
                             int k, double* val_all_h, long long* idx_all_h, double* x_all_h);
int bopl_topk_pack(bopl_handle* h, const double* val, const long long* idx, int k_local, const double* Xc, long long index_offset,
                   int k, double* records, void* stream);
int bopl_topk_merge(bopl_handle* h, const double* records, int R, int k, double* val_all, long long* idx_all, double* x_all,
                    void* stream);

/* ---- host-buffer convenience (the reference-facing call: NumPy arrays in, NumPy arrays out) ------------------
 * Same as bopl_uei (+ bopl_topk when k > 0) with HOST buffers; host<->device copies happen inside, chunked and
 * overlapped with compute on the handle's own streams; returns after the results are on the host.
 * acq_h [N] may be NULL when only the top-k is wanted.  topk_* may be NULL when k == 0.                          */
int bopl_uei_host(bopl_handle* h, const double* Xc_h, int N, const double* Z_h, int nw, int util_kind,
                  const double* theta_h, int Lt, int p, const double* wts_h, const double* best_h, int nH,
                  double* acq_h, int k, long long index_offset, double* topk_val_h, long long* topk_idx_h);

/* Host-buffer value + gradient (the call shape of the L-BFGS inner loop, optimization_services/acquisition_optimizer.py:112 ->
 * GPyOpt/optimization/optimizer.py:182: one point, or one point per anchor in the lock-step driver): bopl_uei_grad with HOST
 * buffers, copies inside, one synchronisation.  acq_h [N], dacq_h [N*d]. */
int bopl_uei_grad_host(bopl_handle* h, const double* Xc_h, int N, const double* Z_h, int nw, int util_kind,
                       const double* theta_h, int Lt, int p, const double* wts_h, const double* best_h, int nH,
                       double* acq_h, double* dacq_h);

/* ---- utility Thompson sampling: one posterior sample path, grown point by point -------------------------------------
 * Replaces the objective of sampling_policies/uTS.py:40-50.  There, every evaluation draws f(x) from a copy of each
 * attribute's GP (models/multi_outputGP.py:286-290 -> GPyOpt/models/gpmodel.py:167-168: hyper-sample 0), appends the draw
 * to the copy's data and re-fits it (GPy/core/gp.py:191-227 set_XY: normaliser mean over ALL targets, Ky, dpotrf, dpotrs —
 * O((n+t)^3) per evaluation), so that later draws are conditioned on earlier ones.  Here the factor is GROWN by one row per
 * evaluation (rank-1 append, matrix-vector work only) and all m attributes advance in the same launches.
 *
 * bopl_path_begin   starts a path on hyper-sample hsel.  Y_h [m*n] HOST: the raw targets the model was fitted to.
 *                   capacity: expected number of evaluations (buffers double when exceeded; <= 0 -> 256).
 *                   Needs the inverse factor (Linv_h of bopl_set_model / want_linv of bopl_fit_model).
 * bopl_path_eval_host  one evaluation at x_h [d] (HOST) with the caller's standard-normal draws z_h [m] (common random
 *                   numbers: nothing is drawn on the device): mean_j, var_j = posterior at x given the data AND the path so
 *                   far (GPy/core/gp.py:833-857 `posterior_samples_f(x, size=1, full_cov=True)`: raw variance, no
 *                   likelihood noise), y_j = mean_j + sqrt(var_j) z_j, then (x, y) joins the path.  y_h [m]; mean_h,
 *                   var_h [m] may be NULL.  Returns BOPL_ERR_NUMERIC (and drops the path) if the grown covariance is not
 *                   positive definite.
 * bopl_path_length  evaluations so far, -1 without a path.   bopl_path_end  frees the path (also done by bopl_set_model). */
int bopl_path_begin(bopl_handle* h, int hsel, int capacity, const double* Y_h);
int bopl_path_eval_host(bopl_handle* h, const double* x_h, const double* z_h, double* y_h, double* mean_h, double* var_h);
int bopl_path_length(const bopl_handle* h);
int bopl_path_end(bopl_handle* h);

/* Roofline denominator of the int8-tensor-pipe posterior kernel, measured on this handle's device: every SM issues back-to-back
 * tcgen05.mma kind::i8 M=128 N=256 K=32 on shared-memory-resident operands for about target_ms milliseconds (<= 0: 200 ms).
 * *tops_out = int8 tera-ops per second (2 per multiply-add), *ms_out (may be NULL) = the timed launch.  Synchronises. */
int bopl_probe_int8_peak(bopl_handle* h, double target_ms, double* tops_out, double* ms_out);

/* Number of kernel launches issued through this handle since creation (bench.py reports the delta). */
long long bopl_launch_count(const bopl_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* BOPL_B200_H */
