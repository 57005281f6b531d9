"""Oracle: uEI / uEI_affine acquisition values and gradients (TEST INFRASTRUCTURE).

Two forms of each: `*_loop` follows the reference's Python loops statement by statement (small shapes
only); `*_vec` is the broadcast NumPy form used at benchmark sizes and as the CPU baseline.
References: bopl/sampling_policies/acquisition_functions/uEI.py and uEI_affine.py under /root/reference.

`model` is an `oracle.gp.MultiOutputGPOracle` (or anything with the same duck-typed surface).
"""
import numpy as np
from scipy.special import erfc
from scipy.stats import norm

UTILITY_KINDS = ("linear", "quadratic", "exponential", "constrained")


# ----------------------------------------------------------------------------- θ aggregation
def aggregate(marginal, use_full_support, prob_dist):
    """uEI.py:56-60 / uEI_affine.py:49-54."""
    if use_full_support:
        acq = np.matmul(marginal, prob_dist)
    else:
        acq = np.sum(marginal, axis=1) / marginal.shape[1]
    return np.reshape(acq, (marginal.shape[0], 1))


def aggregate_gradient(marginal_d, use_full_support, prob_dist):
    """uEI.py:126-133 / uEI_affine.py:63-71."""
    if use_full_support:
        d = np.tensordot(marginal_d, prob_dist, 1)
    else:
        d = np.sum(marginal_d, axis=2) / marginal_d.shape[2]
    return np.reshape(d, marginal_d.shape[:2])


# ----------------------------------------------------------------------------- uEI value, loop-faithful
def uei_marginal_loop(model, func, thetas, W_samples, X, n_h, incumbent="per_h"):
    """Marginal (per-θ) MC improvement, (N, L).

    incumbent="stale": uEI._marginal_acq (uEI.py:64-84) — F = mean at the evaluated points is taken ONCE
    before the h-loop, under whatever hyper-sample the model currently points at.
    incumbent="per_h": uEI._marginal_acq_parallel + _parallel_acq_helper (uEI.py:86-117) — F is recomputed
    inside the helper, i.e. under the current h.  (The pool is replaced by a plain loop; the first,
    discarded `pool.map` of :94 has no effect on the result.)"""
    X = np.atleast_2d(X)
    n_w = W_samples.shape[0]
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    if incumbent == "stale":
        fX_evaluated = model.posterior_mean_at_evaluated_points()
    for h in range(n_h):
        model.set_hyperparameters(h)
        if incumbent == "per_h":
            fX_evaluated = model.posterior_mean_at_evaluated_points()
        muX = model.posterior_mean(X)
        sigmaX = np.sqrt(model.posterior_variance(X))
        for l in range(L):
            max_val = np.max(func(fX_evaluated, thetas[l]))
            for W in W_samples:
                for i in range(X.shape[0]):
                    valx = func(muX[:, i] + sigmaX[:, i] * W, thetas[l])
                    marginal[i, l] += max(valx - max_val, 0)
    marginal /= (n_h * n_w)
    return marginal


def uei_marginal_with_gradient_loop(model, func, gradient, thetas, W_samples, X, n_h):
    """uEI._marginal_acq_with_gradient, uEI.py:136-168 (stale incumbent, strict `>` for the gradient mask)."""
    X = np.atleast_2d(X)
    fX_evaluated = model.posterior_mean_at_evaluated_points()
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    marginal_d = np.zeros((X.shape[0], X.shape[1], L))
    n_w = W_samples.shape[0]
    for h in range(n_h):
        model.set_hyperparameters(h)
        muX = model.posterior_mean(X)
        sigmaX = np.sqrt(model.posterior_variance(X))
        dmuX_dX = model.posterior_mean_gradient(X)
        dvar_dX = model.posterior_variance_gradient(X)
        for l in range(L):
            max_val = np.max(func(fX_evaluated, thetas[l]))
            for W in W_samples:
                for i in range(X.shape[0]):
                    a = muX[:, i] + sigmaX[:, i] * W
                    valx = func(a, thetas[l])
                    marginal[i, l] += max(valx - max_val, 0)
                    if valx > max_val:
                        b = np.multiply((0.5 * W / sigmaX[:, i]), dvar_dX[:, i, :].transpose()).transpose()
                        b += dmuX_dX[:, i, :]
                        marginal_d[i, :, l] += np.matmul(gradient(a, thetas[l]), b)
    marginal /= (n_h * n_w)
    marginal_d /= (n_h * n_w)
    return marginal, marginal_d


# ----------------------------------------------------------------------------- kind-tagged utilities (vector form)
def _U(kind, a, theta, m):
    """a: (m, ...) ; returns (...)  — same formulas as oracle.utilities / the experiment scripts."""
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    if kind == "linear":
        return np.tensordot(th, a, 1)
    if kind == "quadratic":
        return -np.sum(np.square(a - th.reshape((-1,) + (1,) * (a.ndim - 1))), axis=0)
    if kind == "exponential":
        return np.sum(1.0 - np.exp(-th[0] * a), axis=0) / th[0] / m
    raise ValueError(kind)


def _dU(kind, a, theta, m):
    """Gradient wrt y, shape of a."""
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    if kind == "linear":
        return np.broadcast_to(th.reshape((-1,) + (1,) * (a.ndim - 1)), a.shape)
    if kind == "quadratic":
        return -2 * (a - th.reshape((-1,) + (1,) * (a.ndim - 1)))
    if kind == "exponential":
        return np.exp(-th[0] * a) / m
    raise ValueError(kind)


def incumbents(model, kind, thetas, n_h, per_h=True):
    """best[h, l] = max_i U(F^h[:, i], θ_l)  (uEI.py:77,112,153; uEI_affine.py:118-125).
    per_h=False replicates the stale variant: every row is computed under the model's current h."""
    m = model.output_dim
    L = len(thetas)
    best = np.empty((n_h, L))
    h0 = getattr(model, "h", 0)
    for h in range(n_h):
        if per_h:
            model.set_hyperparameters(h)
        F = model.posterior_mean_at_evaluated_points()
        for l in range(L):
            best[h, l] = np.max(_U(kind, F, thetas[l], m))
    model.set_hyperparameters(h0)
    return best


# ----------------------------------------------------------------------------- uEI value, vectorised
def uei_marginal_vec(model, kind, thetas, W_samples, X, n_h, incumbent="per_h", chunk=8192, best=None):
    """Same numbers as `uei_marginal_loop` up to summation order; chunked over candidates."""
    X = np.atleast_2d(X)
    m = model.output_dim
    L = len(thetas)
    n_w = W_samples.shape[0]
    if best is None:
        best = incumbents(model, kind, thetas, n_h, per_h=(incumbent == "per_h"))
    marginal = np.zeros((X.shape[0], L))
    Wt = W_samples.T[:, :, None]                                    # (m, n_w, 1)
    for h in range(n_h):
        model.set_hyperparameters(h)
        for s in range(0, X.shape[0], chunk):
            Xs = X[s:s + chunk]
            mu = model.posterior_mean(Xs)
            sig = np.sqrt(model.posterior_variance(Xs))
            a = mu[:, None, :] + sig[:, None, :] * Wt               # (m, n_w, Nc)
            for l in range(L):
                val = _U(kind, a, thetas[l], m)                     # (n_w, Nc)
                marginal[s:s + chunk, l] += np.maximum(val - best[h, l], 0).sum(0)
    marginal /= (n_h * n_w)
    return marginal


def uei_marginal_with_gradient_vec(model, kind, thetas, W_samples, X, n_h, chunk=2048, best=None):
    """Vector form of uEI.py:136-168 (stale incumbent unless `best` (n_h, L) is given)."""
    X = np.atleast_2d(X)
    m = model.output_dim
    L = len(thetas)
    n_w = W_samples.shape[0]
    if best is None:
        best = incumbents(model, kind, thetas, n_h, per_h=False)
    marginal = np.zeros((X.shape[0], L))
    marginal_d = np.zeros((X.shape[0], X.shape[1], L))
    Wt = W_samples.T[:, :, None]
    for h in range(n_h):
        model.set_hyperparameters(h)
        for s in range(0, X.shape[0], chunk):
            Xs = X[s:s + chunk]
            mu = model.posterior_mean(Xs)
            sig = np.sqrt(model.posterior_variance(Xs))
            dmu = model.posterior_mean_gradient(Xs)                 # (m, Nc, d)
            dvar = model.posterior_variance_gradient(Xs)
            a = mu[:, None, :] + sig[:, None, :] * Wt
            for l in range(L):
                val = _U(kind, a, thetas[l], m)
                marginal[s:s + chunk, l] += np.maximum(val - best[h, l], 0).sum(0)
                mask = (val > best[h, l])[None, :, :]               # strict, uEI.py:160
                g = _dU(kind, a, thetas[l], m) * mask               # (m, n_w, Nc)
                c_mu = g.sum(1)                                     # (m, Nc)
                c_var = (g * (0.5 * Wt / sig[:, None, :])).sum(1)
                marginal_d[s:s + chunk, :, l] += np.einsum("jn,jnq->nq", c_mu, dmu) + np.einsum("jn,jnq->nq", c_var, dvar)
    marginal /= (n_h * n_w)
    marginal_d /= (n_h * n_w)
    return marginal, marginal_d


# ----------------------------------------------------------------------------- uEI_affine (closed form)
def _get_quantiles(fmax, m, s):
    """uEI_affine.py:127-142 (σ floored at 1e-10; erfc form of Φ)."""
    s = np.where(s < 1e-10, 1e-10, s)
    u = (m - fmax) / s
    phi = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    Phi = 0.5 * erfc(-u / np.sqrt(2))
    return phi, Phi, u


def affine_best_so_far(model, thetas):
    """uEI_affine.py:118-125 (under the model's CURRENT hyper-sample)."""
    mu_eval = model.posterior_mean_at_evaluated_points()
    return np.array([max(np.matmul(thetas[l], mu_eval)) for l in range(len(thetas))])


def uei_affine_marginal_loop(model, thetas, X, n_h):
    """uEI_affine._marginal_acq, uEI_affine.py:73-89."""
    X = np.atleast_2d(X)
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    for h in range(n_h):
        model.set_hyperparameters(h)
        meanX, varX = model.predict(X)
        best = affine_best_so_far(model, thetas)
        for l in range(L):
            for i in range(X.shape[0]):
                mu = np.dot(thetas[l], meanX[:, i])
                sigma = np.sqrt(np.dot(np.square(thetas[l]), varX[:, i]))
                phi, Phi, u = _get_quantiles(best[l], mu, sigma)
                marginal[i, l] += sigma * (u * Phi + phi)      # the floor inside _get_quantiles is local (scalar branch)
    return marginal / n_h


def uei_affine_marginal_vec(model, thetas, X, n_h):
    X = np.atleast_2d(X)
    thetas = np.asarray(thetas, dtype=np.float64)
    marginal = np.zeros((X.shape[0], len(thetas)))
    for h in range(n_h):
        model.set_hyperparameters(h)
        meanX, varX = model.predict(X)
        best = affine_best_so_far(model, thetas)
        mu = thetas @ meanX                                         # (L, N)
        sigma = np.sqrt(np.square(thetas) @ varX)
        phi, Phi, u = _get_quantiles(best[:, None], mu, sigma)
        marginal += (sigma * (u * Phi + phi)).T
    return marginal / n_h


def uei_affine_marginal_with_gradient_loop(model, thetas, X, n_h):
    """uEI_affine._marginal_acq_with_gradient, uEI_affine.py:91-116 (norm.pdf/cdf, no σ floor)."""
    X = np.atleast_2d(X)
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    marginal_d = np.zeros((X.shape[0], X.shape[1], L))
    for h in range(n_h):
        model.set_hyperparameters(h)
        meanX, varX = model.predict(X)
        dmean_dX = model.posterior_mean_gradient(X)
        dvar_dX = model.posterior_variance_gradient(X)
        best_all = affine_best_so_far(model, thetas)
        for l in range(L):
            best = best_all[l]
            for i in range(X.shape[0]):
                mu = np.dot(thetas[l], meanX[:, i])
                sigma = np.sqrt(np.dot(np.square(thetas[l]), varX[:, i]))
                phi = norm.pdf((mu - best) / sigma)
                Phi = norm.cdf((mu - best) / sigma)
                marginal[i, l] += (mu - best) * Phi + sigma * phi
                dmu_dX = np.matmul(thetas[l], dmean_dX[:, i, :])
                dsigma_dX = 0.5 * np.matmul(np.square(thetas[l]), dvar_dX[:, i, :]) / sigma
                marginal_d[i, :, l] += dmu_dX * Phi + phi * dsigma_dX
    return marginal / n_h, marginal_d / n_h


# ----------------------------------------------------------------------------- uEI_constrained (closed form)
def constrained_best_so_far(model, func, thetas):
    """uEI_constrained._marginal_max_feasible_val_so_far, uEI_constrained.py:136-146 (model's CURRENT hyper-sample)."""
    F = model.posterior_mean_at_evaluated_points()
    return np.array([np.max(func(F, thetas[l])) for l in range(len(thetas))])


def uei_constrained_marginal_loop(model, func, thetas, X, n_h):
    """uEI_constrained._marginal_acq, uEI_constrained.py:49-73."""
    X = np.atleast_2d(X)
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    best = constrained_best_so_far(model, func, thetas)
    for h in range(n_h):
        model.set_hyperparameters(h)
        meanX, varX = model.predict(X)
        sigmaX = np.sqrt(varX)
        for l in range(L):
            for i in range(X.shape[0]):
                mu, sigma = meanX[0, i], sigmaX[0, i]
                phi = norm.pdf((mu - best[l]) / sigma)
                Phi = norm.cdf((mu - best[l]) / sigma)
                val = (mu - best[l]) * Phi + sigma * phi
                for j in range(1, meanX.shape[0]):
                    val *= norm.cdf((meanX[j, i] - thetas[l][j - 1]) / sigmaX[j, i])
                marginal[i, l] += val
    return marginal / n_h


def _aux_function_for_gradient(mean, var, mean_gradient, var_gradient, theta):
    """uEI_constrained.py:149-164 (recursive product rule over the feasibility factors)."""
    j = len(mean)
    mu = mean[j - 1]
    sigma = np.sqrt(var[j - 1])
    output = norm.pdf((mu - theta[j - 1]) / sigma) * ((mean_gradient[j - 1, :] / sigma)
                                                      - (0.5 * (mu - theta[j - 1]) / np.power(var[j - 1], 1.5)) * var_gradient[j - 1, :])
    if j == 1:
        return output
    Phi_aux = norm.cdf((mu - theta[j - 1]) / sigma)
    for k in range(j - 1):
        output = output * norm.cdf((mean[k] - theta[k]) / np.sqrt(var[k]))
    output = output + Phi_aux * _aux_function_for_gradient(mean[:j - 1], var[:j - 1], mean_gradient[:j - 1, :],
                                                           var_gradient[:j - 1, :], theta[:j - 1])
    return output


def uei_constrained_marginal_with_gradient_loop(model, func, thetas, X, n_h):
    """uEI_constrained._marginal_acq_with_gradient, uEI_constrained.py:92-126."""
    X = np.atleast_2d(X)
    L = len(thetas)
    marginal = np.zeros((X.shape[0], L))
    marginal_d = np.zeros((X.shape[0], X.shape[1], L))
    best = constrained_best_so_far(model, func, thetas)
    for h in range(n_h):
        model.set_hyperparameters(h)
        meanX, varX = model.predict(X)
        sigmaX = np.sqrt(varX)
        dmean_dX = model.posterior_mean_gradient(X)
        dvar_dX = model.posterior_variance_gradient(X)
        for l in range(L):
            for i in range(X.shape[0]):
                mu, sigma = meanX[0, i], sigmaX[0, i]
                phi = norm.pdf((mu - best[l]) / sigma)
                Phi = norm.cdf((mu - best[l]) / sigma)
                val = (mu - best[l]) * Phi + sigma * phi
                val_aux = np.copy(val)
                gradient = dmean_dX[0, i, :] * Phi + phi * (0.5 / sigma) * dvar_dX[0, i, :]
                for j in range(1, meanX.shape[0]):
                    Phi = norm.cdf((meanX[j, i] - thetas[l][j - 1]) / sigmaX[j, i])
                    val *= Phi
                    gradient = gradient * Phi
                gradient = gradient + val_aux * _aux_function_for_gradient(meanX[1:, i], varX[1:, i], dmean_dX[1:, i, :],
                                                                           dvar_dX[1:, i, :], thetas[l])
                marginal[i, l] += val
                marginal_d[i, :, l] += gradient
    return marginal / n_h, marginal_d / n_h


# ----------------------------------------------------------------------------- marginal expected utility (baseline points)
def marginal_objective_loop(model, func, gradient, parameter, Z_samples, X, n_h, affine=False):
    """The val_func / val_func_with_gradient closures of AcquisitionOptimizer._current_marginal_argmax
    (bopl/optimization_services/acquisition_optimizer.py:145-187 affine, :211-247 MC): returns (-value (N,1), -gradient (N,d))."""
    X = np.atleast_2d(X)
    m = model.output_dim
    func_val = np.zeros((X.shape[0], 1))
    func_gradient = np.zeros(X.shape)
    for h in range(n_h):
        model.set_hyperparameters(h)
        if affine:
            muX = model.posterior_mean(X)
            dmu_dX = model.posterior_mean_gradient(X)
            func_val += np.reshape(np.matmul(parameter, muX), (X.shape[0], 1))
            func_gradient += np.tensordot(parameter, dmu_dX, axes=1)
            continue
        mean, var = model.predict_noiseless(X)
        std = np.sqrt(var)
        dmean_dX = model.posterior_mean_gradient(X)
        dstd_dX = model.posterior_variance_gradient(X)
        for i in range(X.shape[0]):
            for j in range(m):
                dstd_dX[j, i, :] /= (2 * std[j, i])
            for Z in Z_samples:
                aux1 = mean[:, i] + np.multiply(Z, std[:, i])
                func_val[i, 0] += func(aux1, parameter)
                aux2 = dmean_dX[:, i, :] + np.multiply(dstd_dX[:, i, :].T, Z).T
                func_gradient[i, :] += np.matmul(gradient(aux1, parameter), aux2)
    return -func_val, -func_gradient


# ----------------------------------------------------------------------------- anchor selection
def top_k(neg_acq, k):
    """GPyOpt/optimization/anchor_points_generator.py:58-62 on scores = -acq (acquisitions/base.py:40).
    The reference's np.argsort is an unstable introsort (tie order unspecified); the oracle fixes
    kind='stable' = lowest index first among exact ties."""
    scores = np.asarray(neg_acq).flatten()
    idx = np.argsort(scores, kind="stable")[:min(len(scores), k)]
    return idx, scores[idx]
