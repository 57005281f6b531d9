"""Hyper-parameter inference for the attributes' GPs — what `GPyOpt/models/gpmodel.py:109-149` does with GPy:
MAP estimate (`model.optimize(max_iters=200)`), 1 % multiplicative jitter, then HMC (`GPy/inference/mcmc/hmc.py:30-68`:
n_burnin + n_samples * subsample_interval iterations of `leapfrog_steps` leap-frog steps, Metropolis accept), keeping
every `subsample_interval`-th draw after burn-in as the hyper-samples.

Two drivers over the same objective:
  * `fit_hyperparameter_samples`          one attribute, likelihood and gradient on the host (NumPy + LAPACK);
  * `fit_hyperparameter_samples_device`   all attributes in LOCK-STEP, every likelihood + gradient evaluation of the m chains
                                          is ONE batched device call (`bopl_log_marginal`: covariance, blocked Cholesky,
                                          K^-1 and the n^2 gradient contraction on the GPU).  At n = 2000 one evaluation
                                          costs ~5 ms for all attributes instead of ~0.45 s per attribute on 16 host cores.
Neither is a bit-for-bit restatement of GPy's optimiser — MCMC output is a random draw in both.  They exist so
that `MultiOutputGP.updateModel(X, Y)` works without pre-supplied hyper-samples, i.e. so that a BO loop written against
the reference runs end to end.  Same model as the reference's default: Matern52 ARD (or the named kernel), variance and
lengthscales positive through GPy's Logexp transform (softplus), Gamma(E=2, V=4) priors on every positive
hyper-parameter (gpmodel.py:71-72), Gaussian noise fixed at 1e-6 (`exact_feval`), fixed at `noise_var`, or free and
initialised at 0.01 Var(Y) (gpmodel.py:67,75-80), mean-centred targets (this fork's normaliser).
"""
import numpy as np
from scipy.linalg import lapack
from scipy.special import gammaln

_SQRT3, _SQRT5 = np.sqrt(3.0), np.sqrt(5.0)
_GAMMA_A, _GAMMA_B = 1.0, 0.5                      # Gamma.from_EV(2, 4): shape E^2/V, rate E/V


def _softplus(x):
    return np.logaddexp(0.0, x)


def _softplus_inv(t):
    return np.where(t > 30.0, t, np.log(np.expm1(np.minimum(t, 30.0))))


def _kernel_and_grads(kind, variance, ell, X):
    """K(X, X) and dK/d(variance), dK/d(ell_q) for the stationary kernels of the path (GPy/kern/src/stationary.py)."""
    n, d = X.shape
    Xs = X / ell
    sq = np.sum(Xs * Xs, 1)
    r2 = np.clip(sq[:, None] + sq[None, :] - 2.0 * Xs @ Xs.T, 0.0, np.inf)
    np.fill_diagonal(r2, 0.0)
    r = np.sqrt(r2)
    if kind == "Mat52":
        e = np.exp(-_SQRT5 * r)
        K = variance * (1.0 + _SQRT5 * r + 5.0 / 3.0 * r2) * e
        dK_dr_over_r = variance * (-5.0 / 3.0) * (1.0 + _SQRT5 * r) * e           # (dK/dr) / r
    elif kind == "Mat32":
        e = np.exp(-_SQRT3 * r)
        K = variance * (1.0 + _SQRT3 * r) * e
        dK_dr_over_r = -3.0 * variance * e
    elif kind in ("rbf", "ExpQuad"):
        K = variance * np.exp(-0.5 * r2)
        dK_dr_over_r = -K
    else:
        raise ValueError("unsupported kernel %r" % (kind,))
    # dr/d ell_q = -(x_q - x'_q)^2 / (ell_q^3 r)  =>  dK/d ell_q = -(dK/dr)/r * (x_q - x'_q)^2 / ell_q^3
    dK_dell = []
    for q in range(d):
        dq = X[:, q][:, None] - X[:, q][None, :]
        dK_dell.append(-dK_dr_over_r * dq * dq / ell[q] ** 3)
    return K, K / variance, dK_dell


class GPObjective(object):
    """Negative log posterior of the hyper-parameters in the unconstrained (softplus-inverse) parametrisation."""

    def __init__(self, X, Y, kind="Mat52", ARD=True, noise_fixed=None):
        self.X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64).reshape(len(self.X), 1)
        self.Yc = Y - Y.mean()
        self.Yvar = float(Y.var())
        self.kind, self.ARD, self.noise_fixed = kind, ARD, noise_fixed
        self.n, self.d = self.X.shape
        self.n_ell = self.d if ARD else 1
        self.size = 1 + self.n_ell + (0 if noise_fixed is not None else 1)

    def unpack(self, x):
        th = _softplus(np.asarray(x, dtype=np.float64))
        variance = th[0]
        ell = th[1:1 + self.n_ell] if self.ARD else np.repeat(th[1], self.d)
        noise = self.noise_fixed if self.noise_fixed is not None else th[-1]
        return variance, ell, noise

    def initial(self):
        th = [1.0] + [1.0] * self.n_ell + ([] if self.noise_fixed is not None else [max(0.01 * self.Yvar, 1e-8)])
        return _softplus_inv(np.array(th))

    def host_likelihood(self, variance, ell, noise):
        """(log p(y), [d/d variance, d/d ell_0..d-1, d/d noise]) or (None, None) when K is not positive definite."""
        K, dK_dvar, dK_dell = _kernel_and_grads(self.kind, variance, ell, self.X)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += noise + 1e-8
        L, info = lapack.dpotrf(Ky, lower=1)
        if info != 0:
            return None, None
        alpha = lapack.dpotrs(L, self.Yc, lower=1)[0]
        Kinv = lapack.dpotri(L, lower=1)[0]
        Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
        ll = -0.5 * float((self.Yc.T @ alpha).item()) - np.sum(np.log(np.diag(L))) - 0.5 * self.n * np.log(2 * np.pi)
        dL_dK = 0.5 * (alpha @ alpha.T - Kinv)
        g = [np.sum(dL_dK * dK_dvar)] + [np.sum(dL_dK * dk) for dk in dK_dell] + [np.trace(dL_dK)]
        return ll, np.array(g)

    def value_and_grad(self, x):
        x = np.asarray(x, dtype=np.float64)
        variance, ell, noise = self.unpack(x)
        ll, g = self.host_likelihood(variance, ell, noise)
        return self.assemble(x, ll, g)

    def assemble(self, x, ll, g_full):
        """Negative log posterior and its gradient in the unconstrained parametrisation from the likelihood `ll` and its
        raw gradient `g_full` = (d/d variance, d/d ell_0..d-1, d/d noise); ll None = not positive definite."""
        x = np.asarray(x, dtype=np.float64)
        if ll is None or not np.isfinite(ll):
            return 1e25, np.zeros_like(x)
        th = _softplus(x)
        g_th = [g_full[0]]
        gl = list(g_full[1:1 + self.d])
        g_th += gl if self.ARD else [float(np.sum(gl))]
        if self.noise_fixed is None:
            g_th.append(g_full[1 + self.d])
        g_th = np.array(g_th)
        # Gamma(a, b) log prior on every positive hyper-parameter
        lp = np.sum(_GAMMA_A * np.log(_GAMMA_B) - gammaln(_GAMMA_A) + (_GAMMA_A - 1.0) * np.log(th) - _GAMMA_B * th)
        g_th = g_th + ((_GAMMA_A - 1.0) / th - _GAMMA_B)
        # ... and, as GPy does for every priored parameter behind a transformation (core/parameterization/priorizable.py:57-65,76-81),
        # the log-Jacobian of the Logexp transform: paramz 0.9.5 Logexp.log_jacobian(theta) = log(exp(theta) - 1) - theta
        # = log(1 - exp(-theta)), gradient 1 / (exp(theta) - 1).  Without it MAP and HMC target a different (and, as theta -> 0,
        # improper) density.
        with np.errstate(over="ignore"):
            lj = np.sum(np.log(-np.expm1(-th)))
            g_th = g_th + 1.0 / np.expm1(th)
        dth_dx = 1.0 / (1.0 + np.exp(-x))                       # derivative of softplus
        return -(ll + lp + lj), -(g_th * dth_dx)


def map_estimate(obj, max_iters=200):
    import scipy.optimize
    res = scipy.optimize.minimize(obj.value_and_grad, obj.initial(), jac=True, method="L-BFGS-B",
                                  options={"maxiter": max_iters})
    return res.x


def hmc_sample(obj, x0, num_samples, hmc_iters=20, stepsize=1e-1, rng=np.random):
    """GPy/inference/mcmc/hmc.py:30-68 with M = I."""
    x = np.array(x0, dtype=np.float64)
    out = np.empty((num_samples, x.size))
    f, g = obj.value_and_grad(x)
    for i in range(num_samples):
        p = rng.normal(size=x.size)
        H_old = f + 0.5 * float(p @ p)
        x_new, f_new, g_new = x.copy(), f, g
        for _ in range(hmc_iters):
            p = p - 0.5 * stepsize * g_new
            x_new = x_new + stepsize * p
            f_new, g_new = obj.value_and_grad(x_new)
            p = p - 0.5 * stepsize * g_new
        H_new = f_new + 0.5 * float(p @ p)
        if np.isfinite(H_new) and (H_old > H_new or rng.rand() < np.exp(H_old - H_new)):
            x, f, g = x_new, f_new, g_new
        out[i] = x
    return out


class DeviceLikelihood(object):
    """Batched likelihood + gradient of several attributes' GPs on the device: one `bopl_log_marginal` call per evaluation."""

    def __init__(self, engine, objs):
        self.engine, self.objs = engine, objs
        self.X = objs[0].X
        self.Yc = np.stack([o.Yc[:, 0] for o in objs])
        self.kinds = [o.kind for o in objs]
        self.calls = 0

    def value_and_grad(self, xs, active=None):
        """xs: one unconstrained vector per attribute -> list of (f, g).  `active`: indices to evaluate (default all)."""
        idx = list(range(len(self.objs))) if active is None else list(active)
        var, ell, noise, ok = [], [], [], []
        for i in idx:
            v, l, s = self.objs[i].unpack(xs[i])
            good = bool(np.isfinite(v) and v > 0 and np.all(np.isfinite(l)) and np.all(l > 0) and np.isfinite(s) and s >= 0)
            ok.append(good)
            var.append(v if good else 1.0)
            ell.append(np.asarray(l, dtype=np.float64) if good else np.ones(self.objs[i].d))
            noise.append(s if good else 1.0)
        ll, grad, info = self.engine.log_marginal(self.X, self.Yc[idx], [self.kinds[i] for i in idx], np.array(var), np.stack(ell),
                                                  np.array(noise))
        self.calls += 1
        out = []
        for k, i in enumerate(idx):
            bad = (not ok[k]) or info[k] != 0
            out.append(self.objs[i].assemble(xs[i], None if bad else float(ll[k]), None if bad else grad[k]))
        return out


def map_estimate_lockstep(lik, max_iters=200):
    """One unmodified SciPy L-BFGS-B per attribute, each on its own thread; the value+gradient requests of all still-running
    attributes are gathered into ONE batched device call (the lock-step driver of bopl_b200/optimization.py)."""
    import threading

    import scipy.optimize

    from .optimization import LockstepBatcher
    objs = lik.objs
    G = len(objs)
    width = max(o.size for o in objs)

    def f_df_batch(Xpad, ids):
        xs = {i: Xpad[k, :objs[i].size] for k, i in enumerate(ids)}
        res = lik.value_and_grad([xs.get(i) for i in range(G)], active=ids)
        f = np.array([r[0] for r in res])
        g = np.zeros((len(ids), width))
        for k, i in enumerate(ids):
            g[k, :objs[i].size] = res[k][1]
        return f, g

    batcher = LockstepBatcher(f_df_batch, G, with_ids=True)
    out, errors = [None] * G, []

    def work(i):
        try:
            def fun(x):
                xp = np.zeros(width)
                xp[:objs[i].size] = x
                f, g = batcher.call(i, xp)
                return f, g[:objs[i].size]
            out[i] = scipy.optimize.minimize(fun, objs[i].initial(), jac=True, method="L-BFGS-B", options={"maxiter": max_iters}).x
        except Exception as e:
            errors.append(e)
        finally:
            batcher.done(i)

    threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


def hmc_sample_lockstep(lik, x0s, num_samples, hmc_iters=20, stepsize=1e-1, rngs=None):
    """`hmc_sample` for every attribute at once: the chains advance together, one batched likelihood call per leap-frog step.
    Every chain draws from its own RandomState, so a chain's trajectory does not depend on the others."""
    G = len(x0s)
    rngs = rngs if rngs is not None else [np.random] * G
    xs = [np.array(x, dtype=np.float64) for x in x0s]
    outs = [np.empty((num_samples, x.size)) for x in xs]
    fg = lik.value_and_grad(xs)
    fs, gs = [r[0] for r in fg], [r[1] for r in fg]
    for it in range(num_samples):
        ps = [rngs[i].normal(size=xs[i].size) for i in range(G)]
        H_old = [fs[i] + 0.5 * float(ps[i] @ ps[i]) for i in range(G)]
        xn, fn, gn = [x.copy() for x in xs], list(fs), list(gs)
        for _ in range(hmc_iters):
            for i in range(G):
                ps[i] = ps[i] - 0.5 * stepsize * gn[i]
                xn[i] = xn[i] + stepsize * ps[i]
            fg = lik.value_and_grad(xn)
            for i in range(G):
                fn[i], gn[i] = fg[i]
                ps[i] = ps[i] - 0.5 * stepsize * gn[i]
        for i in range(G):
            H_new = fn[i] + 0.5 * float(ps[i] @ ps[i])
            if np.isfinite(H_new) and (H_old[i] > H_new or rngs[i].rand() < np.exp(H_old[i] - H_new)):
                xs[i], fs[i], gs[i] = xn[i], fn[i], gn[i]
            outs[i][it] = xs[i]
    return outs


def fit_hyperparameter_samples_device(engine, X, Ys, kinds, ARD, exact_feval, noise_var, n_samples=10, n_burnin=100,
                                      subsample_interval=10, step_size=1e-1, leapfrog_steps=20, max_iters=200, hmc=True, rngs=None):
    """All attributes at once on the device.  Returns variance (m,H), lengthscale (m,H,d), noise (m,H) — the same model, priors
    and sampler settings as `fit_hyperparameter_samples` (gpmodel.py:109-149)."""
    m = len(Ys)
    rngs = rngs if rngs is not None else [np.random] * m
    objs = [GPObjective(X, Ys[j], str(kinds[j]), bool(ARD[j]), 1e-6 if exact_feval[j] else noise_var[j]) for j in range(m)]
    lik = DeviceLikelihood(engine, objs)
    x_map = map_estimate_lockstep(lik, max_iters)
    if not hmc:
        xs = [x[None, :] for x in x_map]
    else:
        starts = []
        for j in range(m):
            th = _softplus(x_map[j]) * (1.0 + rngs[j].normal(size=x_map[j].size) * 0.01)      # gpmodel.py:124
            starts.append(_softplus_inv(th))
        ss = hmc_sample_lockstep(lik, starts, n_burnin + n_samples * subsample_interval, leapfrog_steps, step_size, rngs)
        xs = [s[n_burnin::subsample_interval][:n_samples] for s in ss]
    H = len(xs[0])
    d = objs[0].d
    variance, lengthscale, noise = np.empty((m, H)), np.empty((m, H, d)), np.empty((m, H))
    for j in range(m):
        for h in range(H):
            variance[j, h], lengthscale[j, h], noise[j, h] = objs[j].unpack(xs[j][h])
    return variance, lengthscale, noise


def fit_hyperparameter_samples(X, Y, kind="Mat52", ARD=True, exact_feval=False, noise_var=None, n_samples=10, n_burnin=100,
                               subsample_interval=10, step_size=1e-1, leapfrog_steps=20, max_iters=200, hmc=True, rng=np.random):
    """Returns (variance (H,), lengthscale (H,d), noise (H,)) — the hyper-samples of one attribute (gpmodel.py:109-149).
    hmc=False: the MAP estimate only (H = 1)."""
    noise_fixed = 1e-6 if exact_feval else noise_var
    obj = GPObjective(X, Y, kind, ARD, noise_fixed)
    x_map = map_estimate(obj, max_iters)
    if not hmc:
        xs = x_map[None, :]
    else:
        th = _softplus(x_map) * (1.0 + rng.normal(size=x_map.size) * 0.01)           # gpmodel.py:124
        ss = hmc_sample(obj, _softplus_inv(th), n_burnin + n_samples * subsample_interval, leapfrog_steps, step_size, rng)
        xs = ss[n_burnin::subsample_interval][:n_samples]
    unp = [obj.unpack(x) for x in xs]
    return (np.array([u[0] for u in unp]), np.array([u[1] for u in unp]), np.array([float(u[2]) for u in unp]))
