"""Caller of the hot path: the acquisition optimizer (SURVEY.md §8 a14 + §8f-1).

Drop-in for `bopl/optimization_services/acquisition_optimizer.py:16-137` (AcquisitionOptimizer.optimize) and the pieces
of `GPyOpt/optimization/optimizer.py` it uses (`OptLbfgs` :160-194, `apply_optimizer` :264-302) for continuous box
domains:

    candidates (random design, n_starting) -> batch acquisition scores -> n_anchor best -> L-BFGS-B from every anchor
    -> best optimized point.

Two things differ from the reference, neither changes a number (the final re-evaluation of the rounded optima, which ranks them, is
done one point per call exactly as `apply_optimizer` does — uEI's single-point and batch branches use different incumbents when H > 1):
  * scoring + anchor selection is one fused GPU call when `f` is a bopl_b200 acquisition (anchor_points.py);
  * the per-anchor L-BFGS-B runs (the reference does them one after the other, each calling `f_df` with ONE point,
    acquisition_optimizer.py:112) run in LOCK-STEP: every anchor keeps its own, unmodified `scipy.optimize.fmin_l_bfgs_b`
    (same `factr=1e6`, `maxiter`, bounds), on its own thread, and the `f_df` requests of all still-running anchors are
    gathered into one batched GPU call.  A candidate's value and gradient do not depend on what else is in the batch
    (the kernels are per-candidate deterministic), so every trajectory is bit-identical to the serial loop while the
    number of GPU round trips drops by the number of anchors.  `batched=False` gives the reference's serial loop.
"""
import threading

import numpy as np

from .anchor_points import BoxSpace, ObjectiveAnchorPointsGenerator


class ContextManager(object):
    """acquisition_optimizer.py:287-331 reduced to a context-free continuous box."""

    def __init__(self, space, context=None):
        if context:
            raise NotImplementedError("context variables are not on the accelerated path")
        self.space = space
        self.noncontext_bounds = space.get_bounds()

    def _expand_vector(self, x):
        return np.atleast_2d(x)


class OptLbfgs(object):
    """GPyOpt/optimization/optimizer.py:160-194."""

    def __init__(self, bounds, maxiter=300):
        self.bounds = bounds
        self.maxiter = maxiter

    def optimize(self, x0, f=None, df=None, f_df=None, verbose=False, maxfevals=1000):
        import scipy.optimize
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
        if f_df is None and df is not None:
            f_df = lambda x: (float(np.squeeze(f(x))), np.asarray(df(x)).reshape(-1))
        if f_df is None:
            res = scipy.optimize.fmin_l_bfgs_b(lambda x: float(np.squeeze(f(x))), x0=x0, bounds=self.bounds,
                                               approx_grad=True, maxiter=self.maxiter)
        else:
            def fun(x):
                fx, g = f_df(x)
                return float(np.squeeze(fx)), np.asarray(g, dtype=np.float64).reshape(-1)
            res = scipy.optimize.fmin_l_bfgs_b(fun, x0=x0, bounds=self.bounds, maxiter=self.maxiter, factr=1e6)
        task = res[2]['task']
        task = task.decode() if isinstance(task, bytes) else str(task)
        if 'ABNORMAL' in task:                                   # optimizer.py:185-187: report x0, f(x0)
            return np.atleast_2d(x0), (np.atleast_2d(f(np.atleast_2d(x0))) if f is not None else None)
        return np.atleast_2d(res[0]), np.atleast_2d(res[1])


def cma_es_minimize(f, x0, sigma0, lower, upper, maxfevals=1000, rng=None, tolfun=1e-11, tolx=1e-11):
    """(mu/mu_w, lambda)-CMA-ES with the default strategy parameters of Hansen's tutorial ("The CMA Evolution Strategy: A
    Tutorial", 2016, Table 1 / Algorithm 1): lambda = 4 + floor(3 ln d), log-rank weights, cumulative step-size
    adaptation, rank-one + rank-mu covariance update, eigen-decomposition refreshed lazily.  Box bounds: candidates are
    sampled unconstrained, EVALUATED at their projection onto the box, and ranked with a quadratic distance penalty.
    This stands in for `cma.fmin(f, x0, 0.6, {"bounds": ..., "maxfevals": ...})` (pycma, absent from this image; used by
    GPyOpt/optimization/optimizer.py:229-261) — same algorithm family and budget, not the same random stream."""
    rng = np.random if rng is None else rng
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    lower, upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
    d = x0.size
    lam = 4 + int(3 * np.log(d))
    mu = lam // 2
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w /= w.sum()
    mueff = 1.0 / np.sum(w ** 2)
    cc = (4 + mueff / d) / (d + 4 + 2 * mueff / d)
    cs = (mueff + 2) / (d + mueff + 5)
    c1 = 2 / ((d + 1.3) ** 2 + mueff)
    cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((d + 2) ** 2 + mueff))
    damps = 1 + 2 * max(0.0, np.sqrt((mueff - 1) / (d + 1)) - 1) + cs
    chiN = np.sqrt(d) * (1 - 1 / (4.0 * d) + 1 / (21.0 * d * d))
    mean, sigma = np.clip(x0, lower, upper), float(sigma0)
    pc, ps = np.zeros(d), np.zeros(d)
    B, D, C = np.eye(d), np.ones(d), np.eye(d)
    invsqrtC = np.eye(d)
    eigeneval, counteval = 0, 0
    best_x, best_f = mean.copy(), np.inf
    scale = float(np.mean(upper - lower)) if np.all(np.isfinite(upper - lower)) else 1.0
    while counteval < maxfevals:
        k = min(lam, maxfevals - counteval)
        arz = rng.standard_normal((k, d))
        ary = arz @ (B * D).T
        arx = mean + sigma * ary
        feas = np.clip(arx, lower, upper)
        fit = np.empty(k)
        for i in range(k):
            fit[i] = float(np.squeeze(f(feas[i])))
            if fit[i] < best_f:
                best_f, best_x = fit[i], feas[i].copy()
        counteval += k
        if k < lam:
            break
        spread = max(float(np.max(fit) - np.min(fit)), 1e-12)
        pen = fit + spread * np.sum(((arx - feas) / scale) ** 2, axis=1)
        idx = np.argsort(pen, kind='stable')[:mu]
        old = mean
        mean = w @ arx[idx]
        ymean = (mean - old) / sigma
        ps = (1 - cs) * ps + np.sqrt(cs * (2 - cs) * mueff) * (invsqrtC @ ymean)
        hsig = np.linalg.norm(ps) / np.sqrt(1 - (1 - cs) ** (2.0 * counteval / lam)) / chiN < 1.4 + 2.0 / (d + 1)
        pc = (1 - cc) * pc + hsig * np.sqrt(cc * (2 - cc) * mueff) * ymean
        artmp = ary[idx]
        C = (1 - c1 - cmu) * C + c1 * (np.outer(pc, pc) + (1 - hsig) * cc * (2 - cc) * C) + cmu * (artmp.T * w) @ artmp
        sigma *= np.exp((cs / damps) * (np.linalg.norm(ps) / chiN - 1))
        if counteval - eigeneval > lam / (c1 + cmu) / d / 10.0:
            eigeneval = counteval
            C = np.triu(C) + np.triu(C, 1).T
            Dsq, B = np.linalg.eigh(C)
            D = np.sqrt(np.maximum(Dsq, 1e-300))
            invsqrtC = (B / D) @ B.T
        if spread < tolfun and sigma * np.max(D) < tolx:
            break
    return best_x, best_f


class OptCma(object):
    """GPyOpt/optimization/optimizer.py:229-261: `cma.fmin(f, x0, 0.6, bounds, maxfevals)`; pycma when importable, else the
    restatement above."""

    def __init__(self, bounds, maxiter=1000):
        self.bounds = bounds
        self.maxiter = maxiter

    def optimize(self, x0, f=None, df=None, f_df=None, maxfevals=1000):
        if len(self.bounds) == 1:
            raise IndexError("CMA does not work in problems of dimension 1.")
        lB, uB = np.asarray(self.bounds)[:, 0], np.asarray(self.bounds)[:, 1]
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
        try:
            import cma
        except ImportError:
            x, _ = cma_es_minimize(lambda x: f(np.array([x])), x0, 0.6, lB, uB, maxfevals=maxfevals)
        else:
            x = cma.fmin(lambda x: float(np.squeeze(f(np.array([x])))), x0, 0.6,
                         options={"bounds": [lB, uB], "verbose": -1, 'maxfevals': maxfevals})[0]
        return np.atleast_2d(x), f(np.atleast_2d(x))


def choose_optimizer(optimizer_name, bounds):
    """GPyOpt/optimization/optimizer.py:417-441 ('lbfgs' and 'CMA': the two the BOPL experiments use)."""
    if optimizer_name == 'lbfgs':
        return OptLbfgs(bounds)
    if optimizer_name == 'CMA':
        return OptCma(bounds)
    raise NotImplementedError("optimizer %r: only 'lbfgs' and 'CMA' are provided (DIRECT / sgd / adam are host-side options "
                              "of the reference that no BOPL experiment selects)" % (optimizer_name,))


def _round_optimum(space, x):
    """Continuous box: clip into the bounds (GPyOpt Design_space.round_optimum for continuous variables)."""
    b = np.asarray(space.get_bounds(), dtype=np.float64)
    return np.clip(np.atleast_2d(x), b[:, 0], b[:, 1])


def apply_optimizer(optimizer, x0, f, df=None, f_df=None, duplicate_manager=None, context_manager=None, space=None,
                    verbose=False, maxfevals=1000):
    """optimizer.py:264-302 without context / duplicate managers."""
    x0 = np.atleast_2d(x0)
    optimized_x, _ = optimizer.optimize(x0, f, df, f_df, maxfevals=maxfevals)
    x = _round_optimum(space, optimized_x) if space is not None else np.atleast_2d(optimized_x)
    return x, f(x)


class LockstepBatcher(object):
    """Gathers one `f_df(x)` request from every still-running worker into a single `f_df_batch(X)` call."""

    def __init__(self, f_df_batch, n_workers, with_ids=False):
        self.f_df_batch = f_df_batch
        self.with_ids = with_ids            # True: f_df_batch(X, ids) also gets the worker ids of the rows
        self.active = n_workers
        self.cv = threading.Condition()
        self.pending = {}
        self.results = {}
        self.batches = 0
        self.error = None

    def _flush(self):
        ids = sorted(self.pending)
        X = np.stack([self.pending[i] for i in ids])
        try:
            fx, g = self.f_df_batch(X, ids) if self.with_ids else self.f_df_batch(X)
            fx = np.asarray(fx, dtype=np.float64).reshape(len(ids))
            g = np.asarray(g, dtype=np.float64).reshape(len(ids), -1)
            for k, i in enumerate(ids):
                self.results[i] = (float(fx[k]), g[k].copy())
        except Exception as e:                                   # propagate to every waiting worker
            self.error = e
            for i in ids:
                self.results[i] = None
        self.batches += 1
        self.pending.clear()
        self.cv.notify_all()

    def call(self, wid, x):
        with self.cv:
            self.pending[wid] = np.array(x, dtype=np.float64).reshape(-1)
            if len(self.pending) == self.active:
                self._flush()
            else:
                while wid not in self.results:
                    self.cv.wait()
            out = self.results.pop(wid)
            if out is None:
                raise self.error
            return out

    def done(self, wid):
        with self.cv:
            self.active -= 1
            if self.pending and len(self.pending) == self.active:
                self._flush()


def lockstep_lbfgs(anchors, f_df_batch, bounds, maxiter=300):
    """Run one unmodified scipy L-BFGS-B per anchor, in lock-step.  Returns (list of x_opt (1,d) or None-flag, batches)."""
    n = len(anchors)
    batcher = LockstepBatcher(f_df_batch, n)
    out = [None] * n
    errors = []

    def work(i):
        try:
            opt = OptLbfgs(bounds, maxiter)
            out[i] = opt.optimize(anchors[i], f_df=lambda x: batcher.call(i, x))
        except Exception as e:
            errors.append(e)
        finally:
            batcher.done(i)

    threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out, batcher.batches


class AcquisitionOptimizer(object):
    """Same constructor surface as the reference (acquisition_optimizer.py:26); `optimize(f, df, f_df)` -> (x_min, fx_min)."""

    def __init__(self, space, model=None, utility=None, expectation_utility=None, optimizer='lbfgs', inner_optimizer='lbfgs',
                 parallel=False, n_starting=400, n_anchor=20, include_baseline_points=False, batched=True, **kwargs):
        self.space = space if hasattr(space, "get_bounds") else BoxSpace(space)
        self.model = model
        self.utility = utility
        self.expectation_utility = expectation_utility
        self.optimizer_name = optimizer
        self.inner_optimizer_name = inner_optimizer
        self.parallel = parallel
        self.n_starting = n_starting
        self.n_anchor = n_anchor
        self.include_baseline_points = include_baseline_points
        self.batched = batched
        self.number_of_utility_parameter_samples = 3                                  # acquisition_optimizer.py:38
        if include_baseline_points and (model is None or utility is None):
            raise ValueError("include_baseline_points needs the model and the utility")
        if utility is not None and utility.parameter_distribution is not None:
            self.full_parameter_support = self.utility.parameter_distribution.use_full_support
        if model is not None:
            self.number_of_gp_hyps_samples = min(10, self.model.number_of_hyps_samples())
            self.n_attributes = self.model.output_dim
        self.kwargs = kwargs
        self.context_manager = ContextManager(self.space)
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        self.inner_optimizer = choose_optimizer(self.inner_optimizer_name, self.context_manager.noncontext_bounds)
        self.last_stats = {}

    # ---- local optimisation from a set of anchors (serial = the reference's loop; batched = lock-step)
    def _local_search(self, anchor_points, f, f_df, optimizer, final_f):
        """Returns [(x (1,d), fx)] per anchor.  final_f=True re-evaluates f at the rounded optimum as `apply_optimizer`
        does (optimizer.py:295-300); False returns L-BFGS's own value as `apply_optimizer_inner` does (:343)."""
        if f_df is not None and self.batched and len(anchor_points) > 1:
            results, batches = lockstep_lbfgs(list(anchor_points), f_df, optimizer.bounds, optimizer.maxiter)
            self.last_stats = {"anchors": len(anchor_points), "f_df_batches": batches}
            if final_f:
                # one point per call, as apply_optimizer does (optimizer.py:295-300): with H > 1 hyper-samples uEI's single-point
                # branch (uEI.py:49-50: ONE incumbent under the model's current hyper-sample) and its batch branch (per-hyper-sample
                # incumbents) give different values, and these values rank the optima
                X = np.vstack([_round_optimum(self.space, r[0]) for r in results])
                return [(X[i:i + 1], np.atleast_2d(np.asarray(f(X[i:i + 1]), dtype=np.float64).reshape(1))) for i in range(len(X))]
            return [(r[0], r[1] if r[1] is not None else np.atleast_2d(f(r[0]))) for r in results]
        self.last_stats = {"anchors": len(anchor_points), "f_df_batches": None}
        if final_f:
            return [apply_optimizer(optimizer, a, f=f, df=None, f_df=f_df, space=self.space) for a in anchor_points]
        out = []
        for a in anchor_points:
            x, fx = optimizer.optimize(np.atleast_2d(a), f, None, f_df)
            out.append((x, fx if fx is not None else np.atleast_2d(f(x))))
        return out

    def optimize(self, f=None, df=None, f_df=None, duplicate_manager=None):
        """acquisition_optimizer.py:60-137."""
        import random
        self.f, self.df, self.f_df = f, df, f_df
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        gen = ObjectiveAnchorPointsGenerator(self.space, "random", f, self.n_starting)
        anchor_points, anchor_points_values = gen.get(num_anchor=self.n_anchor, duplicate_manager=duplicate_manager,
                                                      context_manager=self.context_manager, get_scores=True)
        if self.include_baseline_points:                                              # :87-100
            if self.full_parameter_support:
                utility_parameter_samples = self.utility.parameter_distribution.support
            else:
                utility_parameter_samples = self.utility.parameter_distribution.sample(self.number_of_utility_parameter_samples)
            X_baseline = np.atleast_2d([self._current_marginal_argmax(th)[0, :] for th in utility_parameter_samples])
            fX_baseline = np.asarray(f(X_baseline))[:, 0]
            anchor_points = np.vstack((anchor_points, X_baseline))
            anchor_points_values = np.concatenate((np.asarray(anchor_points_values).reshape(-1), fX_baseline))
        optimized_points = self._local_search(anchor_points, f, f_df, self.optimizer, final_f=True)
        x_min, fx_min = min(optimized_points, key=lambda t: float(np.squeeze(t[1])))
        fx_min = np.squeeze(fx_min)
        if self.include_baseline_points:                                              # :118-134
            fx_min_baseline = fX_baseline.min()
            use = fx_min_baseline < fx_min or (fx_min_baseline == fx_min and np.random.binomial(1, 0.5, 1))
            if use:
                index = random.choice(np.atleast_1d(np.argmin(fX_baseline)))
                x_min = np.atleast_2d(X_baseline[index, :])
                fx_min = fX_baseline[index]
        return x_min, fx_min

    # ---- argmax_x E_n[U(f(x)) | theta]  (acquisition_optimizer.py:139-249)
    def marginal_objective(self, parameter, Z_samples=None):
        """The (f, f_df) pair the reference builds inside `_current_marginal_argmax`: minus the expected utility of the
        NOISELESS posterior under the fixed parameter, summed (not averaged) over hyper-samples and draws.
        affine utility: theta . posterior mean (:145-187); ExpectationUtility callables: evaluated on the host from the
        GPU's predict_noiseless and gradients (:189-209); otherwise 50 fixed N(0,1) draws (:211-247)."""
        model, eng = self.model, self.model.engine
        n_h = self.number_of_gp_hyps_samples
        m = self.n_attributes
        theta = np.ascontiguousarray(np.asarray(parameter, dtype=np.float64).reshape(1, -1))
        if self.utility.affine or self.expectation_utility is None:
            if self.utility.affine:
                kind, Z = "linear", np.zeros((1, m))
                if theta.shape[1] != m:
                    raise ValueError("affine utility needs one parameter per attribute")
            else:
                kind = self.utility.resolve_kind(m, theta[0])
                Z = np.random.normal(size=(50, m)) if Z_samples is None else np.asarray(Z_samples, dtype=np.float64)
            th_d, Z_d, w_d = eng.to_dev(theta), eng.to_dev(Z), eng.to_dev(np.ones(1))

            def val_func(X):
                X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
                val, _ = eng.expected_utility_dev(eng.to_dev(X), Z_d, kind, th_d, w_d, n_h, False)
                model.set_hyperparameters(n_h - 1)
                return -val.cpu().numpy().reshape(-1, 1)

            def val_func_with_gradient(X):
                X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
                val, dval = eng.expected_utility_dev(eng.to_dev(X), Z_d, kind, th_d, w_d, n_h, True)
                model.set_hyperparameters(n_h - 1)
                return -val.cpu().numpy().reshape(-1, 1), -dval.cpu().numpy().reshape(X.shape)
            return val_func, val_func_with_gradient

        eu = self.expectation_utility

        def val_func(X):
            X = np.atleast_2d(X)
            func_val = np.zeros((X.shape[0], 1))
            for h in range(n_h):
                model.set_hyperparameters(h)
                mean, var = model.predict_noiseless(X)
                for i in range(X.shape[0]):
                    func_val[i, 0] += eu.func(parameter, mean[:, i], var[:, i])
            return -func_val

        def val_func_with_gradient(X):
            X = np.atleast_2d(X)
            func_val = np.zeros((X.shape[0], 1))
            func_gradient = np.zeros(X.shape)
            for h in range(n_h):
                model.set_hyperparameters(h)
                mean, var = model.predict_noiseless(X)
                aux = np.concatenate((model.posterior_mean_gradient(X), model.posterior_variance_gradient(X)))
                for i in range(X.shape[0]):
                    func_val[i, 0] += eu.func(parameter, mean[:, i], var[:, i])
                    func_gradient[i, :] += np.matmul(eu.gradient(parameter, mean[:, i], var[:, i]), aux[:, i])
            return -func_val, -func_gradient
        return val_func, val_func_with_gradient

    def _current_marginal_argmax(self, parameter):
        f, f_df = self.marginal_objective(parameter)
        return self.optimize_inner_func(f=f, f_df=f_df)[0]

    def optimize_inner_func(self, f=None, df=None, f_df=None, duplicate_manager=None, n_starting=400, n_anchor=20):
        """acquisition_optimizer.py:251-281: latin design -> anchors -> L-BFGS per anchor -> best (value as returned by
        L-BFGS, `apply_optimizer_inner`)."""
        self.inner_optimizer = choose_optimizer(self.inner_optimizer_name, self.context_manager.noncontext_bounds)
        gen = ObjectiveAnchorPointsGenerator(self.space, "latin", f, n_starting)
        anchor_points, _ = gen.get(num_anchor=n_anchor, duplicate_manager=duplicate_manager,
                                   context_manager=self.context_manager, get_scores=True)
        optimized_points = self._local_search(anchor_points, f, f_df, self.inner_optimizer, final_f=False)
        x_min, fx_min = min(optimized_points, key=lambda t: float(np.squeeze(t[1])))
        return x_min, fx_min
