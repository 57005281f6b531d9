"""Drop-in for `bopl/models/multi_outputGP.py:9-290` (class MultiOutputGP) with the posterior on the GPU.

Same constructor, attributes and method names as the reference.  Every `predict / posterior_*` call runs the
sm_100a kernels through the C-ABI; the per-attribute Python loop of the reference (multi_outputGP.py:91-141,
225-246) is replaced by one launch over all attributes.  What stays on the host is what the reference keeps in
LAPACK once per model update — the Cholesky factorisation (`exact_gaussian_inference.py:44-51`), `dpotrs`, `dpotri` —
runs on the device too by default (`Engine.fit_model` -> `bopl_fit_model`); `state_from_hyperparameters` below is the
host/LAPACK form of the same fit (used by `fit_on_device=False`, by `from_reference_model`, and as the tests' comparison).

Hyper-parameter inference (MAP + HMC, `GPyOpt/models/gpmodel.py:109-149`): `updateModel` runs `bopl_b200/fit.py` (same
model, priors and sampler settings as the reference; the optimiser / sampler logic on the host, every likelihood + gradient
evaluation of all attributes as one batched device call) unless the hyper-samples were supplied with
`set_hyperparameter_samples(...)`; a live reference model can be wrapped with `from_reference_model`.
"""
import numpy as np
from scipy.linalg import lapack

from .engine import Engine, GPState

_SQRT3 = np.sqrt(3.0)
_SQRT5 = np.sqrt(5.0)


# ------------------------------------------------------------------------------------------- host fit (LAPACK)
def _train_cov(kind, variance, lengthscale, X):
    """K(X, X) as the reference builds it (GPy/kern/src/stationary.py:104-113,128-166): expanded squared
    distance of the scaled inputs, zero diagonal, clip at 0, sqrt."""
    Xs = X / lengthscale
    sq = np.sum(np.square(Xs), 1)
    r2 = -2.0 * Xs.dot(Xs.T) + (sq[:, None] + sq[None, :])
    r2[np.diag_indices_from(r2)] = 0.0
    r = np.sqrt(np.clip(r2, 0.0, np.inf))
    if kind == "Mat52":
        return variance * (1.0 + _SQRT5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-_SQRT5 * r)
    if kind == "Mat32":
        return variance * (1.0 + _SQRT3 * r) * np.exp(-_SQRT3 * r)
    if kind in ("rbf", "ExpQuad"):
        return variance * np.exp(-0.5 * r ** 2)
    raise ValueError("unsupported kernel %r (Mat52, Mat32, rbf, ExpQuad)" % (kind,))


def _jitchol(A, maxtries=10):
    """GPy/util/linalg.py:53-78."""
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diag = np.diag(A)
    if np.any(diag <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diag.mean() * 1e-6
    for _ in range(maxtries):
        try:
            return np.linalg.cholesky(A + np.eye(A.shape[0]) * jitter)
        except np.linalg.LinAlgError:
            jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def fit_exact_gp(X, Y, kind, variance, lengthscale, noise, want_winv=True):
    """One GPRegression posterior: (L, alpha, ymean, Winv).  Follows GPy/core/gp.py:52-62 (mean-centring normaliser of
    this fork: util/normalizer.py:60-63) and exact_gaussian_inference.py:44-51; Winv: posterior.py:171-191."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(len(X), 1)
    ymean = float(Y.mean())
    Ky = _train_cov(kind, float(variance), np.asarray(lengthscale, dtype=np.float64), X)
    Ky[np.diag_indices_from(Ky)] += float(noise) + 1e-8
    L = _jitchol(Ky)
    alpha = lapack.dpotrs(np.asfortranarray(L), Y - ymean, lower=1)[0]
    Winv = None
    if want_winv:
        Winv, _ = lapack.dpotri(np.asfortranarray(L), lower=1)
        Winv = np.tril(Winv) + np.tril(Winv, -1).T
    return np.tril(L), alpha[:, 0], ymean, Winv


def inverse_factor(L):
    """L^-1 by LAPACK dtrtri (lower, non-unit) — the factor of the small-batch matrix-vector path."""
    Linv, info = lapack.dtrtri(np.asfortranarray(L), lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("dtrtri failed (info=%d)" % info)
    return np.tril(Linv)


def state_from_hyperparameters(X, Y, kinds, variance, lengthscale, noise, want_winv=True, want_linv=True):
    """Build a GPState for m attributes x H hyper-samples.  Y: list / array of m vectors (n,) or (n,1)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    variance = np.asarray(variance, dtype=np.float64)
    m, H = variance.shape
    n, d = X.shape
    lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(m, H, -1)
    if lengthscale.shape[2] == 1 and d > 1:
        lengthscale = np.repeat(lengthscale, d, axis=2)        # non-ARD kernel: one lengthscale for every input
    noise = np.asarray(noise, dtype=np.float64).reshape(m, H)
    kinds = np.asarray(kinds)
    kinds = np.repeat(kinds.reshape(m, 1), H, axis=1) if kinds.size == m else kinds.reshape(m, H)
    L = np.empty((m, H, n, n))
    alpha = np.empty((m, H, n))
    ymean = np.empty((m, H))
    Winv = np.empty((m, H, n, n)) if want_winv else None
    Linv = np.empty((m, H, n, n)) if want_linv else None
    for j in range(m):
        for h in range(H):
            Lj, aj, yj, Wj = fit_exact_gp(X, Y[j], str(kinds[j, h]), variance[j, h], lengthscale[j, h], noise[j, h], want_winv)
            L[j, h], alpha[j, h], ymean[j, h] = Lj, aj, yj
            if want_winv:
                Winv[j, h] = Wj
            if want_linv:
                Linv[j, h] = inverse_factor(Lj)
    Ym = np.stack([np.asarray(Y[j], dtype=np.float64).reshape(n) for j in range(m)])
    return GPState(kinds, X, lengthscale, variance, noise, ymean, L, alpha, Winv, Linv, Y=Ym)


def state_from_reference_model(ref_model, want_winv=True):
    """Extract the cached posterior state from a live reference `MultiOutputGP` (duck-typed):
    per `ref_model.output[j].model_instances[h]`: .X, .kern.name/.variance/.lengthscale, .likelihood.variance,
    .normalizer.mean, .posterior._woodbury_chol, .posterior.woodbury_vector, .posterior.woodbury_inv."""
    m = ref_model.output_dim
    insts = [list(ref_model.output[j].model_instances) for j in range(m)]
    H = min(len(i) for i in insts)
    X = np.asarray(insts[0][0].X, dtype=np.float64)
    n, d = X.shape
    kinds = np.empty((m, H), dtype=object)
    variance, noise, ymean = np.empty((m, H)), np.empty((m, H)), np.empty((m, H))
    lengthscale = np.empty((m, H, d))
    L, alpha = np.empty((m, H, n, n)), np.empty((m, H, n))
    Winv = np.empty((m, H, n, n)) if want_winv else None
    for j in range(m):
        for h in range(H):
            g = insts[j][h]
            kinds[j, h] = str(g.kern.name)
            variance[j, h] = float(np.asarray(g.kern.variance).reshape(-1)[0])
            ls = np.asarray(g.kern.lengthscale, dtype=np.float64).reshape(-1)
            lengthscale[j, h] = ls if ls.size == d else np.repeat(ls, d)
            noise[j, h] = float(np.asarray(g.likelihood.variance).reshape(-1)[0])
            nz = getattr(g, "normalizer", None)
            ymean[j, h] = float(np.asarray(nz.mean).reshape(-1)[0]) if nz is not None else 0.0
            L[j, h] = np.tril(np.asarray(g.posterior._woodbury_chol, dtype=np.float64))
            alpha[j, h] = np.asarray(g.posterior.woodbury_vector, dtype=np.float64).reshape(-1)
            if want_winv:
                Winv[j, h] = np.asarray(g.posterior.woodbury_inv, dtype=np.float64).reshape(n, n)
    Linv = np.stack([np.stack([inverse_factor(L[j, h]) for h in range(H)]) for j in range(m)])
    Y = None
    if all(getattr(insts[j][0], "Y", None) is not None for j in range(m)):
        Y = np.stack([np.asarray(insts[j][0].Y, dtype=np.float64).reshape(n) for j in range(m)])
    return GPState(kinds, X, lengthscale, variance, noise, ymean, L, alpha, Winv, Linv, Y=Y)


class BasicModel(object):
    """models/basic_model.py:6-44: only stores the evaluations (the model ParEGO / Random are paired with)."""
    analytical_gradient_prediction = True

    def __init__(self, output_dim=None, X=None, Y=None):
        self.output_dim, self.X, self.Y = output_dim, X, Y
        self.name = 'basic model'

    def updateModel(self, X, Y):
        self.X, self.Y = X, Y

    def get_X(self):
        return np.copy(self.X)

    def get_Y(self):
        import copy
        return copy.deepcopy(self.Y)

    def get_XY(self):
        return self.get_X(), self.get_Y()

    def get_model_parameters(self):
        pass

    def get_model_parameters_names(self):
        pass


# ------------------------------------------------------------------------------------------- the drop-in class
class MultiOutputGP(object):
    """GPU-backed stand-in for `bopl.models.MultiOutputGP` (multi_outputGP.py:9-290)."""

    analytical_gradient_prediction = True

    def __init__(self, output_dim, kernel=None, noise_var=None, exact_feval=None, n_samples=10, ARD=None,
                 fixed_hyps=False, device=0, gradients=True, fit_options=None, fit_on_device=True, posterior_kernel=None):
        self.output_dim = output_dim
        self.kernel = [None] * output_dim if kernel is None else kernel          # GPy kernel names ('Mat52', 'rbf', ...)
        self.noise_var = [None] * output_dim if noise_var is None else noise_var
        self.exact_feval = [False] * output_dim if exact_feval is None else exact_feval
        self.n_samples = n_samples
        self.ARD = [True] * output_dim if ARD is None else ARD
        self.fixed_hyps = fixed_hyps
        self.name = 'multi-output GP'                                              # checked in core/bopl.py:46,440
        self.gradients = gradients
        self.fit_options = dict(fit_options or {})        # n_burnin, subsample_interval, step_size, leapfrog_steps, max_iters
        self.fit_on_device = fit_on_device                # False: LAPACK on the host + upload (bopl_set_model), as round 1 did
        self.engine = Engine(device, posterior_kernel)    # None: the library default (int8x7, or $BOPL_POSTERIOR); 'fp64' / 'int8'
        self._hyps = None
        self._h = 0
        self.state = None
        self.X_all = None
        self.Y_all = None

    # ---- construction helpers
    @classmethod
    def from_state(cls, state, device=0, posterior_kernel=None):
        self = cls(state.m, n_samples=state.H, device=device, gradients=state.Winv is not None, posterior_kernel=posterior_kernel)
        self._load(state)
        return self

    @classmethod
    def from_reference_model(cls, ref_model, device=0, gradients=True, posterior_kernel=None):
        return cls.from_state(state_from_reference_model(ref_model, gradients), device, posterior_kernel)

    def set_hyperparameter_samples(self, variance, lengthscale, noise=None, kinds=None):
        """variance (m,H), lengthscale (m,H,d) or (m,H,1), noise (m,H) or None -> noise_var / 1e-6 (exact_feval)."""
        variance = np.asarray(variance, dtype=np.float64)
        m, H = variance.shape
        if m != self.output_dim:
            raise ValueError("variance must be (output_dim, H)")
        if noise is None:
            noise = np.empty((m, H))
            for j in range(m):
                if self.exact_feval[j]:
                    noise[j] = 1e-6                                                # gpmodel.py:75-76
                elif self.noise_var[j] is not None:
                    noise[j] = self.noise_var[j]
                else:
                    raise ValueError("noise is required for attribute %d (neither exact_feval nor noise_var given)" % j)
        if kinds is None:
            kinds = [k if k is not None else "Mat52" for k in self.kernel]         # gpmodel.py:61
        self._hyps = (np.asarray(kinds), variance, np.asarray(lengthscale, dtype=np.float64), np.asarray(noise, dtype=np.float64))
        self.n_samples = H

    def _load(self, state):
        self.state = state
        self.engine.set_model(state)
        self.n_samples = state.H
        self._h = 0                                                                 # gpmodel.py:107

    def updateModel(self, X_all, Y_all):
        """multi_outputGP.py:57-62.  Refits every (attribute, hyper-sample) posterior: on the device (`bopl_fit_model`:
        covariance, blocked Cholesky, alpha, K^-1, L^-1 in one batched pass; only X, Y and the hyper-parameters are uploaded)
        or, with `fit_on_device=False`, on the host with LAPACK followed by an upload of the factors."""
        self.X_all, self.Y_all = X_all, Y_all
        if self._hyps is not None:
            kinds, variance, lengthscale, noise = self._hyps
        else:
            kinds, variance, lengthscale, noise = self._infer_hyperparameters(X_all, Y_all)
        if self.fit_on_device:
            self.state = self.engine.fit_model(X_all, Y_all, kinds, variance, lengthscale, noise,
                                               want_winv=self.gradients, want_linv=True)
            self.n_samples = self.state.H
            self._h = 0                                                             # gpmodel.py:107
        else:
            self._load(state_from_hyperparameters(X_all, Y_all, kinds, variance, lengthscale, noise, self.gradients))

    def _infer_hyperparameters(self, X_all, Y_all):
        """Per attribute: MAP + HMC hyper-samples on the host (GPyOpt/models/gpmodel.py:109-149); `fixed_hyps=True` keeps
        the MAP estimate only."""
        from .fit import fit_hyperparameter_samples, fit_hyperparameter_samples_device
        X = np.asarray(X_all, dtype=np.float64)
        m, d = self.output_dim, X.shape[1]
        H = 1 if self.fixed_hyps else self.n_samples
        kinds = np.array([k if k is not None else "Mat52" for k in self.kernel])
        if self.fit_on_device:
            variance, lengthscale, noise = fit_hyperparameter_samples_device(
                self.engine, X, [Y_all[j] for j in range(m)], kinds, self.ARD, self.exact_feval, self.noise_var, n_samples=H,
                hmc=not self.fixed_hyps, **self.fit_options)
            self.n_samples = H
            return kinds, variance, lengthscale, noise
        variance, lengthscale, noise = np.empty((m, H)), np.empty((m, H, d)), np.empty((m, H))
        for j in range(m):
            v, l, s = fit_hyperparameter_samples(X, Y_all[j], str(kinds[j]), bool(self.ARD[j]), bool(self.exact_feval[j]),
                                                 self.noise_var[j], n_samples=H, hmc=not self.fixed_hyps, **self.fit_options)
            variance[j], lengthscale[j], noise[j] = v, l, s
        self.n_samples = H
        return kinds, variance, lengthscale, noise

    # ---- reference surface
    def number_of_hyps_samples(self):
        return self.n_samples

    def set_hyperparameters(self, n):                                              # multi_outputGP.py:70-72
        if not 0 <= n < self.n_samples:
            raise IndexError("hyper-sample index out of range")
        self._h = int(n)

    def get_evaluated_points(self):
        return np.copy(self.state.X)

    def _x(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        if X.shape[1] != self.state.d:
            raise ValueError("X must have %d columns" % self.state.d)
        return self.engine.to_dev(X)

    def predict(self, X, full_cov=False):                                          # multi_outputGP.py:91-102
        if full_cov:
            raise NotImplementedError("full_cov=True is not on the accelerated path")
        mean, var = self.engine.posterior_dev(self._x(X), self._h, include_noise=True, clip=True)
        return mean.cpu().numpy(), var.cpu().numpy()

    def predict_noiseless(self, X, full_cov=False):                                # multi_outputGP.py:104-115
        if full_cov:
            raise NotImplementedError("full_cov=True is not on the accelerated path")
        mean, var = self.engine.posterior_dev(self._x(X), self._h, include_noise=False, clip=True)
        return mean.cpu().numpy(), var.cpu().numpy()

    def posterior_mean(self, X):                                                   # multi_outputGP.py:117-125
        return self.engine.posterior_mean_dev(self._x(X), self._h).cpu().numpy()

    def posterior_mean_at_evaluated_points(self):                                  # multi_outputGP.py:127-131
        return self.engine.train_mean_dev()[self._h].cpu().numpy()

    def posterior_variance(self, X):                                               # multi_outputGP.py:133-141
        _, var = self.engine.posterior_dev(self._x(X), self._h, include_noise=True, clip=True, want_mean=False)
        return var.cpu().numpy()

    def posterior_variance_noiseless(self, X):
        _, var = self.engine.posterior_dev(self._x(X), self._h, include_noise=False, clip=True, want_mean=False)
        return var.cpu().numpy()

    def posterior_mean_gradient(self, X):                                          # multi_outputGP.py:225-235
        dmean, _ = self.engine.posterior_grad_dev(self._x(X), self._h, want_var=False)
        return dmean.cpu().numpy()

    def posterior_variance_gradient(self, X):                                      # multi_outputGP.py:237-246
        _, dvar = self.engine.posterior_grad_dev(self._x(X), self._h, want_mean=False)
        return dvar.cpu().numpy()

    def get_copy_of_model_sample(self):                                            # multi_outputGP.py:286-290
        """One view per attribute onto a fresh device-side posterior sample path under hyper-sample 0
        (gpmodel.py:167-168 copies `model_instances[0]`); see `bopl_b200.sampling_policies.SamplePath`."""
        from .sampling_policies import SamplePath
        return SamplePath(self).attributes

    def get_model_parameters(self):                                                # multi_outputGP.py:268-276
        s = self.state
        return [np.concatenate([[s.variance[j, self._h]], s.lengthscale[j, self._h], [s.noise[j, self._h]]])[None, :]
                for j in range(self.output_dim)]

    def get_model_parameters_names(self):                                          # multi_outputGP.py:278-284
        d = self.state.d
        names = ["Mat52.variance"] + ["Mat52.lengthscale[%d]" % q for q in range(d)] + ["Gaussian_noise.variance"]
        return [list(names) for _ in range(self.output_dim)]
