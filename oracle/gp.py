"""Oracle: exact-GP posterior maths of the reference's vendored GPy/GPyOpt fork (TEST INFRASTRUCTURE).

References are relative to /root/reference/bopl/aux_software/ unless they start with `bopl/`.
All arithmetic fp64; LAPACK through scipy.linalg.lapack exactly where GPy/util/linalg.py uses it.
"""
import numpy as np
from scipy.linalg import lapack

KERNEL_KINDS = ("Mat52", "rbf", "Mat32", "ExpQuad")  # GPy kernel `.name`s: stationary.py:516,427,581; rbf.py:22

_SQRT3 = np.sqrt(3.0)
_SQRT5 = np.sqrt(5.0)


# ----------------------------------------------------------------------------- kernels
def unscaled_dist(X, X2):
    """Euclidean distance in the expanded form ||a||^2+||b||^2-2ab with a clip at 0.
    Follows GPy/kern/src/stationary.py:128-146 (the X2-given branch)."""
    X1sq = np.sum(np.square(X), 1)
    X2sq = np.sum(np.square(X2), 1)
    r2 = -2.0 * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def unscaled_dist_self(X):
    """X2=None branch of stationary.py:134-139 (tdot + forced zero diagonal)."""
    Xsq = np.sum(np.square(X), 1)
    r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
    np.fill_diagonal(r2, 0.0)
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def scaled_dist(X, X2, lengthscale, ARD=True):
    """GPy/kern/src/stationary.py:149-166: ARD divides the inputs, non-ARD divides the distance."""
    lengthscale = np.asarray(lengthscale, dtype=np.float64)
    if ARD:
        if X2 is None:
            return unscaled_dist_self(X / lengthscale)
        return unscaled_dist(X / lengthscale, X2 / lengthscale)
    if X2 is None:
        return unscaled_dist_self(X) / lengthscale
    return unscaled_dist(X, X2) / lengthscale


def K_of_r(kind, variance, r):
    """Mat52 stationary.py:529-530; Mat32 :440-441; ExpQuad :594-595; rbf rbf.py:42-43."""
    if kind == "Mat52":
        return variance * (1 + _SQRT5 * r + 5.0 / 3 * r ** 2) * np.exp(-_SQRT5 * r)
    if kind == "Mat32":
        return variance * (1.0 + _SQRT3 * r) * np.exp(-_SQRT3 * r)
    if kind in ("rbf", "ExpQuad"):
        return variance * np.exp(-0.5 * r ** 2)
    raise ValueError("unknown kernel kind %r" % (kind,))


def dK_dr(kind, variance, r):
    """Mat52 stationary.py:532-533; Mat32 :443-444; ExpQuad :597-598; rbf rbf.py:45-46."""
    if kind == "Mat52":
        return variance * (10.0 / 3 * r - 5.0 * r - 5.0 * _SQRT5 / 3 * r ** 2) * np.exp(-_SQRT5 * r)
    if kind == "Mat32":
        return -3.0 * variance * r * np.exp(-_SQRT3 * r)
    if kind in ("rbf", "ExpQuad"):
        return -r * K_of_r(kind, variance, r)
    raise ValueError("unknown kernel kind %r" % (kind,))


def kern_K(kind, variance, lengthscale, X, X2=None, ARD=True):
    """Stationary.K, stationary.py:104-113."""
    return K_of_r(kind, variance, scaled_dist(X, X2, lengthscale, ARD))


def kern_gradients_X(kind, variance, lengthscale, dL_dK, X, X2, ARD=True):
    """Stationary._gradients_X_pure with X2 given (stationary.py:312-322) + _inv_dist (:227-234) +
    dK_dr_via_X (:115-121).  Chunked over rows of X so the N x n x d temporary stays small; the
    per-row arithmetic is the reference's (sum over the train axis of tmp * (x - x2), / l^2)."""
    lengthscale = np.asarray(lengthscale, dtype=np.float64)
    N = X.shape[0]
    grad = np.empty((N, X.shape[1]))
    dL_dK = np.broadcast_to(dL_dK, (N, X2.shape[0]))
    step = max(1, int(2 ** 22 // max(1, X2.shape[0] * X.shape[1])))
    for s in range(0, N, step):
        Xs = X[s:s + step]
        dist = scaled_dist(Xs, X2, lengthscale, ARD)
        invdist = 1.0 / np.where(dist != 0.0, dist, np.inf)
        tmp = invdist * (dK_dr(kind, variance, dist) * dL_dK[s:s + step])
        d = Xs[:, None, :] - X2[None, :, :]
        grad[s:s + step] = np.sum(tmp[:, :, None] * d, 1) / lengthscale ** 2
    return grad


# ----------------------------------------------------------------------------- LAPACK wrappers
def jitchol(A, maxtries=10):
    """GPy/util/linalg.py:53-78: dpotrf(lower); on failure add jitter = mean(diag)*1e-6 * 10^k."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        try:
            return np.linalg.cholesky(A + np.eye(A.shape[0]) * jitter)
        except np.linalg.LinAlgError:
            jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(A, B, lower=1):
    """GPy/util/linalg.py:92-111."""
    return lapack.dtrtrs(np.asfortranarray(A), B, lower=lower, trans=0, unitdiag=0)[0]


def dpotrs(A, B, lower=1):
    """GPy/util/linalg.py:113-122."""
    return lapack.dpotrs(np.asfortranarray(A), B, lower=lower)[0]


def dpotri(A, lower=1):
    """GPy/util/linalg.py:124-142 (dpotri then symmetrify: copy the computed triangle across)."""
    R, _ = lapack.dpotri(np.asfortranarray(A), lower=lower)
    if lower:
        R = np.tril(R) + np.tril(R, -1).T
    else:
        R = np.triu(R) + np.triu(R, 1).T
    return R


# ----------------------------------------------------------------------------- one exact GP
class ExactGP(object):
    """State of one `GPy.models.GPRegression` instance after `parameters_changed`.

    Fit: GPy/core/gp.py:52-62 (normalizer; this fork's Standardize has std=1: util/normalizer.py:60-63)
    and inference/latent_function_inference/exact_gaussian_inference.py:29-65."""

    def __init__(self, X, Y, kind="Mat52", variance=1.0, lengthscale=None, noise=1e-6, ARD=True, normalizer=True):
        self.X = np.array(X, dtype=np.float64)
        Y = np.array(Y, dtype=np.float64).reshape(len(self.X), 1)
        self.kind = kind
        self.variance = float(variance)
        d = self.X.shape[1]
        if lengthscale is None:
            lengthscale = np.ones(d if ARD else 1)
        self.lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))
        if ARD and self.lengthscale.size == 1:
            self.lengthscale = np.ones(d) * self.lengthscale
        self.ARD = ARD
        self.noise = float(noise)
        self.Y = Y
        self.ymean = float(Y.mean(0)[0]) if normalizer else 0.0
        Yn = Y - self.ymean
        K = kern_K(kind, self.variance, self.lengthscale, self.X, None, ARD)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += self.noise + 1e-8          # exact_gaussian_inference.py:46-47
        self.L = jitchol(Ky)                                        # pdinv -> jitchol, linalg.py:207
        self.alpha = dpotrs(self.L, Yn, lower=1)                    # exact_gaussian_inference.py:51
        self._Winv = None
        self.Ky = Ky

    @property
    def Winv(self):
        """Posterior.woodbury_inv, posterior.py:171-191 (dpotri of the Cholesky factor + symmetrify)."""
        if self._Winv is None:
            self._Winv = dpotri(self.L, lower=1)
        return self._Winv

    def K(self, X, X2=None):
        return kern_K(self.kind, self.variance, self.lengthscale, X, X2, self.ARD)

    # --- GPy/core/gp.py:380-400 + posterior.py:299-305 + normalizer.py:67-68
    def posterior_mean(self, Xnew):
        Kx = self.K(Xnew, self.X)
        return np.dot(Kx, self.alpha) + self.ymean

    def raw_posterior_variance(self, Xnew):
        """posterior.py:308-320: var = Kdiag - colsum(dtrtrs(L, K(X, Xnew))^2)."""
        Kx = self.K(self.X, Xnew)
        Kxx = np.full(Xnew.shape[0], self.variance)                 # stationary.py:168-171
        tmp = dtrtrs(self.L, Kx)
        return (Kxx - np.square(tmp).sum(0))[:, None]

    def posterior_variance(self, Xnew):
        """gp.py:403-418: raw variance + likelihood variance (gaussian.py:110-111)."""
        return self.noise + self.raw_posterior_variance(Xnew)

    def posterior_variance_noiseless(self, Xnew):
        """gp.py:421-435."""
        return self.raw_posterior_variance(Xnew)

    def predict(self, Xnew, include_likelihood=True):
        """gp.py:286-326 + posterior.py:268-296 (diagonal branch): mean un-normalised, variance left as is."""
        Kx = self.K(self.X, Xnew)
        mu = np.dot(Kx.T, self.alpha)
        tmp = dtrtrs(self.L, Kx)
        var = (np.full(Xnew.shape[0], self.variance) - np.square(tmp).sum(0))[:, None]
        if include_likelihood:
            var = var + self.noise                                   # gaussian.py:94-102
        return mu + self.ymean, var

    def posterior_mean_gradient(self, X):
        """gp.py:438-461: kern.gradients_X(woodbury_vector.T, X, Xtrain)."""
        return kern_gradients_X(self.kind, self.variance, self.lengthscale, self.alpha.T, X, self.X, self.ARD)

    def posterior_variance_gradient(self, X):
        """gp.py:464-490.  The `gradients_X(eye(N), X)` term is identically zero for a stationary kernel
        (inverse distance is forced to 0 on the diagonal, stationary.py:233-234), so only the Schur part stays."""
        alpha = -2.0 * np.dot(self.K(X, self.X), self.Winv)
        return kern_gradients_X(self.kind, self.variance, self.lengthscale, alpha, X, self.X, self.ARD)


# ----------------------------------------------------------------------------- GPyOpt GPModel + BOPL MultiOutputGP
class MultiOutputGPOracle(object):
    """m attributes x H hyper-samples of ExactGP with the forwarding/clipping of
    GPyOpt/models/gpmodel.py:164-223 and the stacking of bopl/models/multi_outputGP.py:91-141,225-246."""

    def __init__(self, gps):
        """gps[j][h] : ExactGP for attribute j, hyper-sample h."""
        self.gps = gps
        self.output_dim = len(gps)
        self.n_samples = len(gps[0])
        self.h = 0                                                   # gpmodel.py:107 (set_hyperparameters(0) after create)

    def number_of_hyps_samples(self):
        return self.n_samples

    def set_hyperparameters(self, h):                                # multi_outputGP.py:70-72, gpmodel.py:164-165
        self.h = h

    def _cur(self, j):
        return self.gps[j][self.h]

    def get_evaluated_points(self):
        return np.copy(self.gps[0][0].X)

    def posterior_mean(self, X):                                     # multi_outputGP.py:117-125
        X = np.atleast_2d(X)
        return np.stack([self._cur(j).posterior_mean(X)[:, 0] for j in range(self.output_dim)])

    def posterior_mean_at_evaluated_points(self):                    # multi_outputGP.py:127-131
        return self.posterior_mean(self.gps[0][0].X)

    def posterior_variance(self, X):                                 # multi_outputGP.py:133-141, clip gpmodel.py:214
        X = np.atleast_2d(X)
        return np.stack([np.clip(self._cur(j).posterior_variance(X), 1e-10, np.inf)[:, 0]
                         for j in range(self.output_dim)])

    def posterior_variance_noiseless(self, X):                       # gpmodel.py:217-223
        X = np.atleast_2d(X)
        return np.stack([np.clip(self._cur(j).posterior_variance_noiseless(X), 1e-10, np.inf)[:, 0]
                         for j in range(self.output_dim)])

    def predict(self, X):                                            # multi_outputGP.py:91-102, gpmodel.py:170-179
        X = np.atleast_2d(X)
        ms, vs = [], []
        for j in range(self.output_dim):
            m, v = self._cur(j).predict(X, True)
            ms.append(m[:, 0])
            vs.append(np.clip(v, 1e-10, np.inf)[:, 0])
        return np.stack(ms), np.stack(vs)

    def predict_noiseless(self, X):                                  # multi_outputGP.py:104-115, gpmodel.py:181-190
        X = np.atleast_2d(X)
        ms, vs = [], []
        for j in range(self.output_dim):
            m, v = self._cur(j).predict(X, False)
            ms.append(m[:, 0])
            vs.append(np.clip(v, 1e-10, np.inf)[:, 0])
        return np.stack(ms), np.stack(vs)

    def posterior_mean_gradient(self, X):                            # multi_outputGP.py:225-235
        X = np.atleast_2d(X)
        return np.stack([self._cur(j).posterior_mean_gradient(X) for j in range(self.output_dim)])

    def posterior_variance_gradient(self, X):                        # multi_outputGP.py:237-246
        X = np.atleast_2d(X)
        return np.stack([self._cur(j).posterior_variance_gradient(X) for j in range(self.output_dim)])
