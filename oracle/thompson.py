"""Oracle: the sample-path objective of utility Thompson sampling (TEST INFRASTRUCTURE ONLY).

Restates `bopl/sampling_policies/uTS.py:31-50` on the oracle's `ExactGP`: a copy of hyper-sample 0 of every attribute
(`bopl/models/multi_outputGP.py:286-290` -> `GPyOpt/models/gpmodel.py:167-168`), and per evaluation, attribute by attribute:
`posterior_samples_f(x, size=1, full_cov=True)` (`GPy/core/gp.py:833-857`: `_raw_predict(full_cov=True)` — raw variance
`Kxx - tmp^T tmp`, no likelihood noise, `posterior.py:268-280` — then the normaliser's mean added back and one
`np.random.multivariate_normal` draw), the draw appended to the data and the model RE-FITTED from scratch (`set_XY`,
`gp.py:191-227`: normaliser mean over all targets, then `parameters_changed` -> `exact_gaussian_inference.py:29-65`).
Pinned against the reference's own closure by `tests/golden/uTS_*.npz` (`tests/golden/make_golden.py::main_thompson`).
"""
import numpy as np

from .gp import ExactGP, dtrtrs


class SamplePathOracle(object):
    def __init__(self, gps):
        """gps[j]: the `ExactGP` of attribute j under hyper-sample 0 (copied by re-fitting, never mutated)."""
        self.gps = list(gps)
        g0 = gps[0]
        self.X_aux = np.copy(g0.X)                                                   # uTS.py:33,37
        self.Y_aux = [np.copy(g.Y) for g in gps]                                     # uTS.py:34-38

    @staticmethod
    def _refit(g, X, Y):
        return ExactGP(X, Y, kind=g.kind, variance=g.variance, lengthscale=g.lengthscale, noise=g.noise, ARD=g.ARD)

    def evaluate(self, x, z=None):
        """uTS.py:40-47.  Returns (y, mean, var), each (m,).  z None: draw with np.random.multivariate_normal exactly as the
        reference does; z (m,): use mean + sqrt(var) z (what that call computes for a 1x1 covariance)."""
        x_new = np.atleast_2d(np.asarray(x, dtype=np.float64))
        self.X_aux = np.vstack((self.X_aux, x_new))
        ys, mus, vs = [], [], []
        for j, g in enumerate(self.gps):
            Kx = g.K(g.X, x_new)                                                     # posterior.py:270
            mu = np.dot(Kx.T, g.alpha)                                               # :271
            tmp = dtrtrs(g.L, Kx)                                                    # :277
            var = g.K(x_new) - np.dot(tmp.T, tmp)                                    # :275,278 (Kxx = K(Xnew), tdot)
            mu = mu + g.ymean                                                        # gp.py:847-848, normalizer.py:67-68
            if z is None:
                y = np.random.multivariate_normal(mu.flatten(), var, 1).T            # gp.py:850-857
            else:
                y = mu + np.sqrt(np.abs(var)) * z[j]
            self.Y_aux[j] = np.vstack((self.Y_aux[j], y))                            # uTS.py:46
            self.gps[j] = self._refit(g, self.X_aux, self.Y_aux[j])                  # uTS.py:47
            ys.append(float(y[0, 0]))
            mus.append(float(mu[0, 0]))
            vs.append(float(var[0, 0]))
        return np.array(ys), np.array(mus), np.array(vs)

    def objective(self, x, utility_func, theta, z=None):
        """uTS.py:48-50: minus the utility of the drawn attribute vector."""
        y, _, _ = self.evaluate(x, z)
        return -np.atleast_1d(utility_func(np.squeeze(y), theta))
