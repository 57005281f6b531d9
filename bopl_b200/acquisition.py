"""Drop-ins for the acquisition functions of the hot path, evaluated on the GPU:

  AcquisitionBase  GPyOpt/acquisitions/base.py:6-74          (sign flip; `optimize` hands f / f_df to the optimizer)
  uEI              bopl/sampling_policies/acquisition_functions/uEI.py:11-177
  uEI_affine       bopl/sampling_policies/acquisition_functions/uEI_affine.py:10-150
  uEI_constrained  bopl/sampling_policies/acquisition_functions/uEI_constrained.py:8-164

Same constructor signatures, attributes (`W_samples`, `utility_parameter_samples`, `utility_prob_dist`,
`number_of_gp_hyps_samples`, `use_full_support`) and methods (`_compute_acq`, `_compute_acq_withGradients`,
`update_samples`, `acquisition_function`, `acquisition_function_withGradients`, `optimize`).  The pathos pool
(uEI.py:86-98) is gone: one call evaluates the whole batch.

Incumbent semantics follow the reference exactly: the batch value path (len(X) > 1) takes the
incumbent under each hyper-sample h (uEI.py:93,104); the single-point value path and the gradient path take it once,
under the model's current h, before the h-loop (uEI.py:67,139).  As in the reference the model is left pointing at
hyper-sample n_h-1 after a call.
"""
import numpy as np


def constant_cost_withGradients(x):
    """GPyOpt/core/task/cost.py: unit cost, zero gradient."""
    return np.ones(x.shape[0])[:, None], np.zeros(x.shape)


class AcquisitionBase(object):
    analytical_gradient_prediction = False

    def __init__(self, model, space, optimizer, cost_withGradients=None):
        self.model = model
        self.space = space
        self.optimizer = optimizer
        self.analytical_gradient_acq = self.analytical_gradient_prediction and self.model.analytical_gradient_prediction
        self.cost_withGradients = constant_cost_withGradients if cost_withGradients is None else cost_withGradients

    def acquisition_function(self, x):                         # base.py:33-40
        return -self._compute_acq(x)

    def acquisition_function_withGradients(self, x):           # base.py:43-56
        f_acqu, df_acqu = self._compute_acq_withGradients(x)
        return -f_acqu, -df_acqu

    def optimize(self, duplicate_manager=None):                # base.py:58-66
        if not self.analytical_gradient_acq:
            return self.optimizer.optimize(f=self.acquisition_function, duplicate_manager=duplicate_manager)
        return self.optimizer.optimize(f=self.acquisition_function, f_df=self.acquisition_function_withGradients,
                                       duplicate_manager=duplicate_manager)

    def _compute_acq(self, x):
        raise NotImplementedError('')

    def _compute_acq_withGradients(self, x):
        raise NotImplementedError('')


class _UtilityAcquisition(AcquisitionBase):
    """Shared plumbing: theta / weight arrays, incumbents on the device."""

    analytical_gradient_prediction = True

    def _thetas(self):
        src = self.utility_parameter_samples
        if isinstance(src, np.ndarray) and src.dtype == np.float64 and src.ndim == 2 and src.flags.c_contiguous:
            return src                                         # same object -> its device copy is reused
        th = np.asarray(src, dtype=np.float64)
        return np.ascontiguousarray(th.reshape(len(src), -1))

    def _weights(self, L):
        if self.use_full_support:                              # uEI.py:56-57
            return np.ascontiguousarray(self.utility_prob_dist, dtype=np.float64).reshape(L)
        return np.full(L, 1.0 / L)                             # uEI.py:59

    def _dev_weights(self, wts):
        """Uniform weights 1/L: one cached device vector per L."""
        cache = self.__dict__.setdefault("_dev_cache", {})
        key = ("uniform", len(wts))
        if key not in cache or self.use_full_support:
            cache[key] = (None, self.model.engine.to_dev(wts))
        return cache[key][1]

    def _incumbents(self, kind, theta_dev, per_h):
        """(n_h, L) device tensor of best-so-far utilities.  per_h=False replicates the model's current h."""
        eng = self.model.engine
        best_all = eng.incumbent_dev(kind, theta_dev)          # (H, L)
        n_h = self.number_of_gp_hyps_samples
        if per_h:
            return best_all[:n_h].contiguous()
        return best_all[self.model._h:self.model._h + 1].expand(n_h, -1).contiguous()

    def _x(self, X):
        return np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)

    def _dev(self, name, array):
        """Device copy of a small host array that only changes in `update_samples` (W_samples, theta, weights): uploaded
        once per array OBJECT — the L-BFGS inner loop calls thousands of times between two updates.  Arrays replaced by
        assignment are picked up (identity check); arrays mutated in place must be re-assigned."""
        cache = self.__dict__.setdefault("_dev_cache", {})
        hit = cache.get(name)
        if hit is not None and hit[0] is array:
            return hit[1]
        t = self.model.engine.to_dev(array)
        cache[name] = (array, t)
        return t


class uEI(_UtilityAcquisition):
    """Monte-Carlo expected improvement of the utility under utility-parameter uncertainty."""

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, utility=None):
        self.optimizer = optimizer
        self.utility = utility
        super(uEI, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.cost_withGradients = constant_cost_withGradients            # uEI.py:25-29
        self.n_attributes = self.model.output_dim
        self.W_samples = np.random.normal(size=(25, self.n_attributes))  # uEI.py:31
        self.number_of_gp_hyps_samples = min(10, self.model.number_of_hyps_samples())
        self.use_full_support = self.utility.parameter_distribution.use_full_support
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)

    def _prepare(self, per_h):
        theta = self._thetas()
        kind = self.utility.resolve_kind(self.n_attributes, theta[0])
        th_d = self._dev("theta", self.utility_parameter_samples) if theta is self.utility_parameter_samples \
            else self.model.engine.to_dev(theta)
        best = self._incumbents(kind, th_d, per_h)
        return kind, theta, th_d, self._weights(len(theta)), best

    def _compute_acq(self, X, parallel=True):
        """uEI.py:40-62 -> (N, 1)."""
        X = self._x(X)
        n_h = self.number_of_gp_hyps_samples
        theta = self._thetas()
        kind = self.utility.resolve_kind(self.n_attributes, theta[0])
        best = self._incumbents_host(kind, theta, per_h=bool(parallel and len(X) > 1))
        acq = self.model.engine.uei_host(X, self.W_samples, kind, theta, self._weights(len(theta)), best, n_h)
        self.model.set_hyperparameters(n_h - 1)
        return acq.reshape(X.shape[0], 1)

    def top_candidates(self, X, k, index_offset=0, want_acq=False, out=None):
        """Batch value + the k best candidates in one call (what AnchorPointsGenerator.get does with argsort):
        returns (values (k,), global indices (k,)[, acq (N,)]).  X may be a pinned torch CPU tensor (no staging copy);
        `out` an optional (pinned) buffer for the scores."""
        X = X if hasattr(X, "data_ptr") else self._x(X)
        n_h = self.number_of_gp_hyps_samples
        theta = self._thetas()
        kind = self.utility.resolve_kind(self.n_attributes, theta[0])
        best = self._incumbents_host(kind, theta, per_h=len(X) > 1)
        acq, val, idx = self.model.engine.uei_host(X, self.W_samples, kind, theta, self._weights(len(theta)), best, n_h, k=k,
                                                   index_offset=index_offset, want_acq=want_acq, out=out)
        self.model.set_hyperparameters(n_h - 1)
        return (val, idx, acq) if want_acq else (val, idx)

    def _incumbents_host(self, kind, theta, per_h):
        """Host copy of the (n_h, L) incumbents, cached while theta, the model state and (for the stale variant) the
        current hyper-sample stay the same — thousands of L-BFGS calls happen between two `update_samples`."""
        key = (kind, per_h, None if per_h else self.model._h, self.number_of_gp_hyps_samples)
        cache = self.__dict__.setdefault("_best_cache", {})
        hit = cache.get(per_h)
        if hit is not None and hit[0] == key and hit[1] is theta and hit[3] is self.model.state:
            return hit[2]
        th_d = self._dev("theta", theta) if theta is self.utility_parameter_samples else self.model.engine.to_dev(theta)
        best = self._incumbents(kind, th_d, per_h).cpu().numpy()
        cache[per_h] = (key, theta, best, self.model.state)
        return best

    def _compute_acq_withGradients(self, X):
        """uEI.py:119-134 -> (N, 1), (N, d)."""
        X = self._x(X)
        n_h = self.number_of_gp_hyps_samples
        theta = self._thetas()
        kind = self.utility.resolve_kind(self.n_attributes, theta[0])
        best = self._incumbents_host(kind, theta, per_h=False)
        acq, dacq = self.model.engine.uei_grad_host(X, self.W_samples, kind, theta, self._weights(len(theta)), best, n_h)
        self.model.set_hyperparameters(n_h - 1)
        return acq.reshape(X.shape[0], 1), dacq.reshape(X.shape)

    def update_samples(self):                                            # uEI.py:170-177
        self.W_samples = np.random.normal(size=self.W_samples.shape)
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)


class uEI_affine(_UtilityAcquisition):
    """Closed-form expected improvement for a linear utility theta . f (uEI_affine.py:10-150).
    The incumbent is recomputed under every hyper-sample (uEI_affine.py:80,101)."""

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, utility=None):
        self.optimizer = optimizer
        self.utility = utility
        super(uEI_affine, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.cost_withGradients = constant_cost_withGradients
        self.number_of_gp_hyps_samples = min(10, self.model.number_of_hyps_samples())
        self.use_full_support = self.utility.parameter_distribution.use_full_support
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)

    def _run(self, X, with_grad):
        X = self._x(X)
        eng = self.model.engine
        n_h = self.number_of_gp_hyps_samples
        theta = self._thetas()
        if theta.shape[1] != self.model.output_dim:
            raise ValueError("uEI_affine needs one utility parameter per attribute")
        th_d = eng.to_dev(theta)
        best = self._incumbents("linear", th_d, per_h=True)
        acq, dacq = eng.uei_affine_dev(eng.to_dev(X), th_d, eng.to_dev(self._weights(len(theta))), best, n_h, with_grad)
        self.model.set_hyperparameters(n_h - 1)
        acq = acq.cpu().numpy().reshape(X.shape[0], 1)
        return (acq, dacq.cpu().numpy().reshape(X.shape)) if with_grad else acq

    def _compute_acq(self, X):
        return self._run(X, False)

    def _compute_acq_withGradients(self, X):
        return self._run(X, True)

    def update_samples(self):                                            # uEI_affine.py:144-150
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)


class uEI_constrained(_UtilityAcquisition):
    """Closed-form constrained EI (uEI_constrained.py:8-164): EI of attribute 0 times the feasibility probabilities of
    attributes 1.. against the thresholds theta.  As in the reference the incumbent ("max feasible value so far") is
    computed ONCE, under the model's current hyper-sample, before the h-loop (uEI_constrained.py:52,98)."""

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, utility=None):
        self.optimizer = optimizer
        self.utility = utility
        super(uEI_constrained, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.use_full_support = self.utility.parameter_distribution.use_full_support
        self.number_of_gp_hyps_samples = min(10, self.model.number_of_hyps_samples())   # the reference hard-codes 10
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)

    def _run(self, X, with_grad):
        X = self._x(X)
        eng = self.model.engine
        n_h = self.number_of_gp_hyps_samples
        theta = self._thetas()
        if theta.shape[1] != self.model.output_dim - 1:
            raise ValueError("uEI_constrained needs one threshold per constraint attribute (m - 1)")
        kind = self.utility.resolve_kind(self.model.output_dim, theta[0])
        if kind != "constrained":
            raise NotImplementedError("uEI_constrained needs the threshold-constraint utility (kind='constrained')")
        th_d = eng.to_dev(theta)
        best = self._incumbents("constrained", th_d, per_h=False)
        acq, dacq = eng.uei_constrained_dev(eng.to_dev(X), th_d, eng.to_dev(self._weights(len(theta))), best, n_h, with_grad)
        self.model.set_hyperparameters(n_h - 1)
        acq = acq.cpu().numpy().reshape(X.shape[0], 1)
        return (acq, dacq.cpu().numpy().reshape(X.shape)) if with_grad else acq

    def _compute_acq(self, X):
        return self._run(X, False)

    def _compute_acq_withGradients(self, X):
        return self._run(X, True)

    def update_samples(self):                                            # uEI_constrained.py:128-134
        if self.use_full_support:
            self.utility_parameter_samples = self.utility.parameter_distribution.support
            self.utility_prob_dist = self.utility.parameter_distribution.prob_dist
        else:
            self.utility_parameter_samples = self.utility.sample_parameter(10)
