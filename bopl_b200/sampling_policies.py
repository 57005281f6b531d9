"""Sampling policies of `bopl/sampling_policies/` on the GPU engine: `AcquisitionFunction` (AcquisitionFunction.py:7-36),
`Random` (Random.py:7-26), utility Thompson sampling `uTS` (uTS.py:12-57) and `ParEGO` (ParEGO.py:15-66: a one-attribute
auxiliary GP on a random Chebyshev scalarisation of the observations, maximised with `uEI_affine` — every piece of it is one
of the accelerated classes).

uTS maximises the utility of ONE posterior sample path.  The reference grows the path inside its objective: every
evaluation draws f(x) from a deep copy of each attribute's GP, appends the draw to the copy's data and re-fits the copy
(`set_XY`, O((n+t)^3)), up to 600 times per suggestion (uTS.py:40-55).  Here the path lives on the device
(`bopl_path_begin / bopl_path_eval_host`, csrc/path_kernels.cu): one evaluation is a rank-1 growth of the cached factor —
matrix-vector work, all attributes in the same launches — and costs ~0.1 ms instead of a refit.  The normal draws stay on
the host, taken from `np.random` in the reference's order (one `standard_normal` per attribute and evaluation, which is
what `np.random.multivariate_normal(mean, cov, 1)` consumes for a 1x1 covariance), so a seeded run follows the same path.
"""
import numpy as np

from .anchor_points import initial_design
from .optimization import ContextManager, apply_optimizer, choose_optimizer


class SamplePath(object):
    """A posterior sample path of all attributes under hyper-sample 0, conditioned on its own earlier draws."""

    def __init__(self, model, capacity=640):
        self.model = model
        self.engine = model.engine
        st = model.state
        if st is None or st.Y is None:
            raise ValueError("the model holds no targets (updateModel / a GPState with Y is needed for a sample path)")
        self.m = st.m
        self.X = np.copy(st.X)                                  # grows with the path, like `model_sample[0].X`
        self.Y = [st.Y[j].reshape(-1, 1).copy() for j in range(self.m)]
        self.engine.path_begin(0, st.Y, capacity)               # gpmodel.py:167-168: model_instances[0]
        self._pending = None                                    # (x, y, next attribute) of an evaluation being handed out
        self.attributes = [SamplePathAttribute(self, j) for j in range(self.m)]
        self.last_mean = self.last_var = None

    def evaluate(self, x, z=None):
        """f(x) of every attribute, (m,): drawn given the data and the path so far, then appended to the path."""
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        if z is None:
            z = np.array([np.random.standard_normal(1)[0] for _ in range(self.m)])   # one draw per attribute, in order
        y, mu, var = self.engine.path_eval(x, z)
        self.last_mean, self.last_var = mu, var
        self.X = np.vstack((self.X, x[None, :]))
        for j in range(self.m):
            self.Y[j] = np.vstack((self.Y[j], [[y[j]]]))
        return y

    def __len__(self):
        return self.engine.path_length


class SamplePathAttribute(object):
    """What uTS.py touches on one element of `get_copy_of_model_sample()`: `.X`, `.Y`, `posterior_samples_f`, `set_XY`.
    The m views share one device path; the joint step runs when attribute 0 asks for its draw."""

    def __init__(self, path, j):
        self._path, self._j = path, j

    @property
    def X(self):
        return self._path.X

    @property
    def Y(self):
        return self._path.Y[self._j]

    def posterior_samples_f(self, X, size=1, full_cov=True, **kw):                  # GPy/core/gp.py:833-867
        X = np.atleast_2d(X)
        if X.shape[0] != 1 or size != 1:
            raise NotImplementedError("the device path advances one point and one draw at a time (uTS.py:44)")
        p = self._path
        if p._pending is None or not np.array_equal(p._pending[0], X[0]) or p._pending[2] != self._j:
            if self._j != 0:
                raise RuntimeError("attributes must be sampled in order 0..m-1 at the same point (uTS.py:43-46)")
            p._pending = [X[0].copy(), p.evaluate(X[0]), 0]
        y = p._pending[1][self._j]
        p._pending[2] += 1
        if p._pending[2] == p.m:
            p._pending = None
        return np.array([[y]])

    def set_XY(self, X=None, Y=None):                                               # GPy/core/gp.py:191-227
        """The refit is what the device step already did; only check that the caller appended what was drawn."""
        if Y is not None and (len(Y) != len(self.Y) or Y[-1, 0] != self.Y[-1, 0]):
            raise ValueError("set_XY on a sample-path view accepts only the path's own growth")


class SamplingPolicyBase(object):                                                   # sampling_policies/base.py:5-24
    def __init__(self, model=None, space=None):
        self.model = model
        self.space = space

    def suggest_sample(self, number_of_samples=1):
        raise NotImplementedError('')


class Sequential(object):
    """GPyOpt/core/evaluators/sequential.py: the next point is the acquisition's optimum."""

    def __init__(self, acquisition, batch_size=1):
        self.acquisition = acquisition
        self.batch_size = batch_size

    def compute_batch(self, duplicate_manager=None, context_manager=None):
        x, _ = self.acquisition.optimize(duplicate_manager=duplicate_manager)
        return x


class AcquisitionFunction(SamplingPolicyBase):                                      # AcquisitionFunction.py:7-36
    analytical_gradient_prediction = True

    def __init__(self, model, space, acquisition_func, evaluator=None):
        super(AcquisitionFunction, self).__init__(model, space)
        self.acquisition_func = acquisition_func
        self.evaluator = evaluator if evaluator is not None else Sequential(acquisition_func)

    def suggest_sample(self, number_of_samples=1):
        self.acquisition_func.update_samples()
        self.acquisition_func.optimizer.context_manager = ContextManager(self.space, None)
        x = self.evaluator.compute_batch(duplicate_manager=None)
        return self.space.zip_inputs(x) if hasattr(self.space, "zip_inputs") else np.atleast_2d(x)


class Random(SamplingPolicyBase):                                                   # Random.py:7-26
    analytical_gradient_prediction = True

    def suggest_sample(self, number_of_samples=1):
        return initial_design('random', self.space, 1)


class uTS(SamplingPolicyBase):                                                      # uTS.py:12-57
    analytical_gradient_prediction = True

    def __init__(self, model, space, optimizer, utility, maxfevals=600):
        super(uTS, self).__init__(model, space)
        self.optimizer_name = optimizer
        self.utility = utility
        self.context_manager = ContextManager(self.space)
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        self.maxfevals = maxfevals                                                  # uTS.py:54
        self.X_aux = None
        self.Y_aux = None
        self.last_path_length = 0

    def suggest_sample(self, number_of_samples=1):
        utility_parameter_sample = self.utility.sample_parameter(number_of_samples=1)
        path = SamplePath(self.model, capacity=self.maxfevals + 8)

        def objective_func_sample(x):                                               # uTS.py:40-50
            x_new = np.atleast_2d(x)
            y_new = path.evaluate(x_new[0])
            val = np.atleast_1d(self.utility.eval_func(np.squeeze(y_new), utility_parameter_sample))
            return -val

        try:
            d0 = initial_design('random', self.space, 1)
            suggested_sample = apply_optimizer(self.optimizer, d0, f=objective_func_sample,
                                               context_manager=self.context_manager, space=self.space,
                                               maxfevals=self.maxfevals)[0]
            self.X_aux, self.Y_aux = path.X, path.Y
            self.last_path_length = len(path)
        finally:
            self.model.engine.path_end()
        return np.atleast_2d(suggested_sample)


def chebyshev_scalarization(Y, weight):                                             # others/chebyshev_scalarization.py:3-5
    tmp = np.multiply(weight, Y.T).T
    return np.min(tmp, axis=0) + 0.05 * np.sum(Y, axis=0)


class ParEGO(SamplingPolicyBase):                                                   # ParEGO.py:15-66
    """`model` only has to hand out the evaluations (`get_XY`: the reference pairs ParEGO with `BasicModel`)."""
    analytical_gradient_prediction = True

    def __init__(self, model, space, utility, optimizer='lbfgs', device=0, aux_model_options=None):
        super(ParEGO, self).__init__(model, space)
        self.utility = utility
        self.optimizer = optimizer
        self.device = device
        self.aux_model_options = dict(aux_model_options or {})      # e.g. n_samples / fit_options of the auxiliary GP
        dist = getattr(self.utility, "parameter_distribution", None)
        self.preference_information = getattr(dist, "preference_information", [])

    def suggest_sample(self, number_of_samples=1):
        from .acquisition import uEI_affine
        from .models import MultiOutputGP
        from .optimization import AcquisitionOptimizer
        from .utility import Utility, UtilityDistribution
        X_evaluated, Y_evaluated = self.model.get_XY()
        self.X_aux, self.Y_aux = X_evaluated, Y_evaluated
        weight = np.random.dirichlet(np.ones(len(self.Y_aux)))                      # ParEGO.py:43
        self.Y_aux = np.reshape(self.Y_aux, (len(self.Y_aux), self.Y_aux[0].shape[0]))
        scalarized_fX = chebyshev_scalarization(self.Y_aux, weight)
        scalarized_fX = np.reshape(scalarized_fX, (len(scalarized_fX), 1))
        aux_model = MultiOutputGP(output_dim=1, exact_feval=[True], fixed_hyps=False, device=self.device, **self.aux_model_options)

        def aux_utility_func(y, parameter):
            return np.dot(parameter, y)

        def aux_utility_gradient(y, parameter):
            return parameter

        aux_dist = UtilityDistribution(support=np.ones((1, 1)), prob_dist=np.ones((1,)), utility_func=aux_utility_func)
        aux_utility = Utility(func=aux_utility_func, gradient=aux_utility_gradient, parameter_distribution=aux_dist, affine=True)
        aux_model.updateModel(self.X_aux, [scalarized_fX])                          # core/bopl.py:428-429
        aux_optimizer = AcquisitionOptimizer(space=self.space, model=aux_model, utility=aux_utility, optimizer=self.optimizer,
                                             include_baseline_points=True)
        aux_acquisition = uEI_affine(aux_model, self.space, optimizer=aux_optimizer, utility=aux_utility)
        aux_policy = AcquisitionFunction(aux_model, self.space, aux_acquisition, Sequential(aux_acquisition))
        try:
            return aux_policy.suggest_sample()                                      # core/bopl.py:430
        finally:
            aux_model.engine.close()
