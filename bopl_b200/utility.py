"""Drop-ins for `bopl/utility/utility.py:4-52` (Utility) and `bopl/utility/utility_distribution.py:7-108`
(UtilityDistribution), plus the `kind` tag the GPU kernels need.

In the reference a utility is an arbitrary Python callable; a CUDA kernel needs to know which closed form it is.
`Utility(..., kind=...)` accepts 'linear' / 'quadratic' / 'exponential' (the callables of the named experiment
configs: experiments/test_dtlz1a_linear.py:51-55, test_dtlz2a_quad.py:51-57, test_vlmop3_exp.py:41-52) and 'constrained'
(objective with threshold constraints, experiments/test_portfolio2.py:64-79, used by uEI_constrained).  With
`kind=None` the callable is probed against the closed forms on random points (`infer_kind`); anything else raises —
there is no CPU fallback for arbitrary callables.
"""
import numpy as np

UTILITY_KINDS = ("linear", "quadratic", "exponential", "constrained")


def preference_encoder(a, b):
    """bopl/core/preference_encoder.py:6-15."""
    a_aux, b_aux = np.squeeze(a), np.squeeze(b)
    if a_aux > b_aux:
        return 1
    if a_aux < b_aux:
        return -1
    return 0


def closed_form(kind, y, theta, m):
    """Value of the tagged utility at y (m,) — host-side, used for kind inference and preference bookkeeping only."""
    y = np.asarray(y, dtype=np.float64)
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    if kind == "linear":
        return float(np.dot(th, y))
    if kind == "quadratic":
        return float(-np.sum(np.square(y - th)))
    if kind == "exponential":
        return float(np.sum(1.0 - np.exp(-th[0] * y)) / th[0] / m)
    if kind == "constrained":
        return float(y[0]) if np.all(th < y[1:]) else -np.inf
    raise ValueError(kind)


def infer_kind(func, m, theta, n_probe=6, rtol=1e-10, seed=12345):
    """Identify `func` as one of the three closed forms by probing it at random y with the given theta."""
    rs = np.random.RandomState(seed)
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    cands = []
    if th.size == m:
        cands += ["linear", "quadratic"]
    if th.size == 1:
        cands += ["exponential"]
    if m >= 2 and th.size == m - 1:
        cands += ["constrained"]
    ys = rs.normal(size=(n_probe, m))
    for kind in cands:
        ok = True
        for y in ys:
            try:
                got = float(np.squeeze(func(y, theta)))
            except Exception:
                ok = False
                break
            want = closed_form(kind, y, th, m)
            if np.isinf(want) and got == want:
                continue
            if not np.isfinite(got) or abs(got - want) > rtol * max(1.0, abs(want)):
                ok = False
                break
        if ok:
            return kind
    raise NotImplementedError("utility callable is none of linear / quadratic / exponential: the GPU path has no "
                              "kernel for it and there is no CPU fallback (pass kind=... if it is one of them)")


class Utility(object):
    """Same surface as the reference's Utility; adds `.kind`."""

    def __init__(self, func, gradient=None, parameter_distribution=None, expectation=None, affine=False, kind=None):
        self.func = func
        self.gradient = gradient
        self.parameter_distribution = parameter_distribution
        self.expectation = expectation
        self.affine = affine
        self.counter = 0
        if kind is not None and kind not in UTILITY_KINDS:
            raise ValueError("kind must be one of %s" % (UTILITY_KINDS,))
        self.kind = kind

    def resolve_kind(self, m, theta):
        if self.kind is None:
            self.kind = infer_kind(self.func, m, theta)
        return self.kind

    def eval_func_and_gradient(self, y, theta):
        return self.eval_func(y, theta), self.eval_gradient(y, theta)

    def eval_func(self, y, theta):
        return self.func(y, theta)

    def eval_gradient(self, y, theta):
        return self.gradient(y, theta)

    def sample_parameter_from_prior(self, number_of_samples=1, seed=None):
        return self.parameter_distribution.sample_from_prior(number_of_samples, seed)

    def sample_parameter(self, number_of_samples=1):
        return self.parameter_distribution.sample(number_of_samples)

    def update_parameter_distribution(self, underlying_true_utility_func, Y, number_of_pairwise_comparisons=1):
        self.parameter_distribution.add_preference_information(underlying_true_utility_func, Y,
                                                               number_of_pairwise_comparisons)


class UtilityDistribution(object):
    """Finite support with probabilities, or a prior sampler with rejection on stored pairwise preferences."""

    def __init__(self, support=None, prob_dist=None, prior_sample_generator=None, use_full_support=None,
                 utility_func=None, elicitation_strategy=None):
        if support is None and prior_sample_generator is None:
            raise Exception('Either a finite support or a sample generator have to be provided.')
        self.support = support
        self.prior_prob_dist = np.copy(prob_dist)
        self.prob_dist = np.copy(prob_dist)
        self.prior_sample_generator = prior_sample_generator
        self.utility_func = utility_func
        self.elicitation_strategy = elicitation_strategy
        self.preference_information = []
        if support is not None:
            self.support_cardinality = len(support)
        if use_full_support is not None:
            self.use_full_support = use_full_support
        else:                                                       # utility_distribution.py:27-30
            self.use_full_support = support is not None and self.support_cardinality < 20

    def sample_from_prior(self, number_of_samples=1, seed=None):
        if self.support is None:
            samples = self.prior_sample_generator(number_of_samples, seed)
        else:
            rng = np.random.RandomState(seed) if seed is not None else np.random
            idx = rng.choice(int(len(self.support)), size=number_of_samples, p=self.prior_prob_dist)
            samples = self.support[idx, :]
        if number_of_samples > 1 and samples.ndim < 2:
            samples = samples.reshape((number_of_samples,))
        return samples

    def _agrees(self, theta):
        for item in self.preference_information:
            u1 = self.utility_func(item[0], theta)
            u2 = self.utility_func(item[1], theta)
            if preference_encoder(u1, u2) != item[2]:
                return False
        return True

    def sample(self, number_of_samples=1):
        if self.support is not None:
            idx = np.random.choice(int(len(self.support)), size=number_of_samples, p=self.prob_dist)
            return self.support[idx, :]
        if len(self.preference_information) == 0:
            return self.sample_from_prior(number_of_samples)
        kept = []
        while len(kept) < number_of_samples:                        # rejection on the stored comparisons (:48-77)
            theta = self.sample_from_prior(1)[0]
            if self._agrees(theta):
                kept.append(theta)
        return np.atleast_2d(kept)

    def add_preference_information(self, underlying_true_utility_func, Y, number_of_pairwise_comparisons=1):
        if self.elicitation_strategy is None:
            return
        for _ in range(number_of_pairwise_comparisons):
            pair = self.elicitation_strategy(Y=Y)
            pref = preference_encoder(underlying_true_utility_func(pair[0]), underlying_true_utility_func(pair[1]))
            if self.support is not None:
                keep = [i for i in range(len(self.support))
                        if preference_encoder(self.utility_func(pair[0], self.support[i]),
                                              self.utility_func(pair[1], self.support[i])) == pref]
                self.support = self.support[keep]
                self.prob_dist = self.prob_dist[keep]
                self.prob_dist /= np.sum(self.prob_dist)
            pair.append(pref)
            self.preference_information.append(pair)
