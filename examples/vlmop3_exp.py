#!/usr/bin/env python
"""A BO loop shaped like the reference's experiments/test_vlmop3_exp.py (VLMOP3, d=2, three attributes, exponential utility
with a scalar parameter) with the sampling policy chosen on the command line:

    python examples/vlmop3_exp.py [uTS|uEI|ParEGO|Random] [iterations]

uTS maximises one posterior sample path per iteration (up to 600 objective evaluations, each a rank-1 growth of the factor
on the device); uEI is the Monte-Carlo acquisition; ParEGO fits a one-attribute GP to a random Chebyshev scalarisation.
The loop itself (the reference's bopl/core/bopl.py) is outside the accelerated path.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bopl_b200 import (AcquisitionFunction, AcquisitionOptimizer, BasicModel, MultiOutputGP, ParEGO, Random, Utility,   # noqa: E402
                       UtilityDistribution, uEI, uTS)
from bopl_b200.anchor_points import BoxSpace, samples_multidimensional_uniform                                          # noqa: E402


def vlmop3(X):
    """VLMOP3 (Van Veldhuizen & Lamont), sign-flipped: the reference maximises."""
    X = np.atleast_2d(X)
    x, y = X[:, 0], X[:, 1]
    s = x ** 2 + y ** 2
    f1 = 0.5 * s + np.sin(s)
    f2 = (3 * x - 2 * y + 4) ** 2 / 8.0 + (x - y + 1) ** 2 / 27.0 + 15.0
    f3 = 1.0 / (s + 1.0) - 1.1 * np.exp(-s)
    return -np.stack([f1, f2, f3])


def run(policy_name="uTS", iterations=4, seed=0, verbose=True):
    np.random.seed(seed)
    d, m = 2, 3
    space = BoxSpace([(-3.0, 3.0)] * d)

    def utility_func(y, theta):                          # experiments/test_vlmop3_exp.py:41-46
        theta_aux, y_aux = np.squeeze(theta), np.squeeze(y)
        return np.sum(1. - np.exp(-theta_aux * y_aux), axis=0) / theta_aux / m

    def prior_sample_generator(n, seed=None):            # :55-63
        rs = np.random.RandomState(seed) if seed is not None else np.random
        return 0.1 + 0.4 * rs.rand(n, 1)

    dist = UtilityDistribution(prior_sample_generator=prior_sample_generator, utility_func=utility_func)
    utility = Utility(func=utility_func, parameter_distribution=dist, kind="exponential")
    true_theta = np.array([0.25])
    fit = dict(n_burnin=20, subsample_interval=4, leapfrog_steps=8)
    if policy_name in ("uTS", "uEI"):
        model = MultiOutputGP(output_dim=m, exact_feval=[True] * m, n_samples=3, fit_options=fit)
    else:
        model = BasicModel(output_dim=m)
    if policy_name == "uTS":
        policy = uTS(model, space, optimizer='CMA', utility=utility)
    elif policy_name == "uEI":
        opt = AcquisitionOptimizer(space, model=model, utility=utility, n_starting=400, n_anchor=8)
        policy = AcquisitionFunction(model, space, uEI(model, space, optimizer=opt, utility=utility))
    elif policy_name == "ParEGO":
        policy = ParEGO(model, space, utility, aux_model_options=dict(n_samples=3, fit_options=fit))
    else:
        policy = Random(None, space)
    X = samples_multidimensional_uniform(space.get_bounds(), 2 * (d + 1))
    Y = vlmop3(X)
    true_u = lambda Y: np.array([utility_func(Y[:, i], true_theta) for i in range(Y.shape[1])])
    history = [float(np.max(true_u(Y)))]
    for it in range(iterations):
        t0 = time.perf_counter()
        model.updateModel(X, [Y[j][:, None] for j in range(m)])
        t1 = time.perf_counter()
        x_next = policy.suggest_sample()
        t2 = time.perf_counter()
        X = np.vstack([X, x_next])
        Y = np.hstack([Y, vlmop3(x_next)])
        history.append(float(np.max(true_u(Y))))
        if verbose:
            print("iter %d  %-6s model update %.2f s  suggestion %.2f s  best true utility %.4f"
                  % (it, policy_name, t1 - t0, t2 - t1, history[-1]))
    return X, Y, history


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "uTS", int(sys.argv[2]) if len(sys.argv) > 2 else 4)
