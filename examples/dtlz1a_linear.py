#!/usr/bin/env python
"""A BO loop with the drop-in classes, shaped like the reference's experiments/test_dtlz1a_linear.py (DTLZ1a, d=6, two
attributes, linear utility with a Dirichlet-distributed parameter, uEI acquisition, L-BFGS acquisition optimizer).

The loop itself (the reference's bopl/core/bopl.py) is outside the accelerated path; this script only shows that the
pieces a BOPL experiment builds — model, utility, acquisition, acquisition optimizer — are the reference's, with the
posterior / acquisition arithmetic on the GPU.   python examples/dtlz1a_linear.py [iterations]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bopl_b200 import AcquisitionOptimizer, MultiOutputGP, Utility, UtilityDistribution, uEI   # noqa: E402
from bopl_b200.anchor_points import BoxSpace, samples_multidimensional_uniform                   # noqa: E402


def dtlz1a(X):
    """Two-attribute DTLZ1a, sign-flipped (the reference maximises)."""
    X = np.atleast_2d(X)
    g = 100.0 * (5 + np.sum((X[:, 1:] - 0.5) ** 2, axis=1) - np.sum(np.cos(2 * np.pi * (X[:, 1:] - 0.5)), axis=1))
    return -np.stack([0.5 * X[:, 0] * (1 + g), 0.5 * (1 - X[:, 0]) * (1 + g)])


def run(iterations=5, seed=0, n_samples=3, fit_options=None, verbose=True):
    np.random.seed(seed)
    d, m = 6, 2
    space = BoxSpace([(0.0, 1.0)] * d)

    def utility_func(y, parameter):                      # experiments/test_dtlz1a_linear.py:51-55
        return np.dot(parameter, y)

    def utility_gradient(y, parameter):
        return parameter

    def prior_sample_generator(n, seed=None):            # :57-63
        rs = np.random.RandomState(seed) if seed is not None else np.random
        return rs.dirichlet(np.ones(m), n)

    dist = UtilityDistribution(prior_sample_generator=prior_sample_generator, utility_func=utility_func)
    utility = Utility(func=utility_func, gradient=utility_gradient, parameter_distribution=dist, affine=True)
    true_theta = np.array([0.4, 0.6])

    model = MultiOutputGP(output_dim=m, exact_feval=[True] * m, n_samples=n_samples,
                          fit_options=fit_options or dict(n_burnin=30, subsample_interval=5, leapfrog_steps=10))
    optimizer = AcquisitionOptimizer(space, model=model, utility=utility, n_starting=400, n_anchor=8)
    acquisition = uEI(model, space, optimizer=optimizer, utility=utility)

    X = samples_multidimensional_uniform(space.get_bounds(), 2 * (d + 1))
    Y = dtlz1a(X)
    history = [float(np.max(true_theta @ Y))]
    for it in range(iterations):
        model.updateModel(X, [Y[j][:, None] for j in range(m)])
        acquisition.update_samples()
        x_next, acq_val = acquisition.optimize()
        X = np.vstack([X, x_next])
        Y = np.hstack([Y, dtlz1a(x_next)])
        history.append(float(np.max(true_theta @ Y)))
        if verbose:
            print("iter %d  acq %.4g  best true utility %.4f  (%d L-BFGS batches)"
                  % (it, -float(acq_val), history[-1], optimizer.last_stats.get("f_df_batches") or 0))
    return X, Y, history


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
