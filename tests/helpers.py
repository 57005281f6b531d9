"""Shared test helpers: build the oracle model / the GPU model from a `bopl_b200.synthetic.make_problem` dict."""
import numpy as np

from oracle.gp import ExactGP, MultiOutputGPOracle


def oracle_model(p):
    gps = [[ExactGP(p["X"], p["Y"][j], kind=p["kinds"][j], variance=p["variance"][j, h],
                    lengthscale=p["lengthscale"][j, h], noise=p["noise"][j, h])
            for h in range(p["H"])] for j in range(p["m"])]
    return MultiOutputGPOracle(gps)


def gpu_model(p, gradients=True, device=0, posterior_kernel=None):
    from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
    state = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"],
                                       want_winv=gradients)
    return MultiOutputGP.from_state(state, device=device, posterior_kernel=posterior_kernel)


def make_utility(p):
    """bopl_b200 Utility object for the problem's utility kind, with the oracle's Python callables attached."""
    from bopl_b200.utility import Utility, UtilityDistribution
    from oracle import utilities
    func, grad = utilities.get(p["utility"], p["m"])
    theta = p["theta"]
    if p["use_full_support"]:
        dist = UtilityDistribution(support=theta, prob_dist=np.array(p["prob"]), use_full_support=True)
    else:
        dist = UtilityDistribution(prior_sample_generator=lambda n, seed=None: theta[:n], use_full_support=False)
    return Utility(func=func, gradient=grad, parameter_distribution=dist, kind=p["utility"])


def gpu_acquisition(p, model, cls=None):
    from bopl_b200.acquisition import uEI
    from bopl_b200.anchor_points import BoxSpace
    cls = uEI if cls is None else cls
    util = make_utility(p)
    acq = cls(model, BoxSpace([(p["lo"], p["hi"])] * p["d"]), utility=util)
    acq.utility_parameter_samples = p["theta"]
    if p["use_full_support"]:
        acq.utility_prob_dist = np.array(p["prob"])
    if hasattr(acq, "W_samples"):
        acq.W_samples = p["Z"]
    return acq


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def scaled_err(a, b):
    """max |a-b| / max |b|: error relative to the scale of the array (for quantities with entries near 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
