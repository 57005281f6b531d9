"""Oracle: the utility callables of the reference's named experiment configs (TEST INFRASTRUCTURE).

These are the Python functions the reference hands to `Utility(func=..., gradient=...)`:
  linear       experiments/test_dtlz1a_linear.py:51-55 (also test_portfolio1.py:103-107)
  quadratic    experiments/test_dtlz2a_quad.py:51-57
  exponential  experiments/test_vlmop3_exp.py:41-52   (m is captured from the script's scope)
y is (m,) or (m,k); the value is a scalar or (k,).
"""
import numpy as np


def linear_func(y, parameter):
    return np.dot(parameter, y)


def linear_gradient(y, parameter):
    return parameter


def quadratic_func(y, parameter):
    aux = (y.transpose() - parameter).transpose()
    return -np.sum(np.square(aux), axis=0)


def quadratic_gradient(y, parameter):
    y_aux = np.squeeze(y)
    return -2 * (y_aux - parameter)


def make_exponential(m):
    def exponential_func(y, theta):
        theta_aux = np.squeeze(theta)
        y_aux = np.squeeze(y)
        val = np.sum(1.0 - np.exp(-theta_aux * y_aux), axis=0) / theta_aux
        val /= m
        return val

    def exponential_gradient(y, theta):
        theta_aux = np.squeeze(theta)
        y_aux = np.squeeze(y)
        return np.exp(-theta_aux * y_aux) / m

    return exponential_func, exponential_gradient


def constrained_func(y, parameter):
    """experiments/test_portfolio2.py:64-79 generalised to m-1 constraints: y_0 if parameter_j < y_{j+1} for all j."""
    y_aux = np.squeeze(y)
    par = np.atleast_1d(np.squeeze(parameter))
    if y_aux.ndim > 1:
        ok = np.all(par[:, None] < y_aux[1:], axis=0)
        return np.where(ok, y_aux[0], -np.inf)
    return y_aux[0] if np.all(par < y_aux[1:]) else -np.inf


def get(kind, m):
    """(func, gradient) for kind in {'linear','quadratic','exponential'}."""
    if kind == "linear":
        return linear_func, linear_gradient
    if kind == "quadratic":
        return quadratic_func, quadratic_gradient
    if kind == "exponential":
        return make_exponential(m)
    if kind == "constrained":
        return constrained_func, None
    raise ValueError(kind)
