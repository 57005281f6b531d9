"""Shared test helpers: build the oracle model from a `bopl_b200.synthetic.make_problem` dict."""
import numpy as np

from oracle.gp import ExactGP, MultiOutputGPOracle


def oracle_model(p):
    gps = [[ExactGP(p["X"], p["Y"][j], kind=p["kinds"][j], variance=p["variance"][j, h],
                    lengthscale=p["lengthscale"][j, h], noise=p["noise"][j, h])
            for h in range(p["H"])] for j in range(p["m"])]
    return MultiOutputGPOracle(gps)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
