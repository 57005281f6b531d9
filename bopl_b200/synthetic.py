"""Seeded synthetic problems of the shapes BASELINE.json names.  NumPy only.

`make_problem(name)` returns a dict of plain arrays — the inputs both the CUDA path and the oracle
consume — for:
  "S"             north-star config: d=10, m=3, n=2000, Matern52 ARD, l~U[0.5,0.9], noise 1e-6, H=1
  "dtlz1a_linear" d=6, m=2, linear utility, Dirichlet theta   (experiments/test_dtlz1a_linear.py:23-63)
  "dtlz2a_quad"   d=5, m=4, quadratic utility, 8-point support (experiments/test_dtlz2a_quad.py:24-72)
  "vlmop3_exp"    d=2, m=3, exponential utility, theta in [0.1,0.5] (experiments/test_vlmop3_exp.py:24-63)
  "portfolio1"    d=3, m=5, linear utility theta=(beta^j), synthetic Y (objective needs cvxportfolio + missing blobs)
The objective formulas are the standard DTLZ1a / DTLZ2a / VLMOP3 test functions, sign-flipped
(the reference maximises).  Hyper-parameters are drawn, not fitted (model fitting is out of scope).
"""
import numpy as np

UTILITY_OF = {"S": "linear", "dtlz1a_linear": "linear", "dtlz2a_quad": "quadratic",
              "vlmop3_exp": "exponential", "portfolio1": "linear"}


def _dtlz1a(X, m=2):
    g = 100.0 * (5 + np.sum((X[:, 1:] - 0.5) ** 2, axis=1) - np.sum(np.cos(2 * np.pi * (X[:, 1:] - 0.5)), axis=1))
    return -np.stack([0.5 * X[:, 0] * (1 + g), 0.5 * (1 - X[:, 0]) * (1 + g)])


def _dtlz2a(X, m=4, k=2):
    d = m + k - 1
    g = np.sum((X[:, d - k:] - 0.5) ** 2, axis=1)
    c = np.cos(0.5 * np.pi * X)
    s = np.sin(0.5 * np.pi * X)
    f = np.stack([c[:, 0] * c[:, 1] * c[:, 2], c[:, 0] * c[:, 1] * s[:, 2], c[:, 0] * s[:, 1], s[:, 0]])
    return -f * (1 + g)


def _vlmop3(X):
    q = X[:, 0] ** 2 + X[:, 1] ** 2
    f0 = 0.5 * q + np.sin(q)
    f1 = (3 * X[:, 0] - 2 * X[:, 1] + 4) ** 2 / 8 + (X[:, 0] - X[:, 1] + 1) ** 2 / 27 + 15
    f2 = 1 / (q + 1) - 1.1 * np.exp(X[:, 0] ** 2 - X[:, 1] ** 2)
    return -np.stack([f0, f1, f2])


def _smooth(X, m, rs):
    a = rs.normal(size=(m, X.shape[1]))
    b = rs.normal(size=(m, X.shape[1]))
    return np.sin(2 * np.pi * X @ a.T).T + 0.1 * (X @ b.T).T


def make_problem(name="S", n=None, N=None, n_w=None, L=None, H=1, kind="Mat52", utility=None,
                 seed=0, ell_range=(0.5, 0.9), noise=1e-6):
    """Return dict(X, Y (m,n), lo, hi, kinds, variance (m,H), lengthscale (m,H,d), noise (m,H),
    Xc (N,d), Z (n_w,m), theta (L,p), prob (L,) or None, use_full_support, utility)."""
    rs = lambda k: np.random.RandomState(1000 * seed + k)
    if name == "S":
        d, m, lo, hi = 10, 3, 0.0, 1.0
        n = 2000 if n is None else n
        N = 2 ** 20 if N is None else N
        n_w = 64 if n_w is None else n_w
        L = 32 if L is None else L
    elif name == "dtlz1a_linear":
        d, m, lo, hi = 6, 2, 0.0, 1.0
    elif name == "dtlz2a_quad":
        d, m, lo, hi = 5, 4, 0.0, 1.0
    elif name == "vlmop3_exp":
        d, m, lo, hi = 2, 3, -3.0, 3.0
    elif name == "portfolio1":
        d, m, lo, hi = 3, 5, 0.0, 1.0
    else:
        raise ValueError(name)
    n = 2 * (d + 1) + 50 if n is None else n           # reference: 2(d+1) initial points + up to 150 iterations
    N = 400 if N is None else N                        # acquisition_optimizer.py:26 n_starting
    n_w = 25 if n_w is None else n_w                   # uEI.py:31
    L = 10 if L is None else L                         # uEI.py:38
    utility = UTILITY_OF[name] if utility is None else utility

    X = lo + (hi - lo) * rs(0).uniform(size=(n, d))
    if name == "S" or name == "portfolio1":
        Y = _smooth(X, m, rs(1))
    elif name == "dtlz1a_linear":
        Y = _dtlz1a(X)
    elif name == "dtlz2a_quad":
        Y = _dtlz2a(X)
    else:
        Y = _vlmop3(X)
    span = hi - lo
    r2 = rs(2)
    lengthscale = span * r2.uniform(ell_range[0], ell_range[1], size=(m, H, d))
    variance = np.ones((m, H)) if H == 1 else r2.uniform(0.8, 1.5, size=(m, H))
    if name != "S":
        variance = variance * np.maximum(Y.var(axis=1), 1e-3)[:, None]
    noise_a = np.full((m, H), noise)
    Xc = lo + (hi - lo) * rs(3).uniform(size=(N, d))
    Z = rs(4).normal(size=(n_w, m))
    r5 = rs(5)
    prob, full = None, False
    if utility == "linear":
        if name == "portfolio1":
            beta = 0.5 * r5.uniform(size=L) + 0.5
            theta = np.stack([beta ** j for j in range(m)], axis=1)
        else:
            theta = r5.dirichlet(np.ones(m), L)
    elif utility == "quadratic":
        if name == "dtlz2a_quad":
            g = np.array(np.meshgrid([0., 1. / 3], [1. / 3, 2. / 3], [2. / 3, 1.], [0.5], [0.5])).reshape(5, -1).T
            theta = _dtlz2a(g).T
            prob, full = np.ones(len(theta)) / len(theta), True
        else:
            theta = Y[:, r5.randint(0, n, size=L)].T.copy()
            prob = np.ones(L) / L
    elif utility == "exponential":
        theta = (0.1 + 0.4 * r5.uniform(size=L)).reshape(L, 1)
    elif utility == "constrained":       # attribute 0 = objective, attributes 1.. must exceed theta (test_portfolio2.py:64-79)
        theta = Y[1:].mean(axis=1) - 0.5 * Y[1:].std(axis=1) * np.abs(r5.normal(size=(L, m - 1)))
    else:
        raise ValueError(utility)
    return dict(name=name, d=d, m=m, n=n, N=N, H=H, X=X, Y=Y, lo=lo, hi=hi, kinds=[kind] * m,
                variance=variance, lengthscale=lengthscale, noise=noise_a, Xc=Xc, Z=Z,
                theta=np.ascontiguousarray(theta), prob=prob, use_full_support=full, utility=utility)
