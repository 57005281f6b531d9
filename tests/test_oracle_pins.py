"""CPU tests that pin the ORACLE (test infrastructure) before anything is compared against it.

The reference holds no golden vectors for uEI / MultiOutputGP; the nearest pins are used:
  * closed-form EI known answers, GPyOpt/testing/acquisitions_tests/test_ei_acquisition.py:21-37, reached as the
    m=1, theta=1, H=1 limit of uEI_affine (sign-flipped: stock EI minimises);
  * predict == textbook K** - K*^T (K + s2 I)^-1 K*, GPy/testing/model_tests.py:63-82;
  * finite-difference gradient checks in the style of GPy/testing/kernel_tests.py:352-434 (X 10x6 / X2 20x6);
  * loop-faithful restatement == vectorised restatement; MC uEI -> closed-form uEI_affine as n_w grows;
  * tests/golden/*.npz: outputs of the reference's OWN acquisition code (uEI.py / uEI_affine.py executed from
    /root/reference with stubbed imports, see tests/golden/make_golden.py) on the oracle's GP layer.
"""
import os

import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from oracle import acquisition as oacq
from oracle import gp as ogp
from oracle import utilities
from tests.helpers import oracle_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class MockModel(object):
    """m=1 model with fixed predictions (the reference's EI tests mock the model the same way)."""

    def __init__(self, mean, var, best, dmean=None, dvar=None):
        self.output_dim = 1
        self.mean, self.var, self.best, self.dmean, self.dvar = mean, var, best, dmean, dvar

    def set_hyperparameters(self, h):
        pass

    def predict(self, X):
        N = np.atleast_2d(X).shape[0]
        return np.full((1, N), self.mean), np.full((1, N), self.var)

    def posterior_mean_at_evaluated_points(self):
        return np.array([[self.best, self.best - 1.0]])

    def posterior_mean_gradient(self, X):
        X = np.atleast_2d(X)
        return np.full((1,) + X.shape, self.dmean)

    def posterior_variance_gradient(self, X):
        X = np.atleast_2d(X)
        return np.full((1,) + X.shape, self.dvar)


def test_ei_known_answer_value():
    # stock EI: m=1, s=3, fmin=0.1, jitter=0.01 -> 0.79646919; maximisation form: mean=-(m+jitter), best=-fmin
    model = MockModel(mean=-1.01, var=9.0, best=-0.1)
    X = np.array([[2.0, 2.0], [2.0, 2.0]])
    for fn in (oacq.uei_affine_marginal_loop, oacq.uei_affine_marginal_vec):
        got = fn(model, np.array([[1.0]]), X, 1)
        assert np.isclose(got, 0.79646919).all()


def test_ei_known_answer_gradient():
    # stock EI: m=1, s=1, dmdx=0.1, dsdx=0.1, fmin=0.1, jitter=0.01 -> f=0.0986038, df=0.00822768
    model = MockModel(mean=-1.01, var=1.0, best=-0.1, dmean=-0.1, dvar=0.2)
    X = np.array([[2.0, 2.0], [2.0, 2.0]])
    f, df = oacq.uei_affine_marginal_with_gradient_loop(model, np.array([[1.0]]), X, 1)
    assert np.isclose(f, 0.0986038).all()
    assert np.isclose(df, 0.00822768).all()


@pytest.mark.parametrize("kind", ["rbf", "Mat52", "Mat32", "ExpQuad"])
def test_predict_equals_textbook_formula(kind):
    rs = np.random.RandomState(0)
    X = rs.uniform(-3, 3, (20, 1))
    Y = np.sin(X) + rs.randn(20, 1) * 0.05
    Xn = rs.uniform(-3, 3, (7, 1))
    g = ogp.ExactGP(X, Y, kind=kind, variance=1.3, lengthscale=[0.8], noise=0.5)
    K = g.K(X)
    Kinv = np.linalg.pinv(K + np.eye(20) * (0.5 + 1e-8))
    mu_hat = g.K(Xn, X).dot(Kinv).dot(Y - Y.mean()) + Y.mean()
    var_hat = np.diag(g.K(Xn) - g.K(Xn, X).dot(Kinv).dot(g.K(X, Xn)))[:, None]
    mu, var = g.predict(Xn, include_likelihood=False)
    np.testing.assert_almost_equal(mu_hat, mu)
    np.testing.assert_almost_equal(var_hat, var)
    np.testing.assert_almost_equal(g.posterior_variance(Xn), var + 0.5)
    np.testing.assert_almost_equal(g.posterior_mean(Xn), mu)


@pytest.mark.parametrize("kind", ["rbf", "Mat52", "Mat32", "ExpQuad"])
def test_kernel_gradients_X_finite_difference(kind):
    rs = np.random.RandomState(1)
    X, X2 = rs.randn(10, 6), rs.randn(20, 6)
    ell = rs.uniform(0.7, 1.5, 6)
    dL = rs.randn(10, 20)
    g = ogp.kern_gradients_X(kind, 1.7, ell, dL, X, X2)
    eps = 1e-6
    for i in range(10):
        for q in range(6):
            Xp, Xm = X.copy(), X.copy()
            Xp[i, q] += eps
            Xm[i, q] -= eps
            fd = np.sum(dL * (ogp.kern_K(kind, 1.7, ell, Xp, X2) - ogp.kern_K(kind, 1.7, ell, Xm, X2))) / (2 * eps)
            assert abs(fd - g[i, q]) < 1e-6 * max(1.0, abs(fd))
    K = ogp.kern_K(kind, 1.7, ell, X)
    assert np.all(np.linalg.eigvalsh(K) > -1e-10)         # PSD check, kernel_tests.py:47-53
    assert np.allclose(np.diag(K), 1.7)


def test_posterior_gradients_finite_difference():
    p = make_problem("dtlz2a_quad", N=5)
    om = oracle_model(p)
    X = p["Xc"]
    dm, dv = om.posterior_mean_gradient(X), om.posterior_variance_gradient(X)
    eps = 1e-6
    for q in range(p["d"]):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, q] += eps
        Xm[:, q] -= eps
        fdm = (om.posterior_mean(Xp) - om.posterior_mean(Xm)) / (2 * eps)
        fdv = (om.posterior_variance(Xp) - om.posterior_variance(Xm)) / (2 * eps)
        assert np.allclose(fdm, dm[:, :, q], rtol=1e-5, atol=1e-7)
        assert np.allclose(fdv, dv[:, :, q], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name", ["dtlz1a_linear", "dtlz2a_quad", "vlmop3_exp"])
def test_loop_faithful_equals_vectorised(name):
    p = make_problem(name, N=9, n=20, n_w=5, L=3, H=2)
    om = oracle_model(p)
    func, grad = utilities.get(p["utility"], p["m"])
    th = p["theta"][:3]
    for inc in ("per_h", "stale"):
        om.set_hyperparameters(0)
        a = oacq.uei_marginal_loop(om, func, th, p["Z"], p["Xc"], 2, incumbent=inc)
        om.set_hyperparameters(0)
        b = oacq.uei_marginal_vec(om, p["utility"], th, p["Z"], p["Xc"], 2, incumbent=inc)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)
    om.set_hyperparameters(0)
    f1, d1 = oacq.uei_marginal_with_gradient_loop(om, func, grad, th, p["Z"], p["Xc"], 2)
    om.set_hyperparameters(0)
    f2, d2 = oacq.uei_marginal_with_gradient_vec(om, p["utility"], th, p["Z"], p["Xc"], 2)
    assert np.allclose(f1, f2, rtol=1e-12, atol=1e-14)
    assert np.allclose(d1, d2, rtol=1e-10, atol=1e-13)


def test_uei_gradient_finite_difference():
    p = make_problem("dtlz2a_quad", N=4, n_w=9)
    om = oracle_model(p)
    X = p["Xc"]
    f, d = oacq.uei_marginal_with_gradient_vec(om, "quadratic", p["theta"], p["Z"], X, 1)
    eps = 1e-6
    for q in range(p["d"]):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, q] += eps
        Xm[:, q] -= eps
        fp = oacq.uei_marginal_vec(om, "quadratic", p["theta"], p["Z"], Xp, 1, incumbent="stale")
        fm = oacq.uei_marginal_vec(om, "quadratic", p["theta"], p["Z"], Xm, 1, incumbent="stale")
        assert np.allclose((fp - fm) / (2 * eps), d[:, q, :], rtol=2e-4, atol=1e-7)


def test_mc_uei_converges_to_closed_form_for_linear_utility():
    p = make_problem("dtlz1a_linear", N=40, n_w=200000, L=3)
    om = oracle_model(p)
    mc = oacq.uei_marginal_vec(om, "linear", p["theta"], p["Z"], p["Xc"], 1)
    cf = oacq.uei_affine_marginal_vec(om, p["theta"], p["Xc"], 1)
    assert np.allclose(mc, cf, rtol=0.05, atol=0.02 * cf.max())


def test_top_k_is_stable_argsort():
    a = np.array([0.0, 3.0, 0.0, 3.0, 1.0, -1.0])
    idx, sc = oacq.top_k(-a, 4)
    assert idx.tolist() == [1, 3, 4, 0] and sc.tolist() == [-3.0, -3.0, -1.0, -0.0]


def _golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else []


@pytest.mark.parametrize("fname", _golden_files())
def test_oracle_against_reference_generated_golden(fname):
    """Fixtures made by running the reference's own uEI.py / uEI_affine.py code (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    name, H = str(g["name"]), int(g["H"])
    if str(g["acq"]) == "uTS":
        # the reference's own uTS.suggest_sample closure (uTS.py:40-50) on its own GP.posterior_samples_f / set_XY
        from oracle.thompson import SamplePathOracle
        p = make_problem(name, N=4, n=int(g["n"]), n_w=2, L=len(g["theta"]) if name != "dtlz2a_quad" else 8, H=H)
        om = oracle_model(p)
        func, _ = utilities.get(p["utility"], p["m"])
        sp = SamplePathOracle([om.gps[j][0] for j in range(p["m"])])
        np.random.seed(int(g["seed"]))                               # same global stream as the reference run
        for t, x in enumerate(g["Xseq"]):
            mean_before = [gp.ymean for gp in sp.gps]
            y, mu, var = sp.evaluate(x)
            sc = max(np.abs(g["mean"]).max(), 1.0)
            assert np.allclose(mu, g["mean"][t], rtol=1e-9, atol=1e-10 * sc), (t, mu, g["mean"][t])
            assert np.allclose(var, g["var"][t], rtol=1e-7, atol=1e-11 * p["variance"][:, 0].max()), (t, var, g["var"][t])
            assert np.allclose(y, g["y"][t], rtol=1e-9, atol=1e-10 * sc), (t, y, g["y"][t])
            f = -np.atleast_1d(func(np.squeeze(y), g["theta"]))
            assert np.allclose(f, g["fvals"][t], rtol=1e-8), (t, f, g["fvals"][t])
            assert all(gp.ymean != mb for gp, mb in zip(sp.gps, mean_before))      # the normaliser mean moves with the path
        assert np.allclose(sp.X_aux, g["X_aux"]) and np.allclose(np.stack([y[:, 0] for y in sp.Y_aux]), g["Y_aux"], rtol=1e-9, atol=1e-9)
        return
    if str(g["acq"]) == "big":
        # value path at sizes where the blocked GPU kernels do block-row updates (several 64 / 128-row blocks, N above the small-batch limit)
        p = make_problem(name, N=int(g["N"]), n=int(g["n"]), n_w=int(g["n_w"]), L=int(g["L"]), H=H)
        om = oracle_model(p)
        mu, var = om.predict(p["Xc"])
        assert np.allclose(mu, g["mean"], rtol=1e-10, atol=1e-12 * np.abs(g["mean"]).max())
        assert np.allclose(var, g["var"], rtol=1e-9, atol=1e-13 * p["variance"][:, 0].max())
        assert np.allclose(om.predict_noiseless(p["Xc"])[1], g["var_noiseless"], rtol=1e-9, atol=1e-13 * p["variance"][:, 0].max())
        marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], p["Xc"], 1)
        a = oacq.aggregate(marg, p["use_full_support"], p["prob"])
        assert np.allclose(a, g["acq_sequential"], rtol=1e-9, atol=1e-11 * np.abs(g["acq_sequential"]).max())
        return
    if str(g["acq"]) == "marginal":
        p = make_problem(name, N=int(g["N"]), n=int(g["n"]), n_w=int(g["n_w"]), L=int(g["L"]), H=H)
        assert np.array_equal(p["Xc"], g["Xc"]) and np.array_equal(p["theta"], g["theta"])
        om = oracle_model(p)
        func, grad = utilities.get(p["utility"], p["m"])
        for l in range(2):
            np.random.seed(1234 + l)
            Z = np.random.normal(size=(50, p["m"]))
            v, dv = oacq.marginal_objective_loop(om, func, grad, p["theta"][l], Z, p["Xc"], H, affine=bool(g["affine"]))
            sc = np.abs(g["fdf_f_%d" % l]).max()
            assert np.allclose(v, g["f_%d" % l], rtol=1e-10, atol=1e-12 * sc)
            assert np.allclose(v, g["fdf_f_%d" % l], rtol=1e-10, atol=1e-12 * sc)
            assert np.allclose(dv, g["fdf_df_%d" % l], rtol=1e-8, atol=1e-10 * np.abs(g["fdf_df_%d" % l]).max())
        return
    if str(g["acq"]) == "uEI_constrained":
        p = make_problem(name, N=int(g["N"]), n=int(g["n"]), L=int(g["L"]), H=H, utility="constrained")
        assert np.array_equal(p["Xc"], g["Xc"]) and np.array_equal(p["theta"], g["theta"])
        om = oracle_model(p)
        func, _ = utilities.get("constrained", p["m"])
        om.set_hyperparameters(0)
        marg = oacq.uei_constrained_marginal_loop(om, func, p["theta"], p["Xc"], H)
        assert np.allclose(oacq.aggregate(marg, False, None), g["acq_value"], rtol=1e-11, atol=1e-13)
        om.set_hyperparameters(0)
        f, d = oacq.uei_constrained_marginal_with_gradient_loop(om, func, p["theta"], p["Xc"], H)
        assert np.allclose(oacq.aggregate(f, False, None), g["acq_grad_f"], rtol=1e-11, atol=1e-13)
        assert np.allclose(oacq.aggregate_gradient(d, False, None), g["acq_grad_df"], rtol=1e-9, atol=1e-12 * np.abs(g["acq_grad_df"]).max())
        return
    p = make_problem(name, N=int(g["N"]), n=int(g["n"]), n_w=int(g["n_w"]), L=int(g["L"]), H=H)
    assert np.array_equal(p["Xc"], g["Xc"]) and np.array_equal(p["Z"], g["Z"]) and np.array_equal(p["theta"], g["theta"])
    om = oracle_model(p)
    full, prob = bool(g["use_full_support"]), (g["prob"] if g["prob"].size else None)
    if str(g["acq"]) == "uEI":
        # posterior layer: the reference's MultiOutputGP -> GPModel -> GP -> PosteriorExact -> Stationary code
        for h in range(H):
            om.set_hyperparameters(h)
            mu, var = om.predict(p["Xc"])
            assert np.allclose(mu, g["mean_h%d" % h], rtol=1e-11, atol=1e-12 * np.abs(mu).max())
            assert np.allclose(om.posterior_mean(p["Xc"]), g["pmean_h%d" % h], rtol=1e-11, atol=1e-12 * np.abs(mu).max())
            vtol = 1e-9 * np.abs(var).max()
            assert np.allclose(var, g["var_h%d" % h], rtol=1e-9, atol=vtol)
            assert np.allclose(om.posterior_variance(p["Xc"]), g["pvar_h%d" % h], rtol=1e-9, atol=vtol)
            assert np.allclose(om.predict_noiseless(p["Xc"])[1], g["var_noiseless_h%d" % h], rtol=1e-9, atol=vtol)
            dm, dv = om.posterior_mean_gradient(p["Xc"]), om.posterior_variance_gradient(p["Xc"])
            assert np.allclose(dm, g["dmean_h%d" % h], rtol=1e-9, atol=1e-11 * np.abs(dm).max())
            assert np.allclose(dv, g["dvar_h%d" % h], rtol=1e-7, atol=1e-8 * np.abs(dv).max())
            F = om.posterior_mean_at_evaluated_points()
            assert np.allclose(F, g["F_h%d" % h], rtol=1e-11, atol=1e-12 * np.abs(F).max())
        om.set_hyperparameters(0)
        marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], p["Xc"], H, incumbent="stale")
        assert np.allclose(oacq.aggregate(marg, full, prob), g["acq_sequential"], rtol=1e-11, atol=1e-13)
        om.set_hyperparameters(0)
        marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], p["Xc"], H, incumbent="per_h")
        assert np.allclose(oacq.aggregate(marg, full, prob), g["acq_parallel"], rtol=1e-11, atol=1e-13)
        om.set_hyperparameters(0)
        f, d = oacq.uei_marginal_with_gradient_vec(om, p["utility"], p["theta"], p["Z"], p["Xc"], H)
        assert np.allclose(oacq.aggregate(f, full, prob), g["acq_grad_f"], rtol=1e-11, atol=1e-13)
        assert np.allclose(oacq.aggregate_gradient(d, full, prob), g["acq_grad_df"], rtol=1e-9, atol=1e-12)
    else:
        marg = oacq.uei_affine_marginal_vec(om, p["theta"], p["Xc"], H)
        assert np.allclose(oacq.aggregate(marg, full, prob), g["acq_value"], rtol=1e-11, atol=1e-13)
        f, d = oacq.uei_affine_marginal_with_gradient_loop(om, p["theta"], p["Xc"], H)
        assert np.allclose(oacq.aggregate(f, full, prob), g["acq_grad_f"], rtol=1e-11, atol=1e-13)
        assert np.allclose(oacq.aggregate_gradient(d, full, prob), g["acq_grad_df"], rtol=1e-9, atol=1e-12)
