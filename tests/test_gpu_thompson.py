"""Utility Thompson sampling on the device (`bopl_path_begin / bopl_path_eval_host`, csrc/path_kernels.cu)
against (i) outputs of the reference's OWN `uTS.suggest_sample` closure (tests/golden/uTS_*.npz, made by
tests/golden/make_golden.py::main_thompson from /root/reference) and (ii) the oracle's re-fit-per-evaluation restatement
(oracle/thompson.py) on longer paths and larger models.

The device grows the Cholesky factor by one row per evaluation; the reference re-fits from scratch.  Same algebra, so the
two paths agree to rounding: means and draws to 1e-9 of their scale, variances to 1e-9 relative plus the floor any fp64
implementation has where var << variance (a revisited point: var ~ noise = 1e-6 * variance, cancellation of 1e6).
"""
import os

import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from oracle import utilities
from oracle.thompson import SamplePathOracle
from tests.helpers import gpu_model, make_utility, oracle_model

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9


def _files():
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith("uTS_") and f.endswith(".npz"))


def _check_step(t, got, want, var_scale, cond_slack=1.0):
    y, mu, var = got
    wy, wmu, wvar = want
    sc = max(float(np.max(np.abs(wmu))), 1.0)
    assert np.all(np.abs(mu - wmu) <= RTOL * sc * cond_slack), (t, mu, wmu)
    assert np.all(np.abs(y - wy) <= RTOL * sc * cond_slack), (t, y, wy)
    # variance: 1e-9 relative + 64 ulp of the kernel variance it is the difference of
    assert np.all(np.abs(var - wvar) <= RTOL * np.abs(wvar) + 64 * 2.0 ** -52 * var_scale * cond_slack), (t, var, wvar)


@pytest.mark.parametrize("fname", _files())
def test_sample_path_against_reference_generated_golden(fname):
    """Same seed -> same np.random stream -> the device path must reproduce the reference's draws point by point."""
    from bopl_b200.sampling_policies import SamplePath
    g = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    p = make_problem(str(g["name"]), N=4, n=int(g["n"]), n_w=2, L=8, H=int(g["H"]))
    model = gpu_model(p)
    func, _ = utilities.get(p["utility"], p["m"])
    path = SamplePath(model, capacity=4)                    # smaller than the path: the buffers must grow (twice)
    np.random.seed(int(g["seed"]))
    vs = float(p["variance"][:, 0].max())
    for t, x in enumerate(g["Xseq"]):
        y = path.evaluate(x)
        _check_step(t, (y, path.last_mean, path.last_var), (g["y"][t], g["mean"][t], g["var"][t]), vs)
        f = -np.atleast_1d(func(np.squeeze(y), g["theta"]))
        assert np.allclose(f, g["fvals"][t], rtol=1e-7), (t, f, g["fvals"][t])
    assert len(path) == len(g["Xseq"])
    assert np.allclose(path.X, g["X_aux"]) and np.allclose(np.stack([yy[:, 0] for yy in path.Y]), g["Y_aux"], rtol=1e-8, atol=1e-9)
    model.engine.path_end()
    assert model.engine.path_length == -1


@pytest.mark.parametrize("name,n,T,kind", [("S", 300, 40, "Mat52"), ("dtlz2a_quad", 129, 60, "rbf"), ("portfolio1", 64, 30, "Mat32")])
def test_sample_path_against_refit_oracle(name, n, T, kind):
    """Longer paths, other kernels, n not a multiple of anything: growth vs the oracle's from-scratch refits."""
    from bopl_b200.sampling_policies import SamplePath
    p = make_problem(name, N=4, n=n, n_w=2, L=4, kind=kind, noise=1e-4)
    model = gpu_model(p)
    om = oracle_model(p)
    sp = SamplePathOracle([om.gps[j][0] for j in range(p["m"])])
    path = SamplePath(model, capacity=16)
    rs = np.random.RandomState(5)
    vs = float(p["variance"][:, 0].max())
    for t in range(T):
        x = rs.uniform(p["lo"], p["hi"], size=p["d"])
        if t % 7 == 6:
            x = path.X[-3] + 1e-3 * rs.standard_normal(p["d"])       # crowd the path: nearly dependent rows
        z = rs.standard_normal(p["m"])
        got = path.evaluate(x, z)
        want = sp.evaluate(x, z)
        _check_step(t, (got, path.last_mean, path.last_var), want, vs, cond_slack=16.0)
    model.engine.path_end()


def test_views_follow_the_reference_call_pattern():
    """`get_copy_of_model_sample()` elements driven exactly as uTS.py:40-47 drives GPy models."""
    p = make_problem("vlmop3_exp", N=4, n=30, n_w=2, L=4, H=2)
    model = gpu_model(p)
    om = oracle_model(p)
    sp = SamplePathOracle([om.gps[j][0] for j in range(p["m"])])
    model_sample = model.get_copy_of_model_sample()
    assert len(model_sample) == p["m"] and np.array_equal(model_sample[0].X, p["X"])
    X_aux = np.copy(model_sample[0].X)
    Y_aux = [np.copy(ms.Y) for ms in model_sample]
    rs = np.random.RandomState(11)
    np.random.seed(77)
    st = np.random.get_state()
    for t in range(6):
        x_new = np.atleast_2d(rs.uniform(p["lo"], p["hi"], size=p["d"]))
        X_aux = np.vstack((X_aux, x_new))
        y_new = []
        for j in range(len(model_sample)):
            y_new.append(model_sample[j].posterior_samples_f(x_new, size=1, full_cov=True)[0])
            Y_aux[j] = np.vstack((Y_aux[j], y_new[j]))
            model_sample[j].set_XY(X_aux, Y_aux[j])
        now = np.random.get_state()
        np.random.set_state(st)
        want = sp.evaluate(x_new[0])[0]
        st = np.random.get_state()
        np.random.set_state(now)
        assert np.allclose(np.squeeze(np.asarray(y_new)), want, rtol=1e-8, atol=1e-9)
    with pytest.raises(RuntimeError):
        model_sample[1].posterior_samples_f(np.atleast_2d(rs.uniform(size=p["d"])))
    model.engine.path_end()


@pytest.mark.parametrize("optimizer", ["lbfgs", "CMA"])
def test_uts_policy_suggests_a_point_in_the_box(optimizer):
    from bopl_b200.anchor_points import BoxSpace
    from bopl_b200.sampling_policies import uTS
    p = make_problem("dtlz1a_linear", N=4, n=40, n_w=2, L=4)
    model = gpu_model(p)
    space = BoxSpace([(p["lo"], p["hi"])] * p["d"])
    np.random.seed(3)
    policy = uTS(model, space, optimizer=optimizer, utility=make_utility(p), maxfevals=120)
    x = policy.suggest_sample()
    assert x.shape == (1, p["d"]) and np.all(x >= p["lo"]) and np.all(x <= p["hi"])
    assert policy.last_path_length >= 10 and policy.X_aux.shape[0] == p["n"] + policy.last_path_length
    assert model.engine.path_length == -1                     # the path is released
    # the model itself is untouched by the path: same posterior as before
    mu0 = gpu_model(p).posterior_mean(p["Xc"])
    assert np.array_equal(model.posterior_mean(p["Xc"]), mu0)


def test_path_errors_are_loud():
    from bopl_b200._lib import BoplError
    from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
    p = make_problem("dtlz1a_linear", N=4, n=20, n_w=2, L=4)
    st = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_linv=False)
    model = MultiOutputGP.from_state(st)
    with pytest.raises(BoplError):                            # no inverse factor
        model.engine.path_begin(0)
    model = gpu_model(p)
    with pytest.raises(BoplError):                            # no path yet
        model.engine.path_eval(p["Xc"][0], np.zeros(p["m"]))
    model.engine.path_begin(0)
    model.engine.path_eval(p["Xc"][0], np.zeros(p["m"]))
    assert model.engine.path_length == 1
    model.engine.set_model(model.state)                       # a new model drops the path
    assert model.engine.path_length == -1


def test_path_on_a_device_fitted_model():
    """The path runs on factors that never existed on the host (bopl_fit_model + want_linv)."""
    from bopl_b200.models import MultiOutputGP
    from bopl_b200.sampling_policies import SamplePath
    p = make_problem("dtlz2a_quad", N=4, n=150, n_w=2, L=4, noise=1e-4)
    gm = MultiOutputGP(p["m"], kernel=list(p["kinds"]), n_samples=p["H"])
    gm.set_hyperparameter_samples(p["variance"], p["lengthscale"], p["noise"], kinds=p["kinds"])
    gm.updateModel(p["X"], list(p["Y"]))
    om = oracle_model(p)
    sp = SamplePathOracle([om.gps[j][0] for j in range(p["m"])])
    path = SamplePath(gm)
    rs = np.random.RandomState(2)
    for t in range(10):
        x, z = rs.uniform(p["lo"], p["hi"], size=p["d"]), rs.standard_normal(p["m"])
        _check_step(t, (path.evaluate(x, z), path.last_mean, path.last_var), sp.evaluate(x, z), float(p["variance"].max()), 16.0)
    gm.engine.path_end()


def test_parego_policy_on_the_accelerated_classes():
    """ParEGO.py:32-66 composed of the GPU-backed MultiOutputGP(1) + uEI_affine + AcquisitionOptimizer(baseline points)."""
    from bopl_b200 import BasicModel, ParEGO
    from bopl_b200.anchor_points import BoxSpace
    from bopl_b200.sampling_policies import chebyshev_scalarization
    p = make_problem("vlmop3_exp", N=4, n=24, n_w=2, L=4)
    space = BoxSpace([(p["lo"], p["hi"])] * p["d"])
    model = BasicModel(output_dim=p["m"])
    model.updateModel(p["X"], [p["Y"][j][:, None] for j in range(p["m"])])
    np.random.seed(5)
    pol = ParEGO(model, space, make_utility(p), aux_model_options=dict(
        n_samples=2, fit_options=dict(n_burnin=10, subsample_interval=2, leapfrog_steps=5, max_iters=50)))
    x = pol.suggest_sample()
    assert x.shape == (1, p["d"]) and np.all(x >= p["lo"]) and np.all(x <= p["hi"])
    w = np.array([0.2, 0.3, 0.5])
    Y = np.arange(12.0).reshape(3, 4)
    assert np.allclose(chebyshev_scalarization(Y, w), np.min(w[:, None] * Y, axis=0) + 0.05 * Y.sum(axis=0))
