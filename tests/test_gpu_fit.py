"""Device-side model fit (`bopl_fit_model`, `bopl_log_marginal`) against LAPACK on the host — the
routines the reference itself calls (GPy/util/linalg.py: dpotrf, dpotrs, dpotri; exact_gaussian_inference.py:44-65).

Two kinds of checks:
  * backward-error checks that hold for ANY correct fp64 factorisation whatever the conditioning: L L^T = Ky, Ky alpha = y,
    Winv Ky = I, Linv L = I, each to a few n * eps of the natural scale;
  * forward comparison with the host/LAPACK results and — what matters to the path — posterior and acquisition parity
    against the oracle with the SAME tolerances as tests/test_gpu_parity.py when the model state was fitted on the device.
"""
import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from oracle import acquisition as oacq
from tests.helpers import gpu_acquisition, oracle_model, rel_err, scaled_err
from tests.test_gpu_parity import CONFIGS, RTOL, check_close, check_dvar, variance_floor

pytestmark = pytest.mark.gpu
EPS = 2.0 ** -52


def device_fit_model(p, gradients=True):
    from bopl_b200.models import MultiOutputGP
    gm = MultiOutputGP(p["m"], kernel=list(p["kinds"]), n_samples=p["H"], gradients=gradients)
    gm.set_hyperparameter_samples(p["variance"], p["lengthscale"], p["noise"], kinds=p["kinds"])
    gm.updateModel(p["X"], list(p["Y"]))
    assert gm.state.on_device
    return gm


def host_fit(p, j, h):
    from bopl_b200.models import _train_cov, fit_exact_gp, inverse_factor
    Ky = _train_cov(p["kinds"][j], p["variance"][j, h], p["lengthscale"][j, h], p["X"])
    Ky[np.diag_indices_from(Ky)] += p["noise"][j, h] + 1e-8
    L, alpha, ymean, Winv = fit_exact_gp(p["X"], p["Y"][j], p["kinds"][j], p["variance"][j, h], p["lengthscale"][j, h], p["noise"][j, h])
    return Ky, L, alpha, ymean, Winv, inverse_factor(L)


@pytest.mark.parametrize("name,n", [("dtlz1a_linear", None), ("dtlz2a_quad", None), ("vlmop3_exp", None), ("portfolio1", None),
                                    ("portfolio1", 1), ("portfolio1", 127), ("portfolio1", 128), ("portfolio1", 129),
                                    ("dtlz2a_quad", 300), ("S", 700)])
def test_device_fit_backward_errors_and_lapack_agreement(name, n):
    p = make_problem(name, n=n, N=8, H=2 if name == "dtlz2a_quad" else 1)
    gm = device_fit_model(p)
    nn = p["n"]
    for j in range(p["m"]):
        for h in range(p["H"]):
            Ky, L_h, alpha_h, ymean, Winv_h, Linv_h = host_fit(p, j, h)
            L, alpha, Winv, Linv = gm.engine.get_factor(j, h, want_winv=True, want_linv=True)
            yc = p["Y"][j] - ymean
            nK = np.linalg.norm(Ky, 2)
            assert np.all(np.triu(L, 1) == 0.0)
            # backward errors
            assert np.linalg.norm(L @ L.T - Ky) <= 8 * nn * EPS * nK, "L L^T"
            assert np.linalg.norm(Ky @ alpha - yc) <= 64 * nn * EPS * (nK * np.linalg.norm(alpha) + np.linalg.norm(yc)), "alpha"
            assert np.linalg.norm(Linv @ L - np.eye(nn)) <= 64 * nn * EPS * np.linalg.norm(Linv, 2) * np.linalg.norm(L, 2), "Linv"
            assert np.linalg.norm(Winv @ Ky - np.eye(nn)) <= 64 * nn * EPS * np.linalg.norm(Winv, 2) * nK, "Winv"
            assert np.array_equal(Winv, Winv.T)
            # forward agreement with LAPACK, scaled by the conditioning both computations are subject to
            cond = nK * np.linalg.norm(Winv_h, 2)
            assert scaled_err(L, L_h) <= 8 * nn * EPS * max(cond ** 0.5, 1.0) + 64 * EPS * cond, ("L", scaled_err(L, L_h), cond)
            assert scaled_err(alpha, alpha_h) <= 64 * nn * EPS * cond, ("alpha", scaled_err(alpha, alpha_h), cond)
            assert scaled_err(Winv, Winv_h) <= 64 * nn * EPS * cond, ("Winv", scaled_err(Winv, Winv_h), cond)
            assert scaled_err(Linv, Linv_h) <= 64 * nn * EPS * cond, ("Linv", scaled_err(Linv, Linv_h), cond)
            # log|K|
            assert abs(gm.state.logdet[j, h] - 2.0 * np.sum(np.log(np.diag(L_h)))) <= 1e-10 * max(1.0, abs(gm.state.logdet[j, h]))


@pytest.mark.parametrize("name", CONFIGS)
def test_posterior_and_acquisition_parity_with_device_fitted_model(name):
    """Same assertions as test_posterior_named_configs / test_uei_value_named_configs / test_uei_gradient, model fitted on the GPU."""
    p = make_problem(name)
    om, gm = oracle_model(p), device_fit_model(p)
    X = p["Xc"]
    mu_o, v_o = om.predict(X)
    mu_g, v_g = gm.predict(X)
    check_close(mu_g, mu_o, name + " mean")
    floor = variance_floor(om, X)
    assert np.all(np.abs(v_g - v_o) <= RTOL * np.abs(v_o) + 64 * floor + 1e-300), \
        "%s var: worst excess %.3e" % (name, np.max(np.abs(v_g - v_o) - RTOL * np.abs(v_o) - 64 * floor))
    check_close(gm.posterior_mean_at_evaluated_points(), om.posterior_mean_at_evaluated_points(), name + " F")
    # small-batch path (explicit inverse factor computed on the device)
    mu_s, v_s = gm.predict(X[:7])
    check_close(mu_s, mu_o[:, :7], name + " small mean")
    assert np.all(np.abs(v_s - v_o[:, :7]) <= RTOL * np.abs(v_o[:, :7]) + 64 * floor[:, :7] + 1e-300)
    # gradients (Winv computed on the device)
    check_close(gm.posterior_mean_gradient(X[:40]), om.posterior_mean_gradient(X[:40]), name + " dmean")
    check_dvar(gm.posterior_variance_gradient(X[:40]), om, X[:40], name + " dvar")
    # acquisition value and argmax
    acq = gpu_acquisition(p, gm)
    got = acq._compute_acq(X)[:, 0]
    marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], X, 1)
    want = oacq.aggregate(marg, p["use_full_support"], p["prob"])[:, 0]
    check_close(got, want, name + " uEI")
    assert int(np.argmax(got)) == int(np.argmax(want))


def test_config_S_device_fit_plain_relative():
    """Benchmark shape (n=2000, m=3, d=10), first 2048 candidates, model fitted on the GPU: plain 1e-9 relative."""
    p = make_problem("S", N=2048)
    om, gm = oracle_model(p), device_fit_model(p, gradients=False)
    mu_o, v_o = om.predict(p["Xc"])
    mu_g, v_g = gm.predict(p["Xc"])
    assert rel_err(v_g, v_o) < RTOL, rel_err(v_g, v_o)
    check_close(mu_g, mu_o, "S mean")


def test_device_fit_is_deterministic_and_batch_invariant():
    """The same instance gives bit-identical factors alone (m=1) or inside a batch of attributes / hyper-samples."""
    p = make_problem("dtlz2a_quad", n=200, N=8, H=3)
    gm = device_fit_model(p)
    a = [gm.engine.get_factor(1, 2, True, True) for _ in range(1)][0]
    gm2 = device_fit_model(p)
    b = gm2.engine.get_factor(1, 2, True, True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    q = dict(p)
    q.update(m=1, H=1, Y=p["Y"][1:2], kinds=p["kinds"][1:2], variance=p["variance"][1:2, 2:3],
             lengthscale=p["lengthscale"][1:2, 2:3], noise=p["noise"][1:2, 2:3])
    gm3 = device_fit_model(q)
    c = gm3.engine.get_factor(0, 0, True, True)
    for x, y in zip(a, c):
        assert np.array_equal(x, y)


def test_not_positive_definite_is_reported_and_jitter_retry_works():
    from bopl_b200._lib import BoplError
    p = make_problem("portfolio1", n=60, N=4)
    p["X"][7:20] = p["X"][3]                    # duplicated inputs and zero noise: exactly singular covariance
    p["kinds"] = ["rbf"] * p["m"]
    p["noise"][:] = 0.0
    p["lengthscale"][:] = 50.0                  # nearly constant kernel: rank-deficient to working precision
    p["variance"][:] = 1e10                     # rounding errors of the Schur complements (1e-6) exceed the reference's 1e-8 jitter
    from bopl_b200.models import MultiOutputGP
    gm = MultiOutputGP(p["m"], kernel=list(p["kinds"]), n_samples=1)
    with pytest.raises(BoplError) as ei:
        gm.engine.fit_model(p["X"], list(p["Y"]), p["kinds"], p["variance"], p["lengthscale"], p["noise"], max_tries=0)
    assert ei.value.code == -5 and "not positive definite" in str(ei.value)
    st = gm.engine.fit_model(p["X"], list(p["Y"]), p["kinds"], p["variance"], p["lengthscale"], p["noise"])   # jitchol-style retry
    assert np.all(st.jitter > 0.0)
    L, alpha, _, _ = gm.engine.get_factor(0, 0)
    assert np.all(np.isfinite(L)) and np.all(np.isfinite(alpha))


def test_jitter_retry_touches_only_the_failing_instance():
    """jitchol jitters the matrix that failed, not its neighbours (GPy/util/linalg.py:53-78 is called per GP): with one singular and
    one healthy attribute the healthy attribute's factor is bit-identical to a fit without any retry; a fit that fails for good leaves
    the engine without a state."""
    from bopl_b200._lib import BoplError
    from bopl_b200.models import MultiOutputGP
    p = make_problem("dtlz2a_quad", n=60, N=4)
    m = p["m"]
    X = p["X"].copy()
    X[7:20] = X[3]
    kinds = ["rbf"] * m
    var, ell, noise = p["variance"].copy(), p["lengthscale"].copy(), np.full((m, 1), 1e-3)
    good = MultiOutputGP(m, kernel=kinds, n_samples=1)
    good.engine.fit_model(X, list(p["Y"]), kinds, var, ell, noise)
    L_good = good.engine.get_factor(0, 0)[0]
    var[1], ell[1], noise[1] = 1e10, 50.0, 0.0           # attribute 1: rank-deficient to working precision
    gm = MultiOutputGP(m, kernel=kinds, n_samples=1)
    st = gm.engine.fit_model(X, list(p["Y"]), kinds, var, ell, noise)
    assert st.jitter[1, 0] > 0.0 and np.all(st.jitter[[0, 2, 3], 0] == 0.0)
    assert np.array_equal(gm.engine.get_factor(0, 0)[0], L_good)
    with pytest.raises(BoplError):
        gm.engine.fit_model(X, list(p["Y"]), kinds, var, ell, noise, max_tries=0)
    assert gm.engine.state is None


@pytest.mark.parametrize("kind,ard_shared", [("Mat52", False), ("rbf", False), ("Mat32", True), ("ExpQuad", False)])
def test_log_marginal_and_gradient_match_host_objective(kind, ard_shared):
    """`bopl_log_marginal` against the host NumPy/LAPACK likelihood of bopl_b200/fit.py (itself FD-checked in the CPU suite)
    and against central finite differences of the device value."""
    from bopl_b200.fit import _kernel_and_grads
    from bopl_b200.models import MultiOutputGP
    from scipy.linalg import lapack
    p = make_problem("dtlz2a_quad", n=150, N=4, kind=kind, noise=1e-3)
    m, d, n = p["m"], p["d"], p["n"]
    eng = MultiOutputGP(m).engine
    Yc = p["Y"] - p["Y"].mean(axis=1, keepdims=True)
    var, ell, noise = p["variance"][:, 0], p["lengthscale"][:, 0], p["noise"][:, 0]
    if ard_shared:
        ell = np.repeat(ell[:, :1], d, axis=1)
    ll, grad, info = eng.log_marginal(p["X"], Yc, [kind] * m, var, ell, noise)
    assert np.all(info == 0)
    for j in range(m):
        K, dK_dvar, dK_dell = _kernel_and_grads(kind, var[j], ell[j], p["X"])
        Ky = K + (noise[j] + 1e-8) * np.eye(n)
        L = lapack.dpotrf(Ky, lower=1)[0]
        alpha = lapack.dpotrs(L, Yc[j][:, None], lower=1)[0]
        Kinv = lapack.dpotri(L, lower=1)[0]
        Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
        want_ll = -0.5 * float(Yc[j] @ alpha[:, 0]) - np.sum(np.log(np.diag(L))) - 0.5 * n * np.log(2 * np.pi)
        dL_dK = 0.5 * (alpha @ alpha.T - Kinv)
        want = np.array([np.sum(dL_dK * dK_dvar)] + [np.sum(dL_dK * dk) for dk in dK_dell] + [np.trace(dL_dK)])
        assert abs(ll[j] - want_ll) <= 1e-9 * max(1.0, abs(want_ll)), (ll[j], want_ll)
        assert np.all(np.abs(grad[j] - want) <= 1e-7 * np.max(np.abs(want)) + 1e-9), (grad[j], want)
    # finite differences of the device value in log-space of one lengthscale and the variance
    h = 1e-5
    for (arr, col) in ((ell, 1), (var, None)):
        up, dn = arr.copy(), arr.copy()
        if col is None:
            up *= (1 + h); dn *= (1 - h)
            a_u = eng.log_marginal(p["X"], Yc, [kind] * m, up, ell, noise, want_grad=False)[0]
            a_d = eng.log_marginal(p["X"], Yc, [kind] * m, dn, ell, noise, want_grad=False)[0]
            fd = (a_u - a_d) / (2 * h * arr)
            assert np.all(np.abs(fd - grad[:, 0]) <= 1e-4 * np.abs(grad[:, 0]) + 1e-5)
        elif not ard_shared:
            up[:, col] *= (1 + h); dn[:, col] *= (1 - h)
            a_u = eng.log_marginal(p["X"], Yc, [kind] * m, var, up, noise, want_grad=False)[0]
            a_d = eng.log_marginal(p["X"], Yc, [kind] * m, var, dn, noise, want_grad=False)[0]
            fd = (a_u - a_d) / (2 * h * arr[:, col])
            assert np.all(np.abs(fd - grad[:, 1 + col]) <= 1e-4 * np.abs(grad[:, 1 + col]) + 1e-5)


def test_device_hyperparameter_inference_matches_host_inference():
    """MAP + HMC with every likelihood evaluation on the device (lock-step over the attributes) against the host/LAPACK run
    with the same per-attribute random streams: the MAP estimates agree to optimiser tolerance, the short HMC chains land
    on the same hyper-samples (the two likelihoods differ by rounding only)."""
    from bopl_b200.fit import GPObjective, fit_hyperparameter_samples, fit_hyperparameter_samples_device, map_estimate
    from bopl_b200.models import MultiOutputGP
    p = make_problem("dtlz2a_quad", n=60, N=4)
    m = p["m"]
    eng = MultiOutputGP(m).engine
    Ys = [p["Y"][j] for j in range(m)]
    kinds, ARD, ef, nv = ["Mat52"] * m, [True, True, False, True], [True, True, True, False], [None] * m
    kw = dict(n_samples=3, n_burnin=4, subsample_interval=2, leapfrog_steps=5)
    v, l, s = fit_hyperparameter_samples_device(eng, p["X"], Ys, kinds, ARD, ef, nv,
                                                rngs=[np.random.RandomState(7 + j) for j in range(m)], **kw)
    for j in range(m):
        v1, l1, s1 = fit_hyperparameter_samples(p["X"], Ys[j], kinds[j], ARD[j], ef[j], nv[j], rng=np.random.RandomState(7 + j), **kw)
        assert np.allclose(v[j], v1, rtol=1e-4), (j, v[j], v1)
        assert np.allclose(l[j], l1, rtol=1e-4), (j, l[j], l1)
        assert np.allclose(s[j], s1, rtol=1e-4), (j, s[j], s1)


def test_update_model_infers_hyperparameters_and_fits_on_device():
    """MultiOutputGP.updateModel(X, Y) with no hyper-parameters supplied: MAP (fixed_hyps) on the device, then the device fit;
    the posterior interpolates the (noise-free) training targets."""
    from bopl_b200.models import MultiOutputGP
    p = make_problem("dtlz1a_linear", n=40, N=4)
    gm = MultiOutputGP(p["m"], exact_feval=[True] * p["m"], fixed_hyps=True, fit_options=dict(max_iters=60))
    gm.updateModel(p["X"], [p["Y"][j][:, None] for j in range(p["m"])])
    assert gm.state.on_device and gm.number_of_hyps_samples() == 1
    F = gm.posterior_mean_at_evaluated_points()
    assert np.max(np.abs(F - p["Y"])) <= 1e-3 * np.max(np.abs(p["Y"]))
    mu, var = gm.predict(p["Xc"])
    assert mu.shape == (p["m"], 4) and np.all(var > 0)
