"""CPU tests of the host-side logic: C-ABI library exports, utility objects, anchor selection, sharding helpers,
state extraction.  No compute call touches a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

from bopl_b200 import _lib
from bopl_b200.anchor_points import (BoxSpace, ObjectiveAnchorPointsGenerator, samples_multidimensional_uniform)
from bopl_b200.sharding import merge_topk, pack_records, shard_bounds
from bopl_b200.synthetic import make_problem
from bopl_b200.utility import Utility, UtilityDistribution, infer_kind, preference_encoder
from oracle import utilities

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    _lib.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "bopl_b200.h")).read()
    declared = set(re.findall(r"\b(bopl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bopl_build_info()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.bopl_create(ctypes.byref(h), 0)
    assert rc != 0 and not h
    assert b"no CUDA device" in lib.bopl_last_error(None)
    from bopl_b200.engine import Engine
    with pytest.raises(_lib.BoplError):
        Engine(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bopl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


@pytest.mark.parametrize("kind,m,theta", [("linear", 3, [0.2, 0.3, 0.5]), ("quadratic", 4, [0.1, -0.4, 0.3, 1.0]),
                                           ("exponential", 3, [0.25]), ("constrained", 2, [-0.1]),
                                           ("constrained", 4, [-0.2, 0.1, -0.5])])
def test_utility_kind_inference(kind, m, theta):
    func, grad = utilities.get(kind, m)
    assert infer_kind(func, m, np.array(theta)) == kind
    u = Utility(func=func, gradient=grad, parameter_distribution=None)
    assert u.resolve_kind(m, np.array(theta)) == kind
    y = np.random.RandomState(0).normal(size=m)
    assert np.array_equal(np.asarray(u.eval_func(y, np.array(theta))), np.asarray(func(y, np.array(theta))))


def test_utility_kind_inference_rejects_unknown_callable():
    with pytest.raises(NotImplementedError):
        infer_kind(lambda y, t: np.tanh(np.dot(t, y)), 3, np.ones(3))
    with pytest.raises(ValueError):
        Utility(func=None, kind="cubic")


def test_utility_distribution_full_support_rule_and_sampling():
    sup = np.arange(24.0).reshape(8, 3)
    d = UtilityDistribution(support=sup, prob_dist=np.ones(8) / 8)
    assert d.use_full_support                                    # < 20 support points (utility_distribution.py:27-30)
    d2 = UtilityDistribution(support=np.zeros((25, 3)), prob_dist=np.ones(25) / 25)
    assert not d2.use_full_support
    np.random.seed(3)
    s = d.sample(5)
    assert s.shape == (5, 3) and all(any(np.array_equal(r, q) for q in sup) for r in s)
    a = d.sample_from_prior(4, seed=7)
    b = d.sample_from_prior(4, seed=7)
    assert np.array_equal(a, b)
    with pytest.raises(Exception):
        UtilityDistribution()


def test_utility_distribution_preference_updates():
    func, _ = utilities.get("linear", 2)
    sup = np.array([[1.0, 0.0], [0.0, 1.0], [0.5, 0.5]])
    pair = [np.array([1.0, 0.0]), np.array([0.0, 1.0])]
    d = UtilityDistribution(support=sup.copy(), prob_dist=np.ones(3) / 3, utility_func=func,
                            elicitation_strategy=lambda Y: [pair[0], pair[1]])
    d.add_preference_information(lambda y: y[0], None, 1)        # the true utility prefers the first item
    assert np.array_equal(d.support, sup[:1]) and np.allclose(d.prob_dist, [1.0])
    assert d.preference_information[0][2] == 1
    # prior sampler + rejection
    rs = np.random.RandomState(0)
    gen = lambda n, seed=None: rs.dirichlet(np.ones(2), n)
    d3 = UtilityDistribution(prior_sample_generator=gen, utility_func=func, elicitation_strategy=lambda Y: [pair[0], pair[1]])
    d3.add_preference_information(lambda y: y[0], None, 1)
    th = d3.sample(6)
    assert th.shape == (6, 2) and np.all(th[:, 0] > th[:, 1])
    assert preference_encoder(1.0, 1.0) == 0 and preference_encoder(0.0, 1.0) == -1


def test_candidate_generator_matches_reference_draw():
    bounds = [(-3.0, 3.0), (0.0, 1.0)]
    a = samples_multidimensional_uniform(bounds, 7, seed=11)
    z = np.random.RandomState(11).uniform(size=(7, 2))
    assert np.array_equal(a, np.stack([6 * z[:, 0] - 3, z[:, 1]], axis=1))
    np.random.seed(5)
    b = samples_multidimensional_uniform(bounds, 3)
    np.random.seed(5)
    assert np.array_equal(b[:, 1], np.random.uniform(size=(3, 2))[:, 1])


def test_anchor_generator_host_path_is_stable_argsort():
    space = BoxSpace([(0, 1)] * 2)
    X = np.random.RandomState(0).uniform(size=(50, 2))
    scores = np.round(X[:, 0], 1)                                # many exact ties
    gen = ObjectiveAnchorPointsGenerator(space, "random", lambda Z: np.round(Z[:, :1], 1), 50)
    anchors, sc = gen.select(X, 5, get_scores=True)
    order = np.argsort(scores, kind="stable")[:5]
    assert np.array_equal(anchors, X[order]) and np.array_equal(sc, scores[order])
    np.random.seed(1)
    out = gen.get(num_anchor=60)
    assert out.shape == (50, 2)                                  # min(len(scores), num_anchor)


def test_shard_bounds_cover_and_balance():
    for N, G in [(10, 3), (2 ** 20, 8), (5, 8), (0, 2)]:
        blocks = [shard_bounds(N, G, r) for r in range(G)]
        assert blocks[0][0] == 0 and blocks[-1][1] == N
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(G - 1))
        sizes = [e - s for s, e in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_merge_topk_is_deterministic_value_then_index():
    vals = np.array([[3.0, 1.0, 0.0], [3.0, 2.0, np.nan], [0.0, 0.0, 0.0]])
    idxs = np.array([[10, 11, 12], [5, 6, 7], [1, 2, 3]])
    v, i, pos = merge_topk(vals, idxs, 6)
    assert i.tolist() == [5, 10, 6, 11, 1, 2] and v.tolist() == [3.0, 3.0, 2.0, 1.0, 0.0, 0.0]
    rec = pack_records(np.array([1.5]), np.array([2 ** 40 + 3]), np.array([[0.25, 0.5]]), 3)
    assert rec.shape == (3, 4) and int(rec[0, 1]) == 2 ** 40 + 3 and np.isnan(rec[1:, 0]).all()


def test_state_extractor_on_reference_shaped_objects():
    """`state_from_reference_model` reads the reference's live-object fields; here they are plain
    namespaces filled by the product's own host fit, and the round trip must be the identity."""
    from types import SimpleNamespace as NS
    from bopl_b200.models import state_from_hyperparameters, state_from_reference_model
    p = make_problem("dtlz1a_linear", n=12, H=2)
    st = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"])
    outs = []
    for j in range(st.m):
        insts = [NS(X=st.X, kern=NS(name=st.kinds[j, h], variance=np.array([st.variance[j, h]]), lengthscale=st.lengthscale[j, h]),
                    likelihood=NS(variance=np.array([st.noise[j, h]])), normalizer=NS(mean=np.array([st.ymean[j, h]])),
                    posterior=NS(_woodbury_chol=np.asfortranarray(st.L[j, h]), woodbury_vector=st.alpha[j, h][:, None],
                                 woodbury_inv=st.Winv[j, h])) for h in range(st.H)]
        outs.append(NS(model_instances=insts))
    st2 = state_from_reference_model(NS(output_dim=st.m, output=outs))
    for f in ("X", "lengthscale", "variance", "noise", "ymean", "L", "alpha", "Winv", "Linv"):
        assert np.array_equal(getattr(st, f), getattr(st2, f)), f
    assert (st.kinds == st2.kinds).all()


def test_host_fit_matches_oracle_fit():
    """The product's LAPACK fit (models.fit_exact_gp) and the oracle's are independent restatements of
    exact_gaussian_inference.py:44-51; they must agree to rounding."""
    from bopl_b200.models import fit_exact_gp
    from oracle.gp import ExactGP
    p = make_problem("vlmop3_exp", n=40)
    for j in range(p["m"]):
        L, alpha, ymean, Winv = fit_exact_gp(p["X"], p["Y"][j], "Mat52", p["variance"][j, 0], p["lengthscale"][j, 0], p["noise"][j, 0])
        g = ExactGP(p["X"], p["Y"][j], "Mat52", p["variance"][j, 0], p["lengthscale"][j, 0], p["noise"][j, 0])
        assert np.allclose(L, np.tril(g.L), rtol=1e-12, atol=1e-14)
        assert np.allclose(alpha, g.alpha[:, 0], rtol=1e-9) and ymean == g.ymean
        assert np.allclose(Winv, g.Winv, rtol=1e-9, atol=1e-9 * np.abs(g.Winv).max())


def test_lockstep_lbfgs_equals_serial_scipy():
    """Every anchor's L-BFGS-B trajectory in the batched lock-step driver is the serial one (same f_df values in,
    same iterates out), while the number of f_df round trips drops to the longest single run."""
    from bopl_b200.optimization import OptLbfgs, lockstep_lbfgs
    rs = np.random.RandomState(0)
    A = rs.normal(size=(4, 4))
    A = A @ A.T + np.eye(4)
    calls = {"batch": 0, "points": 0}

    def f_df_batch(X):
        X = np.atleast_2d(X)
        calls["batch"] += 1
        calls["points"] += len(X)
        fs, gs = [], []
        for x in X:                       # row by row: a point's value must not depend on the batch it travels in
            y = x - 0.3
            Ay = A @ y
            fs.append(0.5 * float(y @ Ay) + float(np.sum(np.cos(3 * x))))
            gs.append(Ay - 3 * np.sin(3 * x))
        return np.array(fs)[:, None], np.array(gs)

    bounds = [(0.0, 1.0)] * 4
    anchors = rs.uniform(size=(7, 4))
    serial = [OptLbfgs(bounds).optimize(a, f_df=lambda x: tuple(v[0] for v in f_df_batch(x))) for a in anchors]
    n_serial_calls = calls["points"]
    calls.update(batch=0, points=0)
    batched, nb = lockstep_lbfgs(list(anchors), f_df_batch, bounds)
    for s, b in zip(serial, batched):
        assert np.array_equal(s[0], b[0]) and np.array_equal(s[1], b[1])
    assert calls["points"] == n_serial_calls and nb == calls["batch"] < n_serial_calls


def test_acquisition_optimizer_host_path_returns_best_local_optimum():
    from bopl_b200.optimization import AcquisitionOptimizer
    f = lambda X: np.sum((np.atleast_2d(X) - 0.7) ** 2, axis=1, keepdims=True)
    f_df = lambda X: (f(X), 2 * (np.atleast_2d(X) - 0.7))
    np.random.seed(0)
    for batched in (True, False):
        opt = AcquisitionOptimizer(BoxSpace([(0, 1)] * 3), n_starting=50, n_anchor=5, batched=batched)
        x, fx = opt.optimize(f=f, f_df=f_df)
        assert x.shape == (1, 3) and np.allclose(x, 0.7, atol=1e-5) and fx < 1e-9
    with pytest.raises(ValueError):
        AcquisitionOptimizer(BoxSpace([(0, 1)] * 3), include_baseline_points=True)     # needs model + utility


def test_latin_design_is_a_centred_latin_hypercube():
    from bopl_b200.anchor_points import latin_design
    np.random.seed(4)
    X = latin_design([(0.0, 1.0), (-2.0, 2.0)], 10)
    assert X.shape == (10, 2)
    assert np.allclose(np.sort(X[:, 0]), (np.arange(10) + 0.5) / 10)
    assert np.allclose(np.sort(X[:, 1]), -2 + 4 * (np.arange(10) + 0.5) / 10)


def test_hyperparameter_objective_gradient_and_map_fit():
    """bopl_b200/fit.py (host-side MAP + HMC, gpmodel.py:109-149): analytic gradient == finite differences; the MAP fit
    recovers the anisotropy of a synthetic GP draw; HMC returns n_samples finite, distinct hyper-samples."""
    from bopl_b200.fit import GPObjective, fit_hyperparameter_samples, map_estimate
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(40, 3))
    Y = np.sin(6 * X[:, 0]) + 0.05 * X[:, 2] + 0.01 * rs.normal(size=40)      # input 1 is irrelevant, input 0 matters most
    for kind, ard, nf in [("Mat52", True, None), ("rbf", True, 1e-2), ("Mat32", False, 1e-3)]:
        obj = GPObjective(X, Y, kind, ard, nf)
        x = obj.initial() + 0.3 * rs.normal(size=obj.size)
        f, g = obj.value_and_grad(x)
        for i in range(obj.size):
            e = np.zeros(obj.size)
            e[i] = 1e-6
            fd = (obj.value_and_grad(x + e)[0] - obj.value_and_grad(x - e)[0]) / 2e-6
            assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(fd)), (kind, i, fd, g[i])
    obj = GPObjective(X, Y, "Mat52", True, None)
    var, ell, noise = obj.unpack(map_estimate(obj))
    assert ell[0] < ell[1] and ell[0] < ell[2] and noise < 0.05
    v, l, s = fit_hyperparameter_samples(X, Y, n_samples=4, n_burnin=10, subsample_interval=2, leapfrog_steps=5, rng=rs)
    assert v.shape == (4,) and l.shape == (4, 3) and s.shape == (4,)
    assert np.all(np.isfinite(l)) and np.all(l > 0) and np.all(v > 0) and len(np.unique(np.round(v, 12))) > 1


def test_hyperparameter_objective_includes_the_logexp_jacobian_of_the_reference():
    """GPy adds, for every priored parameter behind a transformation, the transform's log-Jacobian to the log prior and its gradient
    (core/parameterization/priorizable.py:57-65,76-81; paramz 0.9.5 Logexp: log_jacobian = log(exp(theta) - 1) - theta, gradient
    1 / (exp(theta) - 1)).  The objective is -(log p(y) + sum log Gamma(theta) + sum log_jacobian(theta)), written out here
    independently; its gradient is checked by central differences at small theta (where the Jacobian term dominates)."""
    from scipy.stats import gamma
    from bopl_b200.fit import GPObjective, _softplus
    rs = np.random.RandomState(3)
    X = rs.uniform(size=(25, 2))
    Y = np.cos(4 * X[:, 0]) + 0.02 * rs.normal(size=25)
    obj = GPObjective(X, Y, "Mat52", True, None)                  # variance, 2 lengthscales, free noise
    for x in (obj.initial(), np.array([0.3, -1.0, 0.5, -4.0]), np.array([-2.0, 1.5, -0.5, -6.0])):
        th = _softplus(x)
        ll, _ = obj.host_likelihood(th[0], th[1:3], th[3])
        want = -(ll + gamma.logpdf(th, a=1.0, scale=2.0).sum() + (np.log(np.exp(th) - 1.0) - th).sum())     # Gamma.from_EV(2, 4): a = 1, b = 0.5
        f, g = obj.value_and_grad(x)
        assert abs(f - want) <= 1e-9 * max(1.0, abs(want)), (f, want)
        for i in range(obj.size):
            e = np.zeros(obj.size)
            e[i] = 1e-6
            fd = (obj.value_and_grad(x + e)[0] - obj.value_and_grad(x - e)[0]) / 2e-6
            assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(fd)), (i, fd, g[i])
    # the Jacobian term is what keeps a free noise from drifting to zero: d/dx of -(lp + lj) is negative as theta -> 0+
    x = np.array([0.5, 0.5, 0.5, -12.0])
    th = _softplus(x)
    assert 1.0 / np.expm1(th[3]) > 1e4


def test_lockstep_hyperparameter_inference_equals_per_attribute_runs():
    """fit.py's lock-step drivers (one batched likelihood call per step for all attributes) reproduce the sequential
    per-attribute MAP + HMC runs when every attribute draws from its own RandomState.  The batched likelihood is stood in
    for by the host NumPy/LAPACK one (the device one is compared with it in tests/test_gpu_fit.py)."""
    from bopl_b200.fit import GPObjective, fit_hyperparameter_samples, fit_hyperparameter_samples_device

    class HostStandIn(object):
        calls = 0

        def log_marginal(self, X, Yc, kinds, var, ell, noise, want_grad=True):
            G, d = len(var), X.shape[1]
            ll, grad, info = np.zeros(G), np.zeros((G, d + 2)), np.zeros(G, dtype=np.int32)
            for g in range(G):
                l, gr = GPObjective(X, Yc[g], kinds[g], True, None).host_likelihood(var[g], ell[g], noise[g])
                if l is None:
                    info[g] = 1
                else:
                    ll[g], grad[g] = l, gr
            HostStandIn.calls += 1
            return ll, grad, info

    rs = np.random.RandomState(0)
    X = rs.uniform(size=(40, 3))
    Ys = [np.sin(6 * X[:, 0]) + 0.01 * rs.normal(size=40), np.cos(3 * X[:, 1]) + 0.2 * X[:, 2], X[:, 0] * X[:, 1]]
    kinds, ARD, ef, nv = ["Mat52", "rbf", "Mat32"], [True, True, False], [False, True, False], [None, None, 1e-3]
    kw = dict(n_samples=3, n_burnin=6, subsample_interval=2, leapfrog_steps=4)
    v, l, s = fit_hyperparameter_samples_device(HostStandIn(), X, Ys, kinds, ARD, ef, nv,
                                                rngs=[np.random.RandomState(10 + j) for j in range(3)], **kw)
    assert v.shape == (3, 3) and l.shape == (3, 3, 3) and s.shape == (3, 3)
    for j in range(3):
        v1, l1, s1 = fit_hyperparameter_samples(X, Ys[j], kinds[j], ARD[j], ef[j], nv[j], rng=np.random.RandomState(10 + j), **kw)
        assert np.allclose(v[j], v1, rtol=1e-7) and np.allclose(l[j], l1, rtol=1e-7) and np.allclose(s[j], s1, rtol=1e-7)
    # one batched call per leap-frog step, not one per attribute: (6 + 3*2) HMC iterations x 4 steps + 1, plus the MAP steps
    assert HostStandIn.calls < 12 * 4 + 1 + 200


def test_cma_es_restatement_finds_box_constrained_optima():
    """`OptCma` (GPyOpt/optimization/optimizer.py:229-261) without pycma: interior optimum and optimum on the boundary."""
    from bopl_b200.optimization import OptCma, choose_optimizer, cma_es_minimize
    rs = np.random.RandomState(0)
    x, fx = cma_es_minimize(lambda x: float(np.sum((x - 0.3) ** 2)), np.full(5, 0.9), 0.6, np.zeros(5), np.ones(5), maxfevals=800, rng=rs)
    assert np.allclose(x, 0.3, atol=1e-3) and fx < 1e-5
    x, fx = cma_es_minimize(lambda x: float(np.sum((x + 0.3) ** 2)), np.full(3, 0.9), 0.6, np.zeros(3), np.ones(3), maxfevals=600, rng=rs)
    assert np.allclose(x, 0.0, atol=1e-6)
    calls = []

    def f(X):
        calls.append(np.array(X))
        return np.atleast_1d(np.sum((np.atleast_2d(X) - 0.5) ** 2))
    np.random.seed(1)
    opt = choose_optimizer('CMA', [(0.0, 1.0)] * 4)
    assert isinstance(opt, OptCma)
    x, fx = opt.optimize(np.full((1, 4), 0.1), f=f, maxfevals=200)
    assert x.shape == (1, 4) and len(calls) <= 201 and all(c.shape == (1, 4) for c in calls)      # budget respected, 2-D calls
    assert all(np.all(c >= 0) and np.all(c <= 1) for c in calls)                                  # evaluated inside the box only
    with pytest.raises(IndexError):
        OptCma([(0.0, 1.0)]).optimize(np.zeros(1), f=f)
    with pytest.raises(NotImplementedError):
        choose_optimizer('DIRECT', [(0.0, 1.0)] * 2)


def test_sampling_policies_surface():
    """Random / AcquisitionFunction policies (sampling_policies/Random.py, AcquisitionFunction.py) on host-only stand-ins."""
    from bopl_b200.anchor_points import BoxSpace
    from bopl_b200.sampling_policies import AcquisitionFunction, Random, SamplingPolicyBase, uTS
    space = BoxSpace([(0.0, 2.0)] * 3)
    x = Random(None, space).suggest_sample()
    assert x.shape == (1, 3) and np.all(x >= 0) and np.all(x <= 2)
    with pytest.raises(NotImplementedError):
        SamplingPolicyBase().suggest_sample()

    class Acq(object):
        def __init__(self):
            self.optimizer = type("O", (), {})()
            self.updated = 0

        def update_samples(self):
            self.updated += 1

        def optimize(self, duplicate_manager=None):
            return np.array([[1.0, 1.0, 1.0]]), np.array([[0.5]])
    acq = Acq()
    pol = AcquisitionFunction(None, space, acq)
    assert np.array_equal(pol.suggest_sample(), [[1.0, 1.0, 1.0]]) and acq.updated == 1
    assert acq.optimizer.context_manager.noncontext_bounds == space.get_bounds()
    assert uTS.analytical_gradient_prediction and AcquisitionFunction.analytical_gradient_prediction


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys, the same
    `config.workload` string the GPU arm prints, both CPU figures (vectorised port + the reference's loop form), no GPU needed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "150",
                          "--cpu-sample", "256", "--loop-sample", "8"], stdout=subprocess.PIPE, universal_newlines=True, check=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in b, key
    assert b["impl"] == "reference" and b["value"] > 0 and b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1
    assert b["cpu_baseline"]["loop_form"]["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    class A: n, candidates, n_w, L = 150, 2 ** 20, 64, 32
    assert b["config"]["workload"] == bench.workload_string(bench.workload(A), "linear")


def test_c_abi_header_is_plain_c99(tmp_path):
    """include/bopl_b200.h is the drop-in boundary: it must compile as C (no C++ or torch types in any signature)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "bopl_b200.h"\nint main(void) { bopl_handle* h = 0; return bopl_create(&h, 0) == 0 ? 0 : 1; }\n')
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                          str(tmp_path / "hdr.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    assert res.returncode == 0, res.stdout
