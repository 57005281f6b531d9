"""GPU parity tests of the INT8-tensor-pipe posterior kernel (`posterior_kernel='int8'` / `'int8x7'`,
bopl_b200/csrc/posterior_ozaki.cu): the same path as tests/test_gpu_parity.py — GPy posterior.py:299-320, gp.py:380-435 —
with the n^2/2 update of the triangular solve carried out as exact int8 digit products on tcgen05 / TMEM.

Tolerances.  The split keeps 47 (6 digits) or 55 (7 digits) fractional bits of every operand against a power-of-two scale
(row maximum of L, prior standard deviation for V), so its error is ABSOLUTE in units of the prior variance:
  * 7 digits: the same tolerance as the fp64 kernel (1e-9 relative + 64 x the oracle's own one-ulp-of-K* sensitivity);
  * 6 digits: that, plus 1e-10 * prior variance (measured: 2e-11 at the benchmark shape n = 2000, 6e-13 at n = 300).
Mean, acquisition values, argmax: as for the fp64 kernel (1e-9 of scale, identical argmax)."""
import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from oracle import acquisition as oacq
from tests.helpers import gpu_acquisition, gpu_model, oracle_model, rel_err, scaled_err
from tests.test_gpu_parity import CONFIGS, RTOL, check_close, variance_floor

pytestmark = pytest.mark.gpu

KERNELS = ["int8", "int8x7"]
ABS6 = 1e-10          # 6-digit split: absolute term, in units of the prior variance


def check_var(v_g, v_o, om, p, X, kernel, what):
    floor = variance_floor(om, X)
    tol = RTOL * np.abs(v_o) + 64 * floor + 1e-300
    if kernel == "int8":
        tol = tol + ABS6 * p["variance"][:, om.h if hasattr(om, "h") else 0][:, None]
    err = np.abs(v_g - v_o)
    assert np.all(err <= tol), "%s (%s): worst err/tol %.3e, max err %.3e" % (what, kernel, np.max(err / tol), err.max())


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", CONFIGS)
def test_int8_posterior_named_configs(name, kernel):
    p = make_problem(name, N=700)
    om, gm = oracle_model(p), gpu_model(p, posterior_kernel=kernel)
    assert gm.engine.posterior_kernel == kernel
    mu_o, v_o = om.predict(p["Xc"])
    mu_g, v_g = gm.predict(p["Xc"])
    check_close(mu_g, mu_o, name + " mean")
    check_var(v_g, v_o, om, p, p["Xc"], kernel, name + " var")
    _, vn_o = om.predict_noiseless(p["Xc"])
    _, vn_g = gm.predict_noiseless(p["Xc"])
    check_var(vn_g, vn_o, om, p, p["Xc"], kernel, name + " noiseless var")


@pytest.mark.parametrize("kernel", KERNELS)
def test_int8_posterior_config_S_plain_relative(kernel):
    """Benchmark shape (n=2000, m=3, d=10), first 4096 candidates: plain 1e-9 relative on mean AND variance, like the fp64 kernel;
    the int8 kernels against the fp64 kernel itself: 1e-10 (6 digits) / 1e-12 (7 digits) relative."""
    p = make_problem("S", N=4096)
    om = oracle_model(p)
    gm, g64 = gpu_model(p, gradients=False, posterior_kernel=kernel), gpu_model(p, gradients=False, posterior_kernel="fp64")
    mu_o, v_o = om.predict(p["Xc"])
    mu_g, v_g = gm.predict(p["Xc"])
    mu_d, v_d = g64.predict(p["Xc"])
    assert rel_err(v_g, v_o) < RTOL, rel_err(v_g, v_o)
    check_close(mu_g, mu_o, "S mean")
    assert scaled_err(mu_g, mu_o) < 1e-11
    assert rel_err(v_g, v_d) < (1e-10 if kernel == "int8" else 1e-12), rel_err(v_g, v_d)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n,N", [(1, 33), (5, 40), (63, 129), (64, 128), (65, 257), (127, 129), (129, 33), (300, 257), (513, 1000)])
def test_int8_posterior_ragged_shapes(n, N, kernel):
    """Training sets that are not multiples of the 64-row block (front padding), one block row only (no tensor work at all),
    candidate counts that are not multiples of the 128-candidate tile."""
    p = make_problem("portfolio1", n=n, N=N)
    om, gm = oracle_model(p), gpu_model(p, posterior_kernel=kernel)
    mu_o, v_o = om.predict(p["Xc"])
    mu_g, v_g = gm.predict(p["Xc"])
    check_close(mu_g, mu_o, "mean n=%d N=%d" % (n, N))
    check_var(v_g, v_o, om, p, p["Xc"], kernel, "var n=%d N=%d" % (n, N))


@pytest.mark.parametrize("kind", ["Mat52", "rbf", "Mat32", "ExpQuad"])
def test_int8_posterior_kernel_kinds(kind):
    p = make_problem("dtlz2a_quad", kind=kind, n=150, N=77)
    om, gm = oracle_model(p), gpu_model(p, posterior_kernel="int8")
    mu_o, v_o = om.predict(p["Xc"])
    mu_g, v_g = gm.predict(p["Xc"])
    check_close(mu_g, mu_o, kind + " mean")
    check_var(v_g, v_o, om, p, p["Xc"], "int8", kind + " var")


def test_int8_posterior_hyper_samples_select():
    p = make_problem("dtlz1a_linear", H=3, n=200, N=300)
    om, gm = oracle_model(p), gpu_model(p, posterior_kernel="int8")
    for h in (2, 0, 1):
        om.set_hyperparameters(h)
        gm.set_hyperparameters(h)
        mu_o, v_o = om.predict(p["Xc"])
        mu_g, v_g = gm.predict(p["Xc"])
        check_close(mu_g, mu_o, "mean h=%d" % h)
        floor = variance_floor(om, p["Xc"])
        assert np.all(np.abs(v_g - v_o) <= RTOL * np.abs(v_o) + 64 * floor + ABS6 * p["variance"][:, h][:, None])


def test_int8_posterior_is_batch_and_permutation_invariant():
    """Integer accumulation is exact and every other stage is per candidate: a candidate's numbers do not depend on its batch."""
    p = make_problem("S", n=500, N=3000)
    gm = gpu_model(p, gradients=False, posterior_kernel="int8")
    mu, v = gm.predict(p["Xc"])
    sl = slice(777, 777 + 1301)
    mu_s, v_s = gm.predict(p["Xc"][sl])
    assert np.array_equal(mu_s, mu[:, sl]) and np.array_equal(v_s, v[:, sl])
    perm = np.random.RandomState(4).permutation(3000)
    mu_p, v_p = gm.predict(p["Xc"][perm])
    assert np.array_equal(mu_p, mu[:, perm]) and np.array_equal(v_p, v[:, perm])


@pytest.mark.parametrize("kernel", KERNELS)
def test_int8_posterior_of_a_model_fitted_on_the_device(kernel):
    """bopl_fit_model builds the digit images on the device (oz_scale / oz_digits / oz_stage kernels): same numbers as the images
    the host builds from an uploaded factor, up to the difference of the two Cholesky factorisations."""
    from bopl_b200.models import MultiOutputGP
    p = make_problem("S", n=700, N=900)
    kinds, var, ell, noise = list(p["kinds"]), p["variance"], p["lengthscale"], p["noise"]
    models = {}
    for k in ("fp64", kernel):
        gm = MultiOutputGP(p["m"], kernel=kinds, n_samples=1, device=0, gradients=False, posterior_kernel=k)
        gm.set_hyperparameter_samples(var, ell, noise, kinds=kinds)
        gm.updateModel(p["X"], list(p["Y"]))
        models[k] = gm
    mu_d, v_d = models["fp64"].predict(p["Xc"])
    mu_g, v_g = models[kernel].predict(p["Xc"])
    assert scaled_err(mu_g, mu_d) < 1e-12
    assert rel_err(v_g, v_d) < (1e-10 if kernel == "int8" else 1e-12), rel_err(v_g, v_d)
    om = oracle_model(p)
    mu_o, v_o = om.predict(p["Xc"])
    check_close(mu_g, mu_o, "device-fit mean")
    assert rel_err(v_g, v_o) < RTOL


def test_int8_uei_values_and_argmax():
    """The whole acquisition through the int8 posterior kernel: 1e-9 of scale against the oracle, identical argmax and anchors as the
    fp64 kernel's."""
    p = make_problem("S", n=600, N=5000)
    om = oracle_model(p)
    gm, g64 = gpu_model(p, gradients=False, posterior_kernel="int8"), gpu_model(p, gradients=False, posterior_kernel="fp64")
    a8 = gpu_acquisition(p, gm)._compute_acq(p["Xc"])[:, 0]
    a64 = gpu_acquisition(p, g64)._compute_acq(p["Xc"])[:, 0]
    marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], p["Xc"], 1)
    ref = oacq.aggregate(marg, p["use_full_support"], p["prob"])[:, 0]
    check_close(a8, ref, "int8 uEI")
    assert np.argmax(a8) == np.argmax(ref)
    assert np.array_equal(np.argsort(-a8, kind="stable")[:20], np.argsort(-a64, kind="stable")[:20])


@pytest.mark.parametrize("kernel", ["fp64"] + KERNELS)
@pytest.mark.parametrize("fname", ["big_S.npz", "big_dtlz2a_quad.npz"])
def test_blocked_kernels_against_reference_generated_golden(fname, kernel):
    """All three posterior kernels against outputs of the reference's OWN code (tests/golden/make_golden.py big) at sizes where they do
    block-row updates (n = 300 / 200, N = 200 / 150): mean, variance (+ noiseless), uEI values, argmax — not via the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname), allow_pickle=False)
    p = make_problem(str(g["name"]), N=int(g["N"]), n=int(g["n"]), n_w=int(g["n_w"]), L=int(g["L"]), H=int(g["H"]))
    gm = gpu_model(p, gradients=False, posterior_kernel=kernel)
    om = oracle_model(p)                                    # only for the conditioning floor of the tolerance
    floor = variance_floor(om, p["Xc"])
    extra = ABS6 * p["variance"][:, 0][:, None] if kernel == "int8" else 0.0
    mu, var = gm.predict(p["Xc"])
    check_close(mu, g["mean"], fname + " mean")
    assert np.all(np.abs(var - g["var"]) <= RTOL * np.abs(g["var"]) + 64 * floor + extra), np.max(np.abs(var - g["var"]) / np.abs(g["var"]))
    assert np.all(np.abs(gm.predict_noiseless(p["Xc"])[1] - g["var_noiseless"]) <= RTOL * np.abs(g["var"]) + 64 * floor + extra)
    a = gpu_acquisition(p, gm)._compute_acq(p["Xc"])
    check_close(a, g["acq_sequential"], fname + " uEI")
    assert np.argmax(a) == np.argmax(g["acq_sequential"])


@pytest.mark.parametrize("kernel", KERNELS)
def test_int8_cluster_pairs_equal_single_ctas_bit_for_bit(monkeypatch, kernel):
    """Row-paired launch (clusters of two CTAs per candidate tile: even block rows in one, odd ones in the other, the solved panel's
    k-steps multicast into both, partial sums combined through the peer's shared memory) against one CTA per tile: the arithmetic per
    candidate is the same in the same order, so the results are equal bit for bit — several tiles per pair, ragged last tile — and so
    is whatever the default picks (the launch form is chosen by the number of work items)."""
    p = make_problem("S", n=300, N=149 * 128 - 37)          # 149 tiles x 3 attributes: six work items per pair
    res = {}
    for mode in ("2", "0", "1"):                            # always paired / never / by item count (default)
        monkeypatch.setenv("BOPL_OZ_CLUSTER", mode)
        res[mode] = gpu_model(p, gradients=False, posterior_kernel=kernel).predict(p["Xc"])
    for mode in ("0", "1"):
        assert np.array_equal(res["2"][0], res[mode][0]) and np.array_equal(res["2"][1], res[mode][1]), mode
    monkeypatch.delenv("BOPL_OZ_CLUSTER")
    mu_d, v_d = gpu_model(p, gradients=False, posterior_kernel="fp64").predict(p["Xc"])
    assert scaled_err(res["2"][0], mu_d) < 1e-12 and rel_err(res["2"][1], v_d) < (1e-10 if kernel == "int8" else 1e-12)
    # a small batch (fewer work items than SM pairs: paired by default) and a mid-size one (one CTA per item needs fewer rounds)
    gm = gpu_model(p, gradients=False, posterior_kernel=kernel)
    for N in (400, 4096):
        mu, v = gm.predict(p["Xc"][:N])
        assert np.array_equal(mu, res["2"][0][:, :N]) and np.array_equal(v, res["2"][1][:, :N])


def test_int8_posterior_many_block_rows():
    """n = 3000 (47 block rows of 64, front padding 8): more block rows, longer accumulation chains and larger accumulators than
    the benchmark shape; 6- and 7-digit kernels against the fp64 kernel and the oracle."""
    p = make_problem("S", n=3000, N=400)
    om = oracle_model(p)
    mu_o, v_o = om.predict(p["Xc"])
    g64 = gpu_model(p, gradients=False, posterior_kernel="fp64")
    mu_d, v_d = g64.predict(p["Xc"])
    for kernel, tol in (("int8", 2e-10), ("int8x7", 2e-12)):
        gm = gpu_model(p, gradients=False, posterior_kernel=kernel)
        mu_g, v_g = gm.predict(p["Xc"])
        check_close(mu_g, mu_o, "n=3000 mean")
        assert rel_err(v_g, v_o) < RTOL, (kernel, rel_err(v_g, v_o))
        assert rel_err(v_g, v_d) < tol, (kernel, rel_err(v_g, v_d))


def test_int8_kernel_selection_rules():
    from bopl_b200._lib import BoplError
    p = make_problem("portfolio1", n=80, N=50)
    gm = gpu_model(p, posterior_kernel="fp64")          # fp64 model: no digit images
    with pytest.raises(BoplError):
        gm.engine.set_posterior_kernel("int8")
    assert gpu_model(p).engine.posterior_kernel == "int8x7"     # the library default
    g8 = gpu_model(p, posterior_kernel="int8")
    with pytest.raises(BoplError):
        g8.engine.set_posterior_kernel("int8x7")        # images of the other width
    mu8, v8 = g8.predict(p["Xc"])
    g8.engine.set_posterior_kernel("fp64")              # always allowed: the fp64 operands are always there
    mu64, v64 = g8.predict(p["Xc"])
    mu_r, v_r = gm.predict(p["Xc"])
    assert np.array_equal(mu64, mu_r) and np.array_equal(v64, v_r)
    assert scaled_err(v8, v64) < 1e-9
    # wide inputs (d up to 24: three ring stages) run the int8 kernels too — against the fp64 kernel of the same library
    from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
    rs = np.random.RandomState(5)
    for d in (14, 24):
        m, n = 2, 200
        X, Xc = rs.uniform(size=(n, d)), rs.uniform(size=(300, d))
        Y = np.stack([np.sin(X @ rs.normal(size=d)) for _ in range(m)])
        st = state_from_hyperparameters(X, Y, ["Mat52"] * m, np.ones((m, 1)), np.full((m, 1, d), 3.0), np.full((m, 1), 1e-4))
        b = MultiOutputGP.from_state(st, posterior_kernel="fp64")
        mu_b, v_b = b.predict(Xc)
        for kernel, tol in (("int8", 1e-9), ("int8x7", 1e-11)):
            a = MultiOutputGP.from_state(st, posterior_kernel=kernel)
            assert a.engine.posterior_kernel == kernel
            mu_a, v_a = a.predict(Xc)
            assert not np.array_equal(v_a, v_b)              # it is the int8 kernel that ran
            assert scaled_err(mu_a, mu_b) < 1e-12 and rel_err(v_a, v_b) < tol, (d, kernel, rel_err(v_a, v_b))


@pytest.mark.parametrize("kernel", ["int8"])
def test_int8_full_size_config_S(kernel):
    """Config S at BASELINE.json's full size through the reference-facing call with the 6-digit kernel (the 7-digit default runs
    tests/test_gpu_parity.py::test_full_size_config_S_properties_and_subset_parity): finite non-negative scores,
    anchors = stable argsort, bitwise sub-batch reproducibility, and the same top-20 anchors as the fp64 kernel."""
    p = make_problem("S")
    gm = gpu_model(p, gradients=False, posterior_kernel=kernel)
    acq = gpu_acquisition(p, gm)
    val, idx, a = acq.top_candidates(p["Xc"], 20, want_acq=True)
    assert a.shape == (2 ** 20,) and np.all(np.isfinite(a)) and np.all(a >= 0.0)
    want = np.argsort(-a, kind="stable")[:20]
    assert np.array_equal(idx, want) and np.array_equal(val, a[want])
    sl = slice(300000, 300000 + 70001)
    assert np.array_equal(acq._compute_acq(p["Xc"][sl])[:, 0], a[sl])
    g64 = gpu_model(p, gradients=False, posterior_kernel="fp64")
    val64, idx64, a64 = gpu_acquisition(p, g64).top_candidates(p["Xc"], 20, want_acq=True)
    assert np.array_equal(idx, idx64)
    assert scaled_err(a, a64) < 1e-10


@pytest.mark.parametrize("seed", range(8))
def test_int8x7_randomised_shapes_against_fp64_kernel_and_oracle(seed):
    """Random (n, d, m, N, kernel kind, launch form) draws for the default posterior kernel: mean and variance of BOTH full-precision
    kernels against the oracle under the usual bounds, and against each other.  Exercises odd tile counts, several work items per CTA
    pair, 2 ... 20 block rows, wide inputs, badly conditioned models and both launch forms."""
    import os
    from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
    rs = np.random.RandomState(100 + seed)
    n = int(rs.choice([70, 130, 200, 333, 520, 777, 1100, 1290]))
    d = int(rs.choice([1, 3, 7, 12, 13, 24]))
    m = int(rs.choice([1, 2, 5]))
    N = int(rs.choice([33, 129, 500, 1537, 4000, 9001]))
    kind = str(rs.choice(["Mat52", "rbf", "Mat32", "ExpQuad"]))
    mode = str(rs.choice(["0", "1", "2"]))
    X, Xc = rs.uniform(size=(n, d)), rs.uniform(size=(N, d))
    Y = np.stack([np.sin(X @ rs.normal(size=d)) + 0.1 * rs.normal(size=n) for _ in range(m)])
    ell = rs.uniform(0.4, 1.2, size=(m, 1, d)) * np.sqrt(d)
    var = rs.uniform(0.5, 2.0, size=(m, 1))
    st = state_from_hyperparameters(X, Y, [kind] * m, var, ell, np.full((m, 1), 1e-4), want_winv=False)
    os.environ["BOPL_OZ_CLUSTER"] = mode
    try:
        a = MultiOutputGP.from_state(st, posterior_kernel="int8x7")
    finally:
        del os.environ["BOPL_OZ_CLUSTER"]
    b = MultiOutputGP.from_state(st, posterior_kernel="fp64")
    assert a.engine.posterior_kernel == "int8x7"
    mu_a, v_a = a.predict(Xc)
    mu_b, v_b = b.predict(Xc)
    what = "seed %d: n=%d d=%d m=%d N=%d %s mode %s" % (seed, n, d, m, N, kind, mode)
    # both kernels against the oracle under the usual bounds (the draws include badly conditioned models — one input dimension, 1000+
    # points — where two fp64 summation orders of K* alpha already differ by 1e-10 of the scale), and against each other within the same
    p = dict(X=X, Y=Y, kinds=[kind] * m, variance=var, lengthscale=ell, noise=np.full((m, 1), 1e-4), H=1, m=m)
    om = oracle_model(p)
    mu_o, v_o = om.predict(Xc)
    floor = variance_floor(om, Xc)
    bound = RTOL * np.abs(v_o) + 64 * floor + 1e-300
    check_close(mu_a, mu_o, what + " mean (int8x7)")
    check_close(mu_b, mu_o, what + " mean (fp64)")
    assert np.all(np.abs(v_a - v_o) <= bound), (what, float(np.max(np.abs(v_a - v_o) / bound)))
    assert np.all(np.abs(v_b - v_o) <= bound), (what, float(np.max(np.abs(v_b - v_o) / bound)))
    assert np.all(np.abs(v_a - v_b) <= bound), (what, float(np.max(np.abs(v_a - v_b) / bound)))
    print("%s: worst |dv| / bound: int8x7 %.3f, fp64 %.3f, int8x7 vs fp64 %.3f; mean int8x7 vs fp64 %.2e of scale"
          % (what, np.max(np.abs(v_a - v_o) / bound), np.max(np.abs(v_b - v_o) / bound), np.max(np.abs(v_a - v_b) / bound), scaled_err(mu_a, mu_b)))


def test_int8auto_probes_the_model_at_upload():
    """`posterior_kernel='int8auto'` (BOPL_POSTERIOR_INT8_AUTO): the 6-digit split only where an upload-time probe at the training
    points meets 1e-9 relative on the variance against the fp64 kernel; otherwise the 7-digit images are built.  The benchmark model
    (noise 1e-6: variance down to 1e-6 sigma^2 at the training points) must come out with 7 digits, a noisy model keeps 6; both then
    pass the full-precision bound at and near the training points — host upload and device fit."""
    from bopl_b200.models import MultiOutputGP
    from tests.test_gpu_low_variance import near_training_points
    for noise, want in ((1e-6, "int8x7"), (3e-2, "int8")):
        p = make_problem("S", n=700, N=16, noise=noise)
        om = oracle_model(p)
        Xq = near_training_points(p, 60)
        _, v_o = om.predict_noiseless(Xq)
        bound = RTOL * np.abs(v_o) + 64 * variance_floor(om, Xq)
        gm = gpu_model(p, gradients=False, posterior_kernel="int8auto")
        assert gm.engine.posterior_kernel == want, (noise, gm.engine.posterior_kernel)
        _, v_g = gm.predict_noiseless(Xq)
        assert np.all(np.abs(v_g - v_o) <= bound), (noise, float(np.max(np.abs(v_g - v_o) / bound)))
        kinds = list(p["kinds"])
        gd = MultiOutputGP(p["m"], kernel=kinds, n_samples=1, device=0, gradients=False, posterior_kernel="int8auto")
        gd.set_hyperparameter_samples(p["variance"], p["lengthscale"], p["noise"], kinds=kinds)
        gd.updateModel(p["X"], list(p["Y"]))
        assert gd.engine.posterior_kernel == want
        _, v_d = gd.predict_noiseless(Xq)
        assert np.all(np.abs(v_d - v_o) <= 4 * bound), (noise, float(np.max(np.abs(v_d - v_o) / bound)))
        # the next model on the same handle is probed afresh
        gd.updateModel(p["X"], list(p["Y"]))
        assert gd.engine.posterior_kernel == want
