"""Low-variance parity: candidates AT and within 1e-6 ... 1e-1 of training points, where the raw posterior variance
`variance - |L^-1 K*|^2` (posterior.py:308-320) cancels from the prior variance down to the noise level (var / sigma^2 from 1 to 1e-6).

This is where a posterior kernel that carries fewer bits than the reference shows: an ABSOLUTE error e * sigma^2 becomes a RELATIVE error
e * sigma^2 / var.  All three posterior kernels run the same candidates against the oracle (the reference's algorithm, LAPACK dtrtrs):

  * `int8x7` (library default: the solve's update as exact 55-bit digit products on the int8 tensor pipe) and `fp64` (DMMA) are held
    to the SAME bound, with no extra term: |v - v_oracle| <= 1e-9 |v_oracle| + 64 * floor, floor = the change of the oracle's OWN
    variance under a one-ulp perturbation of its K* input + one ulp of the prior variance (tests/test_gpu_parity.py) — the reference's
    own rounding sensitivity, below which two fp64 implementations with different summation orders cannot agree (at var = 1e-6 sigma^2
    one ulp of sigma^2 is already 2e-10 relative);
  * `int8` (6 digits, opt-in reduced precision) is checked against the same bound + its documented absolute term 1e-10 sigma^2, and the
    test records by how much it MISSES the full-precision bound at low variance — that is why it is not the default.

Every run prints (and, when gpurun_out/ exists, saves to gpurun_out/low_variance.json) the largest element-relative error per decade of
var / sigma^2 and the largest error in units of the bound; profiles/r02_low_variance.json is the committed copy."""
import json
import os

import numpy as np
import pytest

from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model, oracle_model
from tests.test_gpu_parity import RTOL, check_close, variance_floor

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DELTAS = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1]
ABS6 = 1e-10


def near_training_points(p, per_delta, seed=7):
    """Candidates X_i + delta * u (u a random unit direction) for `per_delta` random training points per delta."""
    rs = np.random.RandomState(seed)
    span = p["hi"] - p["lo"]
    out = []
    for delta in DELTAS:
        idx = rs.choice(p["n"], size=per_delta, replace=False)
        u = rs.normal(size=(per_delta, p["d"]))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        out.append(p["X"][idx] + delta * span * u)
    return np.concatenate(out)


def per_decade(v_o, err, sigma2):
    """{decade exponent e: (count, max |err| / |v_o|)} with var / sigma^2 in [10^e, 10^(e+1))."""
    ratio = v_o / sigma2
    dec = np.floor(np.log10(np.maximum(ratio, 1e-300))).astype(int)
    rel = np.abs(err) / np.abs(v_o)
    return {int(e): (int(np.sum(dec == e)), float(rel[dec == e].max())) for e in sorted(set(dec.ravel().tolist()))}


def _record(key, payload):
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out_dir):
        return
    path = os.path.join(out_dir, "low_variance.json")
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[key] = payload
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("kernel", ["int8x7", "fp64", "int8"])
# dtlz2a_quad at n = 162 = the 2 (d + 1) initial points + the experiment's 150 iterations (test_dtlz2a_quad.py:24-26,103-110): three 64-row
# block rows; at the default n = 62 the model is one block row and the int8 kernels do no tensor work at all
@pytest.mark.parametrize("name,n,per_delta", [("S", 2000, 300), ("dtlz2a_quad", 162, 100), ("dtlz2a_quad", None, 40)])
def test_low_variance_candidates(name, n, per_delta, kernel):
    p = make_problem(name, n=n, N=16)
    Xq = near_training_points(p, min(per_delta, p["n"]))
    om, gm = oracle_model(p), gpu_model(p, gradients=False, posterior_kernel=kernel)
    assert gm.engine.posterior_kernel == kernel
    sigma2 = p["variance"][:, 0][:, None]
    floor = variance_floor(om, Xq)
    report = {}
    worst_full = 0.0
    fails = []
    for what, o_call, g_call in (("noiseless", om.predict_noiseless, gm.predict_noiseless), ("with_noise", om.predict, gm.predict)):
        mu_o, v_o = o_call(Xq)
        mu_g, v_g = g_call(Xq)
        check_close(mu_g, mu_o, "%s %s mean" % (name, what))
        err = v_g - v_o
        bound = RTOL * np.abs(v_o) + 64 * floor
        ratio_full = float(np.max(np.abs(err) / bound))
        worst_full = max(worst_full, ratio_full)
        dec = per_decade(v_o, err, sigma2)
        report[what] = {"max_err_over_full_precision_bound": ratio_full, "max_abs_err_over_sigma2": float(np.max(np.abs(err) / sigma2)),
                        "min_var_over_sigma2": float(np.min(v_o / sigma2)),
                        "per_decade_of_var_over_sigma2": {str(e): {"count": c, "max_rel_err": r} for e, (c, r) in dec.items()}}
        print("%s n=%d %s kernel=%s: max |err| / (1e-9 |v| + 64 floor) = %.3f, max |err| / sigma^2 = %.2e" %
              (name, p["n"], what, kernel, ratio_full, report[what]["max_abs_err_over_sigma2"]))
        for e, (c, r) in dec.items():
            print("    var/sigma^2 in [1e%+d, 1e%+d): %5d candidates, max rel. err %.2e" % (e, e + 1, c, r))
        if kernel == "int8":
            if not np.all(np.abs(err) <= bound + ABS6 * sigma2):
                fails.append("%s %s int8: beyond the documented absolute term" % (name, what))
        elif not np.all(np.abs(err) <= bound):
            fails.append("%s %s %s: %.3f x the full-precision bound" % (name, what, kernel, ratio_full))
    _record("%s/n=%d/%s" % (name, p["n"], kernel), report)
    assert not fails, fails
    assert report["noiseless"]["min_var_over_sigma2"] < 1e-4      # the sweep does reach the noise level
    if kernel == "int8":
        # recorded, not required: the 6-digit split exceeds the full-precision bound once var << sigma^2
        print("    6-digit split: %.1f x the full-precision bound at its worst candidate" % worst_full)
