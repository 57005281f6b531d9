#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN CODE for the hot path, in the build container.

The reference cannot be imported as a package here (vendored GPy needs `paramz`, uEI.py imports `pathos`; SURVEY.md
§8c), so this script executes the reference's source for the path *piecewise*, straight from /root/reference, with
nothing copied into the repo:

  * whole modules that only need a stub or two are imported from their files with the missing imports stubbed:
      GPy/util/linalg.py, GPy/util/diag.py  (config stub: cython not working — the repo ships no built .so either)
      bopl/utility/utility.py, bopl/utility/utility_distribution.py (stub `core.preference_encoder` <- the real file)
      GPyOpt/acquisitions/base.py, bopl/sampling_policies/acquisition_functions/uEI.py, uEI_affine.py
      (stubs: pathos ProcessingPool -> serial map; GPyOpt.core.task.cost.constant_cost_withGradients)
  * classes whose base classes need paramz are rebuilt from their METHOD SOURCE (ast-extracted FunctionDefs, compiled
    unchanged, decorators dropped) on plain containers:
      Stationary / Matern52 / Matern32 / ExpQuad (stationary.py), RBF (rbf.py): K, _scaled_dist, _unscaled_dist, Kdiag,
        _inv_dist, dK_dr_via_X, gradients_X, _gradients_X_pure, K_of_r, dK_dr
      ExactGaussianInference.inference; Posterior / PosteriorExact (whole ClassDefs); Gaussian.predictive_values,
        predictive_variance2, gaussian_variance; Standardize (whole ClassDef)
      GP: _raw_predict, predict, predict_noiseless, posterior_mean, posterior_variance, posterior_variance_noiseless,
        posterior_mean_gradient, posterior_variance_gradient
      GPModel (gpmodel.py): predict, predict_noiseless, posterior_mean, posterior_variance,
        posterior_variance_noiseless, posterior_mean_gradient, posterior_variance_gradient, set_hyperparameters
      MultiOutputGP (multi_outputGP.py): every predict / posterior_* method + set_hyperparameters
  * utility callables: the `utility_func` / `utility_gradient` FunctionDefs of experiments/test_dtlz1a_linear.py,
    test_dtlz2a_quad.py, test_vlmop3_exp.py, compiled unchanged.

Hyper-parameters are supplied (the fit's MAP + HMC is outside the path); everything downstream of them — Ky, Cholesky,
woodbury vector / inverse, predictions, gradients, the acquisition loops — is the reference's own arithmetic.

Outputs (committed): one .npz per case with the inputs' identifying parameters (regenerated bit-exactly by
bopl_b200.synthetic.make_problem) and the reference's outputs.  `python tests/golden/make_golden.py` rewrites them.
"""
import ast
import configparser
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GPY = os.path.join(REF, "bopl/aux_software/GPy")
GPYOPT = os.path.join(REF, "bopl/aux_software/GPyOpt")


# ------------------------------------------------------------------------------------------- loading machinery
def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _tree(path):
    with open(path) as f:
        return ast.parse(f.read(), filename=path)


def _find_class(tree, cls):
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            return node
    raise KeyError(cls)


def harvest_methods(path, cls, names, ns):
    """Compile the named methods of class `cls` in file `path` unchanged (decorators dropped) inside namespace ns."""
    node = _find_class(_tree(path), cls)
    out = {}
    for item in node.body:
        if isinstance(item, ast.FunctionDef) and item.name in names:
            item.decorator_list = []
            mod = ast.Module(body=[item], type_ignores=[])
            ast.fix_missing_locations(mod)
            scope = dict(ns)
            exec(compile(mod, path, "exec"), scope)
            out[item.name] = scope[item.name]
    missing = set(names) - set(out)
    if missing:
        raise KeyError("%s: %s missing in %s" % (cls, sorted(missing), path))
    return out


def exec_classes(path, classes, ns):
    tree = _tree(path)
    scope = dict(ns)
    for cls in classes:
        mod = ast.Module(body=[_find_class(tree, cls)], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, path, "exec"), scope)
    return scope


def harvest_functions(path, names):
    """Nested / module-level FunctionDefs by name (the experiment scripts define the utilities inside __main__)."""
    tree = _tree(path)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in out:
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            out[node.name] = mod
    return out


def build_reference():
    """Returns a namespace of reference-code-backed constructors."""
    R = types.SimpleNamespace()
    # --- GPy.util.linalg / diag (real files; config stub says cython is not working, as in the shipped repo)
    cfg = configparser.ConfigParser()
    cfg.read_dict({"cython": {"working": "False"}})
    _pkg("refGPy_util", os.path.join(GPY, "util"))
    cmod = types.ModuleType("refGPy_util.config")
    cmod.config = cfg
    sys.modules["refGPy_util.config"] = cmod
    linalg = _load("refGPy_util.linalg", os.path.join(GPY, "util/linalg.py"))
    diag = _load("refGPy_util.diag", os.path.join(GPY, "util/diag.py"))
    util_ns = types.SimpleNamespace(diag=diag, linalg=linalg)
    R.linalg = linalg

    # --- kernels: method source from stationary.py / rbf.py
    kns = dict(np=np, tdot=linalg.tdot, util=util_ns, config=cfg)
    stat = os.path.join(GPY, "kern/src/stationary.py")
    base = harvest_methods(stat, "Stationary", ["K", "dK_dr_via_X", "_unscaled_dist", "_scaled_dist", "Kdiag", "_inv_dist",
                                                "gradients_X", "_gradients_X_pure"], kns)
    kinds = {"Mat52": harvest_methods(stat, "Matern52", ["K_of_r", "dK_dr"], kns),
             "Mat32": harvest_methods(stat, "Matern32", ["K_of_r", "dK_dr"], kns),
             "ExpQuad": harvest_methods(stat, "ExpQuad", ["K_of_r", "dK_dr"], kns),
             "rbf": harvest_methods(os.path.join(GPY, "kern/src/rbf.py"), "RBF", ["K_of_r", "dK_dr"], kns)}

    def make_kern(kind, variance, lengthscale, input_dim):
        cls = type("Ref" + kind, (object,), dict(base, **kinds[kind]))
        k = cls()
        k.name, k.variance, k.lengthscale, k.ARD, k.input_dim = kind, np.array([variance]), np.asarray(lengthscale, float), True, input_dim
        return k
    R.make_kern = make_kern

    # --- Posterior classes, inference, likelihood, normalizer
    pns = dict(np=np, pdinv=linalg.pdinv, dpotrs=linalg.dpotrs, dpotri=linalg.dpotri, symmetrify=linalg.symmetrify,
               jitchol=linalg.jitchol, dtrtrs=linalg.dtrtrs, tdot=linalg.tdot,
               VariationalPosterior=type("VariationalPosterior", (object,), {}))
    post = exec_classes(os.path.join(GPY, "inference/latent_function_inference/posterior.py"), ["Posterior", "PosteriorExact"], pns)
    ins = dict(np=np, pdinv=linalg.pdinv, dpotrs=linalg.dpotrs, tdot=linalg.tdot, diag=diag,
               log_2_pi=np.log(2 * np.pi), Posterior=post["PosteriorExact"])
    R.inference = harvest_methods(os.path.join(GPY, "inference/latent_function_inference/exact_gaussian_inference.py"),
                                  "ExactGaussianInference", ["inference"], ins)["inference"]
    lik = harvest_methods(os.path.join(GPY, "likelihoods/gaussian.py"), "Gaussian",
                          ["predictive_values", "predictive_variance2", "gaussian_variance"], dict(np=np))
    lik["exact_inference_gradients"] = lambda self, dL_dKdiag, Y_metadata=None: dL_dKdiag.sum()
    R.Gaussian = type("RefGaussian", (object,), lik)
    norm = exec_classes(os.path.join(GPY, "util/normalizer.py"), ["_Norm", "Standardize"], dict(np=np))
    R.Standardize = norm["Standardize"]

    # --- GP (BOPL-added posterior_* methods + predict)
    gp_m = harvest_methods(os.path.join(GPY, "core/gp.py"), "GP",
                           ["_raw_predict", "predict", "predict_noiseless", "posterior_mean", "posterior_variance",
                            "posterior_variance_noiseless", "posterior_mean_gradient", "posterior_variance_gradient"], dict(np=np))
    R.GP = type("RefGP", (object,), gp_m)

    # --- GPyOpt GPModel forwarding (+ clip) and BOPL MultiOutputGP stacking
    gm = harvest_methods(os.path.join(GPYOPT, "models/gpmodel.py"), "GPModel",
                         ["predict", "predict_noiseless", "posterior_mean", "posterior_variance", "posterior_variance_noiseless",
                          "posterior_mean_gradient", "posterior_variance_gradient", "set_hyperparameters"], dict(np=np))
    R.GPModel = type("RefGPModel", (object,), gm)
    mo = harvest_methods(os.path.join(REF, "bopl/models/multi_outputGP.py"), "MultiOutputGP",
                         ["number_of_hyps_samples", "set_hyperparameters", "get_evaluated_points", "predict", "predict_noiseless",
                          "posterior_mean", "posterior_mean_at_evaluated_points", "posterior_variance",
                          "posterior_mean_gradient", "posterior_variance_gradient"], dict(np=np))
    mo["analytical_gradient_prediction"] = True
    R.MultiOutputGP = type("RefMultiOutputGP", (object,), mo)

    # --- utility objects (real modules)
    core = _pkg("core")
    core.preference_encoder = _load("refcore_pe", os.path.join(REF, "bopl/core/preference_encoder.py")).preference_encoder
    R.Utility = _load("ref_utility", os.path.join(REF, "bopl/utility/utility.py")).Utility
    R.UtilityDistribution = _load("ref_utility_distribution", os.path.join(REF, "bopl/utility/utility_distribution.py")).UtilityDistribution

    # --- acquisitions (real modules, stubbed imports)
    pathos = _pkg("pathos")
    pm = types.ModuleType("pathos.multiprocessing")

    class ProcessingPool(object):                      # the pool only distributes rows; a serial map gives the same numbers
        def __init__(self, *a, **k):
            pass

        def map(self, f, X):
            return [f(x) for x in X]
    pm.ProcessingPool = ProcessingPool
    sys.modules["pathos.multiprocessing"] = pm
    pathos.multiprocessing = pm
    for name in ["aux_software", "aux_software.GPyOpt", "aux_software.GPyOpt.acquisitions", "aux_software.GPyOpt.core",
                 "aux_software.GPyOpt.core.task"]:
        _pkg(name)
    cost = types.ModuleType("aux_software.GPyOpt.core.task.cost")
    cost.constant_cost_withGradients = lambda x: (np.ones(x.shape[0])[:, None], np.zeros(x.shape))
    sys.modules["aux_software.GPyOpt.core.task.cost"] = cost
    _load("aux_software.GPyOpt.acquisitions.base", os.path.join(GPYOPT, "acquisitions/base.py"))
    R.uEI = _load("ref_uEI", os.path.join(REF, "bopl/sampling_policies/acquisition_functions/uEI.py")).uEI
    R.uEI_affine = _load("ref_uEI_affine", os.path.join(REF, "bopl/sampling_policies/acquisition_functions/uEI_affine.py")).uEI_affine
    R.uEI_constrained = _load("ref_uEI_constrained",
                              os.path.join(REF, "bopl/sampling_policies/acquisition_functions/uEI_constrained.py")).uEI_constrained

    # --- utility callables of the experiment scripts
    def exp_utils(script, m):
        mods = harvest_functions(os.path.join(REF, "experiments", script), ["utility_func", "utility_gradient"])
        scope = dict(np=np, m=m)
        for mod in mods.values():
            exec(compile(mod, script, "exec"), scope)
        return scope["utility_func"], scope["utility_gradient"]
    R.exp_utils = exp_utils
    return R


def reference_model(R, p):
    """The reference's object graph for problem dict p: MultiOutputGP -> GPModel -> GP(kern, likelihood, posterior)."""
    outs = []
    for j in range(p["m"]):
        insts = []
        for h in range(p["H"]):
            g = R.GP()
            g.X = p["X"].copy()
            g._predictive_variable = g.X
            g.kern = R.make_kern(p["kinds"][j], p["variance"][j, h], p["lengthscale"][j, h], p["d"])
            g.likelihood = R.Gaussian()
            g.likelihood.variance = np.array([p["noise"][j, h]])
            g.mean_function = None
            g.normalizer = R.Standardize()
            Y = p["Y"][j].reshape(-1, 1)
            g.normalizer.scale_by(Y)                               # GPy/core/gp.py:52-62
            g.Y_normalized = g.normalizer.normalize(Y)
            g.posterior, g._log_marginal_likelihood, g.grad_dict = R.inference(
                None, g.kern, g.X, g.likelihood, g.Y_normalized, None, None)   # gp.py:247-260
            insts.append(g)
        gm = R.GPModel()
        gm.model_instances = insts
        gm.model = insts[0]
        outs.append(gm)
    mo = R.MultiOutputGP()
    mo.output, mo.output_dim, mo.n_samples, mo.name = outs, p["m"], p["H"], "multi-output GP"
    return mo


CASES = [  # name, kwargs for make_problem, experiment script providing the utility callables
    ("dtlz1a_linear", dict(N=12, n=30, n_w=6, L=4, H=2), "test_dtlz1a_linear.py"),
    ("dtlz2a_quad", dict(N=12, n=30, n_w=6, L=8, H=2), "test_dtlz2a_quad.py"),
    ("vlmop3_exp", dict(N=12, n=30, n_w=6, L=4, H=2), "test_vlmop3_exp.py"),
    ("portfolio1", dict(N=60, n=30, n_w=6, L=4, H=1), "test_dtlz1a_linear.py"),
]


def main():
    from bopl_b200.synthetic import make_problem
    R = build_reference()
    for name, kw, script in CASES:
        p = make_problem(name, **kw)
        model = reference_model(R, p)
        func, grad = R.exp_utils(script, p["m"])
        theta = p["theta"]
        if p["use_full_support"]:
            dist = R.UtilityDistribution(support=theta, prob_dist=np.array(p["prob"]), utility_func=func)
        else:
            dist = R.UtilityDistribution(prior_sample_generator=lambda n_s, seed=None: theta[:n_s], utility_func=func,
                                         use_full_support=False)
        util = R.Utility(func=func, gradient=grad, parameter_distribution=dist)
        X = p["Xc"]
        out = dict(name=name, acq="uEI", N=kw["N"], n=kw["n"], n_w=kw["n_w"], L=kw["L"], H=kw["H"], Xc=X, Z=p["Z"], theta=theta,
                   use_full_support=p["use_full_support"], prob=np.array(p["prob"]) if p["prob"] is not None else np.empty(0))
        # posterior layer under every hyper-sample (reference MultiOutputGP surface)
        for h in range(p["H"]):
            model.set_hyperparameters(h)
            mu, var = model.predict(X)
            out["mean_h%d" % h], out["var_h%d" % h] = mu, var
            out["pmean_h%d" % h] = model.posterior_mean(X)
            out["pvar_h%d" % h] = model.posterior_variance(X)
            _, out["var_noiseless_h%d" % h] = model.predict_noiseless(X)
            out["dmean_h%d" % h] = model.posterior_mean_gradient(X)
            out["dvar_h%d" % h] = model.posterior_variance_gradient(X)
            out["F_h%d" % h] = model.posterior_mean_at_evaluated_points()
        # acquisition layer
        acq = R.uEI(model, None, optimizer=None, utility=util)
        acq.W_samples = p["Z"].copy()
        acq.utility_parameter_samples = theta
        acq.number_of_gp_hyps_samples = p["H"]
        model.set_hyperparameters(0)
        out["acq_sequential"] = acq._compute_acq(X, parallel=False)
        model.set_hyperparameters(0)
        out["acq_parallel"] = acq._compute_acq(X, parallel=True)
        model.set_hyperparameters(0)
        out["acq_grad_f"], out["acq_grad_df"] = acq._compute_acq_withGradients(X)
        np.savez(os.path.join(HERE, "uEI_%s.npz" % name), **out)
        print("wrote uEI_%s.npz  acq range [%.3e, %.3e]" % (name, out["acq_parallel"].min(), out["acq_parallel"].max()))
        if p["utility"] == "linear":
            aff = R.uEI_affine(model, None, optimizer=None, utility=util)
            aff.utility_parameter_samples = theta
            aff.number_of_gp_hyps_samples = p["H"]                    # the reference hard-codes 10 (uEI_affine.py:36)
            o2 = {k: out[k] for k in ("name", "N", "n", "n_w", "L", "H", "Xc", "Z", "theta", "use_full_support", "prob")}
            o2["acq"] = "uEI_affine"
            o2["acq_value"] = aff._compute_acq(X)
            o2["acq_grad_f"], o2["acq_grad_df"] = aff._compute_acq_withGradients(X)
            np.savez(os.path.join(HERE, "uEI_affine_%s.npz" % name), **o2)
            print("wrote uEI_affine_%s.npz" % name)


BIG_CASES = [  # several 64-row block rows and more candidates than the small-batch limit: the blocked posterior kernels' own regime
    ("S", dict(N=200, n=300, n_w=8, L=4, H=1), "test_dtlz1a_linear.py"),
    ("dtlz2a_quad", dict(N=150, n=200, n_w=6, L=8, H=1), "test_dtlz2a_quad.py"),
]


def main_big():
    """Value path only (posterior + uEI values) at sizes where the blocked posterior kernels do block-row updates: big_*.npz."""
    from bopl_b200.synthetic import make_problem
    R = build_reference()
    for name, kw, script in BIG_CASES:
        p = make_problem(name, **kw)
        model = reference_model(R, p)
        func, grad = R.exp_utils(script, p["m"])
        theta = p["theta"]
        if p["use_full_support"]:
            dist = R.UtilityDistribution(support=theta, prob_dist=np.array(p["prob"]), utility_func=func)
        else:
            dist = R.UtilityDistribution(prior_sample_generator=lambda n_s, seed=None: theta[:n_s], utility_func=func,
                                         use_full_support=False)
        util = R.Utility(func=func, gradient=grad, parameter_distribution=dist)
        X = p["Xc"]
        out = dict(name=name, acq="big", N=kw["N"], n=kw["n"], n_w=kw["n_w"], L=kw["L"], H=kw["H"])
        model.set_hyperparameters(0)
        out["mean"], out["var"] = model.predict(X)
        _, out["var_noiseless"] = model.predict_noiseless(X)
        acq = R.uEI(model, None, optimizer=None, utility=util)
        acq.W_samples = p["Z"].copy()
        acq.utility_parameter_samples = theta
        acq.number_of_gp_hyps_samples = p["H"]
        model.set_hyperparameters(0)
        out["acq_sequential"] = acq._compute_acq(X, parallel=False)
        np.savez(os.path.join(HERE, "big_%s.npz" % name), **out)
        print("wrote big_%s.npz  var range [%.3e, %.3e]  acq max %.3e" % (name, out["var"].min(), out["var"].max(), out["acq_sequential"].max()))


CONSTRAINED_CASES = [  # (problem name, kwargs): m = 2 uses the reference's own callable of experiments/test_portfolio2.py
    ("dtlz1a_linear", dict(N=40, n=30, L=5, H=2, utility="constrained")),       # m = 2: one constraint
    ("vlmop3_exp", dict(N=40, n=30, L=5, H=1, utility="constrained")),          # m = 3: two constraints
]


def main_constrained():
    from bopl_b200.synthetic import make_problem
    from oracle import utilities
    R = build_reference()
    for name, kw in CONSTRAINED_CASES:
        p = make_problem(name, **kw)
        model = reference_model(R, p)
        if p["m"] == 2:      # the reference's own utility callable (one scalar threshold)
            mods = harvest_functions(os.path.join(REF, "experiments", "test_portfolio2.py"), ["utility_func"])
            scope = dict(np=np)
            np.infty = np.inf                                    # numpy 2 dropped the alias the script uses
            exec(compile(mods["utility_func"], "test_portfolio2.py", "exec"), scope)
            func = scope["utility_func"]
        else:                # the script's callable handles one constraint only; same rule for m-1 thresholds
            func = utilities.constrained_func
        theta = p["theta"]
        dist = R.UtilityDistribution(prior_sample_generator=lambda n_s, seed=None: theta[:n_s], utility_func=func,
                                     use_full_support=False)
        util = R.Utility(func=func, parameter_distribution=dist)
        acq = R.uEI_constrained(model, None, optimizer=None, utility=util)
        acq.utility_parameter_samples = theta
        acq.number_of_gp_hyps_samples = p["H"]
        X = p["Xc"]
        out = dict(name=name, acq="uEI_constrained", N=kw["N"], n=kw["n"], L=kw["L"], H=kw["H"], Xc=X, theta=theta)
        model.set_hyperparameters(0)
        out["acq_value"] = acq._compute_acq(X)
        model.set_hyperparameters(0)
        out["acq_grad_f"], out["acq_grad_df"] = acq._compute_acq_withGradients(X)
        np.savez(os.path.join(HERE, "uEIc_%s.npz" % name), **out)
        print("wrote uEIc_%s.npz  acq range [%.3e, %.3e]" % (name, out["acq_value"].min(), out["acq_value"].max()))


MARGINAL_CASES = [  # (problem, kwargs, experiment script with the utility callables, affine flag)
    ("dtlz1a_linear", dict(N=10, n=30, n_w=6, L=3, H=2), "test_dtlz1a_linear.py", True),
    ("dtlz2a_quad", dict(N=10, n=30, n_w=6, L=8, H=2), "test_dtlz2a_quad.py", False),
    ("vlmop3_exp", dict(N=10, n=30, n_w=6, L=3, H=1), "test_vlmop3_exp.py", False),
]


def main_marginal():
    """The objective closures of the reference's AcquisitionOptimizer._current_marginal_argmax (the baseline-point search,
    acquisition_optimizer.py:139-249), captured by intercepting `optimize_inner_func`, evaluated on the test candidates."""
    from bopl_b200.synthetic import make_problem
    R = build_reference()
    for name in ["aux_software.GPyOpt.experiment_design", "aux_software.GPyOpt.optimization",
                 "aux_software.GPyOpt.optimization.optimizer", "aux_software.GPyOpt.optimization.anchor_points_generator"]:
        _pkg(name)
    sys.modules["aux_software.GPyOpt.experiment_design"].initial_design = None
    om = sys.modules["aux_software.GPyOpt.optimization.optimizer"]
    om.apply_optimizer = om.apply_optimizer_inner = None
    om.choose_optimizer = lambda name, bounds: None
    ag = sys.modules["aux_software.GPyOpt.optimization.anchor_points_generator"]
    ag.ObjectiveAnchorPointsGenerator = ag.ThompsonSamplingAnchorPointsGenerator = None
    AO = _load("ref_acquisition_optimizer", os.path.join(REF, "bopl/optimization_services/acquisition_optimizer.py")).AcquisitionOptimizer
    for name, kw, script, affine in MARGINAL_CASES:
        p = make_problem(name, **kw)
        model = reference_model(R, p)
        func, grad = R.exp_utils(script, p["m"])
        theta = p["theta"]
        dist = R.UtilityDistribution(prior_sample_generator=lambda n_s, seed=None: theta[:n_s], utility_func=func,
                                     use_full_support=False)
        util = R.Utility(func=func, gradient=grad, parameter_distribution=dist, affine=affine)
        space = types.SimpleNamespace(model_dimensionality=p["d"], config_space_expanded=[None] * p["d"],
                                      get_bounds=lambda: [(p["lo"], p["hi"])] * p["d"])
        opt = AO(space, model, util)
        opt.number_of_gp_hyps_samples = p["H"]
        got = {}

        def capture(f=None, df=None, f_df=None, **k):
            got["f"], got["f_df"] = f, f_df
            return np.zeros((1, p["d"])), 0.0
        opt.optimize_inner_func = capture
        out = dict(name=name, acq="marginal", N=kw["N"], n=kw["n"], n_w=kw["n_w"], L=kw["L"], H=kw["H"], affine=affine,
                   Xc=p["Xc"], theta=theta)
        for l in range(min(2, len(theta))):
            np.random.seed(1234 + l)                      # the MC branch draws its 50 Z samples from the global stream
            opt._current_marginal_argmax(theta[l])
            out["f_%d" % l] = got["f"](p["Xc"].copy())
            v, dv = got["f_df"](p["Xc"].copy())
            out["fdf_f_%d" % l], out["fdf_df_%d" % l] = v, dv
        np.savez(os.path.join(HERE, "marginal_%s.npz" % name), **out)
        print("wrote marginal_%s.npz  value range [%.3e, %.3e]" % (name, out["f_0"].min(), out["f_0"].max()))


THOMPSON_CASES = [("dtlz1a_linear", dict(N=4, n=30, n_w=2, L=4, H=2), "test_dtlz1a_linear.py"),
                  ("dtlz2a_quad", dict(N=4, n=30, n_w=2, L=8, H=2), "test_dtlz2a_quad.py"),
                  ("vlmop3_exp", dict(N=4, n=30, n_w=2, L=4, H=2), "test_vlmop3_exp.py")]


def main_thompson(T=14):
    """The objective closure of the reference's uTS.suggest_sample (sampling_policies/uTS.py:27-57), executed from the
    reference's own file: `apply_optimizer` is replaced by a recorder that calls the closure on a fixed sequence of points
    (one of them repeated, so the path revisits a point it already holds).  The per-attribute GP copies are the reference's
    `GP.posterior_samples_f` / `GP.set_XY` (method source compiled unchanged; `update_model(True)` runs the first statement of
    `parameters_changed`, gp.py:254 — the inference; the gradient plumbing after it needs paramz and has no effect on the
    path).  np.random.multivariate_normal is wrapped to record the (mean, cov, draw) of every call."""
    import copy
    from bopl_b200.synthetic import make_problem
    R = build_reference()
    gns = dict(np=np, ObsAr=lambda a: np.asarray(a), VariationalPosterior=type("VariationalPosterior", (object,), {}))
    gp_extra = harvest_methods(os.path.join(GPY, "core/gp.py"), "GP", ["posterior_samples_f", "set_XY"], gns)

    def update_model(self, flag):                                                   # gp.py:254
        if flag:
            self.posterior, self._log_marginal_likelihood, self.grad_dict = R.inference(
                None, self.kern, self.X, self.likelihood, self.Y_normalized, None, None)
    body = dict(R.GP.__dict__)
    body = {k: v for k, v in body.items() if not k.startswith("__")}
    body.update(gp_extra)
    body.update(update_model=update_model, output_dim=1, parameters=(),
                _predictive_variable=property(lambda self: self.X))                # gp.py:187-189
    RefGPPath = type("RefGPPath", (object,), body)
    gm_copy = harvest_methods(os.path.join(GPYOPT, "models/gpmodel.py"), "GPModel", ["get_copy_of_model_sample"],
                              dict(np=np, deepcopy=copy.deepcopy))
    mo_copy = harvest_methods(os.path.join(REF, "bopl/models/multi_outputGP.py"), "MultiOutputGP", ["get_copy_of_model_sample"],
                              dict(np=np))
    GPModelP = type("RefGPModelP", (R.GPModel,), gm_copy)
    MOP = type("RefMultiOutputGPP", (R.MultiOutputGP,), mo_copy)

    # --- the reference's uTS module, plumbing imports stubbed
    _pkg("sampling_policies")
    _load("sampling_policies.base", os.path.join(REF, "bopl/sampling_policies/base.py"))
    _pkg("optimization_services")
    ao = types.ModuleType("optimization_services.acquisition_optimizer")
    ao.ContextManager = lambda space, context=None: types.SimpleNamespace(noncontext_bounds=space.get_bounds())
    sys.modules["optimization_services.acquisition_optimizer"] = ao
    for name in ["aux_software", "aux_software.GPyOpt", "aux_software.GPyOpt.optimization"]:
        if name not in sys.modules:
            _pkg(name)
    ed = types.ModuleType("aux_software.GPyOpt.experiment_design")
    sys.modules["aux_software.GPyOpt.experiment_design"] = ed
    om = types.ModuleType("aux_software.GPyOpt.optimization.optimizer")
    sys.modules["aux_software.GPyOpt.optimization.optimizer"] = om
    rec = {}

    def apply_optimizer(optimizer, x0, f=None, context_manager=None, space=None, maxfevals=None, **k):
        rec["maxfevals"] = maxfevals
        rec["fvals"] = np.array([np.asarray(f(x)).reshape(-1)[0] for x in rec["Xseq"]])
        return np.atleast_2d(rec["Xseq"][-1]), None
    om.apply_optimizer = apply_optimizer
    om.choose_optimizer = lambda name, bounds: None
    ed.initial_design = lambda kind, space, n: np.atleast_2d(rec["Xseq"][0])
    uTS = _load("ref_uTS", os.path.join(REF, "bopl/sampling_policies/uTS.py")).uTS

    for name, kw, script in THOMPSON_CASES:
        p = make_problem(name, **kw)
        outs = []
        for j in range(p["m"]):
            insts = []
            for h in range(p["H"]):
                g = RefGPPath()
                g.kern = R.make_kern(p["kinds"][j], p["variance"][j, h], p["lengthscale"][j, h], p["d"])
                g.likelihood = R.Gaussian()
                g.likelihood.variance = np.array([p["noise"][j, h]])
                g.mean_function = None
                g.normalizer = R.Standardize()
                g.X = p["X"].copy()
                g.set_XY(p["X"].copy(), p["Y"][j].reshape(-1, 1).copy())           # gp.py:52-62 + 247-260 through the real set_XY
                insts.append(g)
            gm = GPModelP()
            gm.model_instances = insts
            gm.model = insts[0]
            outs.append(gm)
        mo = MOP()
        mo.output, mo.output_dim, mo.n_samples, mo.name = outs, p["m"], p["H"], "multi-output GP"
        func, grad = R.exp_utils(script, p["m"])
        theta = p["theta"]
        dist = R.UtilityDistribution(prior_sample_generator=lambda n_s, seed=None: theta[:n_s], utility_func=func,
                                     use_full_support=False)
        util = R.Utility(func=func, gradient=grad, parameter_distribution=dist)
        space = types.SimpleNamespace(get_bounds=lambda: [(p["lo"], p["hi"])] * p["d"])
        rs = np.random.RandomState(99)
        Xseq = rs.uniform(p["lo"], p["hi"], size=(T, p["d"]))
        Xseq[9] = Xseq[3]                                                           # a revisited point
        rec["Xseq"] = Xseq
        calls = []
        real_mvn = np.random.multivariate_normal

        def mvn(mean, cov, size=None):
            out = real_mvn(mean, cov, size)
            calls.append((float(np.asarray(mean).reshape(-1)[0]), float(np.asarray(cov).reshape(-1)[0]),
                          float(np.asarray(out).reshape(-1)[0])))
            return out
        np.random.seed(4321)
        np.random.multivariate_normal = mvn
        try:
            policy = uTS(mo, space, optimizer="CMA", utility=util)
            suggested = policy.suggest_sample()
        finally:
            np.random.multivariate_normal = real_mvn
        c = np.array(calls).reshape(T, p["m"], 3)
        out = dict(name=name, acq="uTS", n=kw["n"], H=kw["H"], seed=4321, Xseq=Xseq, theta=theta[:1], mean=c[:, :, 0],
                   var=c[:, :, 1], y=c[:, :, 2], fvals=rec["fvals"], suggested=suggested, maxfevals=rec["maxfevals"],
                   X_aux=policy.X_aux, Y_aux=np.stack([np.asarray(y)[:, 0] for y in policy.Y_aux]))
        np.savez(os.path.join(HERE, "uTS_%s.npz" % name), **out)
        print("wrote uTS_%s.npz  fvals [%.3e, %.3e]  min var %.3e" % (name, out["fvals"].min(), out["fvals"].max(), c[:, :, 1].min()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        {"main": main, "big": main_big, "constrained": main_constrained, "marginal": main_marginal, "thompson": main_thompson}[sys.argv[1]]()
    else:
        main()
        main_constrained()
        main_marginal()
        main_thompson()
