"""Latency of one Thompson-sampling objective evaluation (uTS.py:40-50) along a growing sample path: the device's rank-1
growth (`bopl_path_eval_host`) vs the reference's algorithm on the host (append + refit from scratch with LAPACK — the
oracle's restatement, all host cores).  Prints one JSON object.  GPU only."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from bopl_b200.synthetic import make_problem
from bopl_b200.sampling_policies import SamplePath
from oracle.thompson import SamplePathOracle
from tests.helpers import gpu_model, oracle_model

out = {}
for n, T in ((164, 600), (2000, 600)):
    p = make_problem("S", n=n, N=4, n_w=2, L=4, noise=1e-4)
    gm = gpu_model(p)
    path = SamplePath(gm, capacity=T + 8)
    rs = np.random.RandomState(0)
    X = rs.uniform(p["lo"], p["hi"], size=(T, p["d"]))
    Z = rs.standard_normal((T, p["m"]))
    marks, t0 = {}, time.perf_counter()
    seg0 = t0
    for t in range(T):
        path.evaluate(X[t], Z[t])
        if (t + 1) % 100 == 0:
            now = time.perf_counter()
            marks["t=%d..%d" % (t - 99, t)] = (now - seg0) / 100 * 1e3
            seg0 = now
    total = time.perf_counter() - t0
    gm.engine.path_end()
    # reference algorithm on the host: a few evaluations at the start and (n = 164 only) deep into a path
    om = oracle_model(p)
    sp = SamplePathOracle([om.gps[j][0] for j in range(p["m"])])
    k = 5 if n > 1000 else 40
    t0 = time.perf_counter()
    for t in range(k):
        sp.evaluate(X[t], Z[t])
    host_ms = (time.perf_counter() - t0) / k * 1e3
    out["n=%d,m=%d" % (n, p["m"])] = {"device_ms_per_eval_by_path_segment": marks, "device_ms_per_eval_mean": total / T * 1e3,
                                      "device_600_evals_s": total, "host_refit_ms_per_eval_at_path_start": host_ms,
                                      "host_evals_timed": k}
print(json.dumps(out))
