"""Model-update latency: host LAPACK fit + upload (bopl_set_model) vs the device fit (bopl_fit_model), and one batched
likelihood+gradient evaluation (bopl_log_marginal) vs the host objective of bopl_b200/fit.py.  Run on the GPU box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from bopl_b200.fit import GPObjective
from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
from bopl_b200.synthetic import make_problem

out = {}
for n, H in ((164, 1), (500, 1), (2000, 1), (2000, 10)):
    p = make_problem("S", n=n, N=8, H=H)
    gm = MultiOutputGP(p["m"], n_samples=H)
    t = time.perf_counter()
    st = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"])
    t_host_fit = time.perf_counter() - t
    t = time.perf_counter()
    gm.engine.set_model(st)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        gm.engine.fit_model(p["X"], list(p["Y"]), p["kinds"], p["variance"], p["lengthscale"], p["noise"])
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    rec = {"host_fit_ms": 1e3 * t_host_fit, "upload_ms": 1e3 * t_upload, "device_fit_ms": 1e3 * best}
    if H == 1:
        Yc = p["Y"] - p["Y"].mean(axis=1, keepdims=True)
        a = (p["X"], Yc, p["kinds"], p["variance"][:, 0], p["lengthscale"][:, 0], p["noise"][:, 0])
        gm.engine.log_marginal(*a)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            gm.engine.log_marginal(*a)
            best = min(best, time.perf_counter() - t)
        rec["device_lml_grad_ms_all_attributes"] = 1e3 * best
        obj = GPObjective(p["X"], p["Y"][0], "Mat52", True, 1e-6)
        x0 = obj.initial()
        t = time.perf_counter()
        obj.value_and_grad(x0)
        rec["host_lml_grad_ms_per_attribute"] = 1e3 * (time.perf_counter() - t)
    out["n=%d,m=3,H=%d" % (n, H)] = rec
    print(n, H, rec, file=sys.stderr)
print(json.dumps(out))
