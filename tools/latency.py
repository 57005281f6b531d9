"""Latency of the small-batch calls the L-BFGS inner loop makes (value+gradient at N = 1 ... 64 points) and of a model
upload.  Prints one JSON line per shape.  GPU only."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model, gpu_acquisition

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

for n in (164, 2000):
    p = make_problem("S", n=n, N=4096, n_w=25, L=10)
    t0 = time.perf_counter(); gm = gpu_model(p); t_model = (time.perf_counter() - t0) * 1e3
    st = gm.state
    t0 = time.perf_counter(); gm.engine.set_model(st); t_up = (time.perf_counter() - t0) * 1e3
    acq = gpu_acquisition(p, gm)
    out = {"n": n, "fit_plus_upload_ms": t_model, "set_model_ms": t_up}
    for N in (1, 4, 20, 64, 256, 4096):
        X = p["Xc"][:N]
        out["predict_N%d_ms" % N] = timeit(lambda: gm.predict(X))
        out["acq_N%d_ms" % N] = timeit(lambda: acq._compute_acq(X))
        if N <= 256:
            out["acq_grad_N%d_ms" % N] = timeit(lambda: acq._compute_acq_withGradients(X))
    print(json.dumps(out))
