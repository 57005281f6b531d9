"""Posterior + uEI latency across batch sizes (the reference's own anchor batch is n_starting = 400 candidates, acquisition_optimizer.py:26) for the three
posterior kernels and both launch forms of the int8 kernel.  usage: python tools/batch_latency.py [n]   -> JSON on stdout"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_acquisition, gpu_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
p = make_problem("S", n=n, N=8192)
out = {"n": n, "m": p["m"], "d": p["d"], "what": "ms per call: posterior_dev (device tensors) / uEI._compute_acq (NumPy in, NumPy out)", "rows": {}}
variants = [("int8x7", None), ("int8x7_single_cta", "0"), ("fp64", None), ("int8", None)]
for name, cl in variants:
    if cl is not None:
        os.environ["BOPL_OZ_CLUSTER"] = cl
    gm = gpu_model(p, gradients=False, posterior_kernel=name.split("_")[0])
    os.environ.pop("BOPL_OZ_CLUSTER", None)
    acq = gpu_acquisition(p, gm)
    eng = gm.engine
    Xd = eng.to_dev(p["Xc"])
    for N in (33, 128, 400, 1024, 4096, 8192):
        Xs = Xd[:N].contiguous()
        for _ in range(3):
            eng.posterior_dev(Xs, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.posterior_dev(Xs, 0)
        e1.record()
        torch.cuda.synchronize()
        Xh = p["Xc"][:N]
        acq._compute_acq(Xh)
        t0 = time.perf_counter()
        for _ in range(5):
            acq._compute_acq(Xh)
        t = (time.perf_counter() - t0) / 5
        out["rows"].setdefault(str(N), {})[name] = {"posterior_ms": e0.elapsed_time(e1) / 10, "uei_call_ms": t * 1e3}
print(json.dumps(out, indent=1))
