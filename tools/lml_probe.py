"""`bopl_log_marginal` alone at the benchmark shape: wall time per call (plain run) and the command for its ncu launch list."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bopl_b200.models import MultiOutputGP
from bopl_b200.synthetic import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = make_problem("S", n=n, N=8)
gm = MultiOutputGP(p["m"], n_samples=1)
Yc = p["Y"] - p["Y"].mean(axis=1, keepdims=True)
a = (p["X"], Yc, p["kinds"], p["variance"][:, 0], p["lengthscale"][:, 0], p["noise"][:, 0])
gm.engine.log_marginal(*a)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    t = time.perf_counter()
    gm.engine.log_marginal(*a)
    ts.append(time.perf_counter() - t)
tv = []
for _ in range(reps):
    t = time.perf_counter()
    gm.engine.log_marginal(*a, want_grad=False)
    tv.append(time.perf_counter() - t)
print("n=%d lml+grad ms: min %.3f median %.3f | value only: min %.3f median %.3f"
      % (n, 1e3 * min(ts), 1e3 * float(np.median(ts)), 1e3 * min(tv), 1e3 * float(np.median(tv))))
