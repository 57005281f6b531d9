import sys, time, numpy as np
sys.path.insert(0, ".")
import torch
from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model, gpu_acquisition
p = make_problem("S", n=2000, N=64, n_w=25, L=10)
gm = gpu_model(p); acq = gpu_acquisition(p, gm)
for N in (1, 20):
    X = p["Xc"][:N]
    for _ in range(5): acq._compute_acq_withGradients(X)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): acq._compute_acq_withGradients(X)
    torch.cuda.synchronize(); print("N=%d value+gradient %.3f ms" % (N, (time.perf_counter() - t0) / 50 * 1e3))
