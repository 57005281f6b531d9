"""Kernel-level view of one value+gradient call at N=20, n=2000 (run under ncu --metrics gpu__time_duration.sum)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model, gpu_acquisition
p = make_problem("S", n=2000, N=64, n_w=25, L=10)
gm = gpu_model(p)
acq = gpu_acquisition(p, gm)
X = p["Xc"][:20]
for _ in range(3):
    acq._compute_acq_withGradients(X)
torch.cuda.synchronize()
