"""One device fit (+ one likelihood/gradient evaluation) at the benchmark shape — the command the ncu launch list of the fit is taken from."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bopl_b200.models import MultiOutputGP
from bopl_b200.synthetic import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = make_problem("S", n=n, N=8, H=H)
gm = MultiOutputGP(p["m"], n_samples=H)
for _ in range(2):
    gm.engine.fit_model(p["X"], list(p["Y"]), p["kinds"], p["variance"], p["lengthscale"], p["noise"])
if H == 1:
    Yc = p["Y"] - p["Y"].mean(axis=1, keepdims=True)
    for _ in range(2):
        gm.engine.log_marginal(p["X"], Yc, p["kinds"], p["variance"][:, 0], p["lengthscale"][:, 0], p["noise"][:, 0])
torch.cuda.synchronize()
print("ok")
