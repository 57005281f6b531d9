"""Cost of the int8 posterior kernel's model preparation: bopl_set_model (host digit split, threaded) and bopl_fit_model (device
oz_scale / oz_digits / oz_stage kernels) with posterior_kernel='int8' vs 'fp64'.  Run on the GPU box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
from bopl_b200.synthetic import make_problem

out = {}
for n, H in ((2000, 1), (2000, 10)):
    p = make_problem("S", n=n, N=8, H=H)
    st = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_winv=False, want_linv=False)
    rec = {}
    for k in ("fp64", "int8", "int8x7"):
        gm = MultiOutputGP(p["m"], n_samples=H, posterior_kernel=k)
        gm.engine.set_model(st)
        best_u, best_f = 1e9, 1e9
        for _ in range(3):
            t = time.perf_counter()
            gm.engine.set_model(st)
            torch.cuda.synchronize()
            best_u = min(best_u, time.perf_counter() - t)
            t = time.perf_counter()
            gm.engine.fit_model(p["X"], list(p["Y"]), p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_winv=False, want_linv=False)
            torch.cuda.synchronize()
            best_f = min(best_f, time.perf_counter() - t)
        rec[k] = {"set_model_ms": 1e3 * best_u, "fit_model_ms": 1e3 * best_f}
    out["n%d_H%d" % (n, H)] = rec
print(json.dumps(out))
