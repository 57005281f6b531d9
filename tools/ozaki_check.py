"""INT8-tensor-pipe posterior kernel ($OZ_KERNEL = int8x7 (default) | int8; $BOPL_OZ_CLUSTER=0: single-CTA launch) against the fp64
DMMA kernel and the oracle.   usage: python tools/ozaki_check.py n N [time]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model, oracle_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
p = make_problem("S", n=n, N=N)
goz = gpu_model(p, gradients=False, posterior_kernel=os.environ.get("OZ_KERNEL", "int8x7"))
gdm = gpu_model(p, gradients=False, posterior_kernel="fp64")
mo, vo = goz.predict(p["Xc"])
md, vd = gdm.predict(p["Xc"])
print("ozaki vs dmma: mean %.3e  var rel %.3e" % (np.max(np.abs(mo - md)) / np.max(np.abs(md)), np.max(np.abs(vo - vd) / np.abs(vd))), flush=True)
if n <= 600 and N <= 4096:
    om = oracle_model(p)
    mr, vr = om.predict(p["Xc"])
    print("ozaki vs oracle: mean %.3e var rel %.3e | dmma vs oracle: var rel %.3e"
          % (np.max(np.abs(mo - mr)) / np.max(np.abs(mr)), np.max(np.abs(vo - vr) / np.abs(vr)), np.max(np.abs(vd - vr) / np.abs(vr))))
if len(sys.argv) > 3:
    Xd = goz.engine.to_dev(p["Xc"])
    for name, g in (("ozaki", goz), ("dmma", gdm)):
        for _ in range(2):
            g.engine.posterior_dev(Xd, 0)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            g.engine.posterior_dev(Xd, 0)
        torch.cuda.synchronize()
        print("%s: %.2f ms per posterior call (N=%d, n=%d)" % (name, (time.perf_counter() - t) / 3 * 1e3, N, n))
