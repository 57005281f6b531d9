"""Phase timeline of the int8 posterior kernel: CTA 0, first work item (BOPL_OZ_TRACE -> posterior_ozaki_kernel<S, true>).
usage: python tools/ozaki_trace.py [n] [N] [out]   -> per block row: durations in clocks of K*, wait, recombine, diagonal solve, digits,
and what the MMA thread did meanwhile."""
import os
import sys

out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/oz_trace.txt"
os.environ["BOPL_OZ_TRACE"] = out
os.environ["BOPL_POSTERIOR"] = "int8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bopl_b200.synthetic import make_problem
from tests.helpers import gpu_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
p = make_problem("S", n=n, N=N)
g = gpu_model(p, gradients=False)
g.predict(p["Xc"])
t = np.loadtxt(out, dtype=np.int64)
t0 = t[0, 0]
print("row   K*    wait_acc  recomb  diag  digits | row_total | mma_start-after-recomb(prev)  mma_busy  | gap acc_full->next K* start")
tot = 0
for i in range(len(t)):
    r = t[i]
    kst, kdone, accf, rec, diag, pub, m0, m1 = r[:8]
    if i == 0:
        print("%3d %6d %8s %7s %6d %6d | %8d" % (i, kdone - kst, "-", "-", diag - rec, pub - diag, pub - kst))
        continue
    prev_pub = t[i - 1, 5]
    print("%3d %6d %8d %7d %6d %6d | %8d | mma: start %+7d after prev recomb, busy %6d, done %+6d vs K* done"
          % (i, kdone - kst, accf - kdone, rec - accf, diag - rec, pub - diag, pub - prev_pub, m0 - t[i - 1, 3], m1 - m0, m1 - kdone))
    print("      K* done per warp (vs warp 2):", " ".join("%+6d" % (x - kdone) for x in r[16:24]), "| TMEM released per warp (vs warp 2):",
          " ".join("%+6d" % (x - rec) for x in r[8:16]), "| mma start vs LAST release %+6d" % ((t[i + 1, 6] - max(r[8:16])) if i + 1 < len(t) else 0))
print("item total clocks:", t[-1, 5] - t0)
