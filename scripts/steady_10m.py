"""Steady-state cost per inner iteration at 10M centroids (BASELINE configs[4])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
d = synth.make_pair(n, with_clouds=False)
ctx = P.Context(0)
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
t = {}
for it in (1, 2, 3, 4, 5, 6, 51, 151, 251):
    t[it] = min(ctx.icp_run(P.icp_params(max_iter=it, force_iters=1))["device_ms"] for _ in range(2))
m = len(d["ct2"])
print("n=%d it1 %.2f ms (+%.2f +%.2f +%.2f +%.2f +%.2f ms) it51 %.2f | steady us/iter %.1f = %.2f TB/s at 80 B/corr, %.2f TB/s algorithmic (48 B)" % (
    m, t[1], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[5]-t[4], t[6]-t[5], t[51], (t[251]-t[151])*10, 80*m/((t[251]-t[151])/100*1e-3)/1e12, 48*m/((t[251]-t[151])/100*1e-3)/1e12))
