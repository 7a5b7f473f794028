"""A/B harness: per-iteration cost of the inner loop for several builds of libpwicp on the SAME box.
usage: python scripts/ab.py libA.so libB.so ...   (each runs in its own process)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
P._lib = P.load_library(sys.argv[1])
ctx = P.Context(0)
out = []
for n in (113664, 1000000):
    d = synth.make_pair(n, with_clouds=False)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])
    for _ in range(10):                      # clocks up, caches and allocations warm
        ctx.icp_run(P.icp_params(max_iter=300, force_iters=1))
    t = {}
    for it in (1, 2, 3, 4, 5, 6, 51, 520, 1020):
        t[it] = min(ctx.icp_run(P.icp_params(max_iter=it, force_iters=1))["device_ms"] for _ in range(4))
    nn_ms = []
    for _ in range(3):
        ctx.nn(d["ct2"]); nn_ms.append(ctx.last_device_ms())
    out.append("n=%%d: nn %%.3f ms, it1 %%.3f ms (+%%.0f +%%.0f +%%.0f +%%.0f +%%.0f us), it51 %%.3f | steady us/iter %%.2f" %% (
        len(d["ct2"]), min(nn_ms), t[1], (t[2]-t[1])*1e3, (t[3]-t[2])*1e3, (t[4]-t[3])*1e3, (t[5]-t[4])*1e3, (t[6]-t[5])*1e3, t[51], (t[1020]-t[520])/500*1e3))
print(os.path.basename(sys.argv[1]), " || ".join(out))
''' % ROOT
for rep in range(2):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(lib)])
