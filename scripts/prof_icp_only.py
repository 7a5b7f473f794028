"""Profiling target: one pair, the inner loop only (3 launches of icp_persistent_kernel; profile the last).
python scripts/prof_icp_only.py [n] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = synth.make_pair(n, with_clouds=False)
ctx = P.Context(0)
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
for _ in range(3):
    r = ctx.icp_run(P.icp_params(max_iter=iters, force_iters=1))
    print("icp ms", r["device_ms"], "kernel ms", r["kernel_ms"], r["grid_blocks"])
