import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = synth.make_pair(n, with_clouds=False)
ctx = P.Context(0)
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
for it in (1, 1, 2, 3, 6, 11, 21, 51):
    r = ctx.icp_run(P.icp_params(max_iter=it, force_iters=1)); print("iters", it, "icp ms", round(r["device_ms"], 3))
