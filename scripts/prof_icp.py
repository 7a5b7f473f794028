"""Profiling driver: 1M pair, a few stand-alone NN batches and one 50-iteration inner loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = synth.make_pair(n, with_clouds=False)
ctx = P.Context(0)
if len(sys.argv) > 3:
    ctx.set_cells_per_point(float(sys.argv[3]))
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
for _ in range(3):
    ctx.nn(d["ct2"]); print("nn ms", ctx.last_device_ms())
for _ in range(2):
    r = ctx.icp_run(P.icp_params(max_iter=iters, force_iters=1)); print("icp ms", r["device_ms"], r["grid_blocks"])
