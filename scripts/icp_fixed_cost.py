import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
if os.environ.get('PWICP_LIB'):
    P._lib = P.load_library(os.environ['PWICP_LIB'])
ctx = P.Context(0)
for n in (113664, 1000000):
    d = synth.make_pair(n, with_clouds=False)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])
    ctx.icp_run(P.icp_params(max_iter=5, force_iters=1))
    a = ctx.icp_run(P.icp_params(max_iter=11, force_iters=1))
    b = ctx.icp_run(P.icp_params(max_iter=51, force_iters=1))
    print(f"n={len(d['ct2'])} grid={b['grid_blocks']} per-iteration us = {(b['device_ms']-a['device_ms'])/40*1e3:.1f}")
