import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
d = synth.make_pair(n)
ctx = P.Context(0)
for rep in range(2):
    ctx.upload_pair(d)
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    g = ctx.piecewise_icp(pp, 1, 0.05)
    print("rep", rep, "outer iters", g["n_outer"], "device ms", round(g["device_ms"], 3), [round(s.device_ms, 3) for s in g["stats"]])
