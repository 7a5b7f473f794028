import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
d = synth.make_pair(300000)
ctx = P.Context(0)
ctx.upload_pair(d)
pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
for _ in range(2):
    ctx.nn(d["cloud2"], P.TGT_CLOUD1); print("all cloud2 -> cloud1 ms", ctx.last_device_ms())
st = P.State(0.05, 0, 0, 0, 0)
T, V, flags, stats = ctx.single_iteration(pp, st)
print("iter ms", stats.device_ms, "stable", stats.n_stable)
pid = np.repeat(np.arange(len(d["ct2"])), 8)
sub = d["cloud2"][flags[pid] == 1]
for _ in range(2):
    ctx.nn(sub, P.TGT_CLOUD1); print("stable subset", len(sub), "ms", ctx.last_device_ms())
uns = d["cloud2"][flags[pid] == 0]
ctx.nn(uns, P.TGT_CLOUD1); print("unstable subset", len(uns), "ms", ctx.last_device_ms())
import time
t0 = time.time(); v = ctx.percentile_nn(d["cloud1"], d["cloud2"], 0.75); print("percentile_nn api s", time.time() - t0, v)
t0 = time.time(); v = ctx.percentile_nn(d["cloud1"], d["cloud2"], 0.75); print("percentile_nn api s", time.time() - t0, v)
